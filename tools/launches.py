"""Summarise an ncu --csv launch list: the last CP pass (from scan_valid_kernel on)."""
import csv, re, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
r=csv.DictReader(lines)
cur={}; order=[]
for row in r:
    key=row['ID']
    if key not in cur:
        cur[key]={'name':re.sub(r'\(CUtensor.*|\(ofx::A.*|\(const.*|\(Assem.*|\(Attn.*','',row['Kernel Name'])[:72]}; order.append(key)
    cur[key][row['Metric Name']]=float(row['Metric Value'].replace(',',''))
starts=[i for i,k in enumerate(order) if 'scan_valid' in cur[k]['name']]
s=starts[-1] if starts else 0
n=int(sys.argv[2]) if len(sys.argv)>2 else 24
def show(k):
    d=cur[k]
    g=lambda m: d.get(m,0.0)
    print(f"{d['name']:74s} {g('gpu__time_duration.sum')/1e3:8.1f}us rd {g('dram__bytes_read.sum')/1e6:7.1f}MB wr {g('dram__bytes_write.sum')/1e6:7.1f}MB lsu {g('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):5.1f}% inst {g('smsp__inst_executed.sum')/1e6:6.1f}M")
for k in order[s:s+n]: show(k)
print('...')
for k in order[-9:]: show(k)
print('pass total us', sum(cur[k]['gpu__time_duration.sum'] for k in order[s:])/1e3)
