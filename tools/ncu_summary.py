"""Key metrics of every kernel in an .ncu-rep (run where ncu is installed, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more-metric-prefixes...]"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_xu", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "lts__t_sectors_srcunit_tex.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__average_warp"]
rep = sys.argv[1]
want = WANT + sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = [i for i, h in enumerate(hdr) if any(h.startswith(w) for w in want)]
kn = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[kn][:110])
    for i in idx:
        print(f"  {hdr[i]:78s} {units[i]:16s} {r[i]}")
