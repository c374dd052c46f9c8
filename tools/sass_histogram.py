"""Opcode histogram of the tensor-core / TMA / TMEM instructions per kernel in the shipped library:
    python tools/sass_histogram.py [outfitx_b200/libofx.so] > profiles/r2_sass_opcodes.txt
(cuobjdump -sass; UTCHMMA = tcgen05.mma kind::f16, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st,
UTMALDG/UTMASTG = TMA bulk tensor load/store, UBLKCP = cp.async.bulk, SYNCS = mbarrier, HMMA = mma.sync.)"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "outfitx_b200/libofx.so"
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA",
        "FFMA", "MUFU", "REDG", "ATOMG", "DFMA"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, hist, total = None, {}, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn)[:100]
        hist[fn] = collections.Counter(); total[fn] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        total[fn] += 1
        op = m.group(1)
        for w in WANT:
            if op.startswith(w):
                hist[fn][w] += 1
print(f"{'kernel':100s} {'instrs':>7s}  " + " ".join(f"{w:>8s}" for w in WANT))
for fn in sorted(hist):
    print(f"{fn:100s} {total[fn]:7d}  " + " ".join(f"{hist[fn][w]:8d}" if hist[fn][w] else f"{'.':>8s}" for w in WANT))
