"""Writes profiles/r2_traffic.json (read by bench.py's `traffic` fields) from ncu captures of the shipped kernels:
    python tools/make_traffic.py gpurun_out/r2ev
DRAM bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum of the named kernel (ncu --set full report,
or for the feed-forward block the in-step --metrics launch list, which holds the LN-emitting form bench.py times)."""
import csv, json, os, re, subprocess, sys

src = sys.argv[1]
out = {}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def rep_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
        yield d


def dram(d):
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v, u = d[m]
        tot += float(v.replace(",", "")) * UNIT[u]
    return int(tot)


for f in sorted(os.listdir(src)):
    m = re.match(r"search_(\d+)\.ncu-rep", f)
    if m:
        for d in rep_rows(os.path.join(src, f)):
            if "EpiTopK" in d["Kernel Name"][0]:
                out[f"search_sweep_{m.group(1)}"] = {
                    "dram_bytes": dram(d), "kernel": d["Kernel Name"][0][:80],
                    "us": float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}[d["gpu__time_duration.sum"][1]],
                    "source": f"ncu --set full --clock-control none of tools/prof_step.py --skip-cp --rows {m.group(1)} "
                              f"({4096 if m.group(1) == '1000000' else 8192} queries, top-10), summary in "
                              f"profiles/r2_ncu_search_{m.group(1)}.txt"}
lc = os.path.join(src, "launches_cp.csv")
if os.path.exists(lc):
    lines = [l for l in open(lc) if not l.startswith("==")]
    per = {}
    for row in csv.DictReader(lines):
        if "ffn_block_kernel" in row["Kernel Name"]:
            per.setdefault(row["ID"], {})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    big = [v for v in per.values() if v.get("gpu__time_duration.sum", 0) > 200e3]      # layers 0-4 (82158 rows), not the pruned layer 5
    tot = [v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for v in big]
    out["ffn_block_82158"] = {
        "dram_bytes": int(sum(tot) / len(tot)), "launches_averaged": len(tot),
        "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none launch list of "
                  "tools/prof_step.py --skip-cir (the LN-emitting form inside the CP pass, 82158 rows), "
                  "profiles/r2_launches_cp_pass.txt"}
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
