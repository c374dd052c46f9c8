"""BASELINE.json configs[4]: large-encoder variant (d_model 1024 = clip + concat, 6 layers, 16 items per outfit),
CP throughput sweep over the batch size on 1 B200, inputs resident in HBM.  One JSON line per batch size."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import outfitx_b200 as o
from outfitx_b200 import synth
import bench

dev = torch.device("cuda", 0)
cfg = o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type="clip", aggregation_method="concat"))
m = o.OutfitX(cfg, precision="bf16")
sd = synth.make_state_dict(1024, 1024, seed=0)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
m = m.to(dev)
peak = bench.peaks()["burst"]
for B in (256, 1024, 4096, 8192, 16384, 32768):
    g = torch.Generator(device=dev).manual_seed(B)
    emb = torch.nn.functional.normalize(torch.randn(B, 16, 2, 512, device=dev, generator=g), dim=-1).reshape(B, 16, 1024)
    mask = torch.zeros(B, 16, dtype=torch.bool, device=dev)          # n = 16 valid items
    for _ in range(3):
        m.score_cp(emb, mask)
    torch.cuda.synchronize()
    reps = max(3, min(20, 65536 // B))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m.score_cp(emb, mask)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = float(bench.flops_alg(np.full(B, 16), dm=1024, task="cp").sum())
    print(json.dumps({"workload": "configs[4] large encoder CP", "batch": B, "ms": ms, "outfits_per_s": B / ms * 1e3,
                      "tflops_alg": fl / ms / 1e9, "frac_of_burst_peak": fl / ms / 1e9 / peak}))
