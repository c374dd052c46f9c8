"""Fixed workload for ncu: 3 CP passes of the large encoder (configs[4]: d_model 1024, 16 items per outfit).
Usage: [ncu ...] python tools/prof_large.py [--batch 2048] [--task cp|cir]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2048)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model, _ = bench.make_model(dev, 1024)
g = torch.Generator(device=dev).manual_seed(a.batch)
emb = torch.nn.functional.normalize(torch.randn(a.batch, 16, 2, bench.DPM, device=dev, generator=g), dim=-1).reshape(a.batch, 16, 1024)
mask = torch.zeros(a.batch, 16, dtype=torch.bool, device=dev)
for _ in range(2):
    model.score_cp(emb, mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); model.score_cp(emb, mask); e1.record(); torch.cuda.synchronize()
print("large cp pass ms", e0.elapsed_time(e1))
