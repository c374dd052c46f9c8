"""A/B of the host pipeline's H2D modes on configs[1] (8192 outfits, CP + FITB): whole padded tensors by DMA
vs valid slots fetched in place by the SMs.  Usage: python tools/time_e2e.py [chunk]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from outfitx_b200.pipeline import HostScoringPipeline

dev = torch.device("cuda", 0)
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
model, _ = bench.make_model(dev)
B = 8192
img, txt, mask, text, cand, lengths = bench.make_cp_inputs(B, dev, 1000)
host = [t.cpu().pin_memory() for t in (img, txt, mask, text, cand)]
valid = float((~host[2]).float().mean())
for mode in (False, True, False, True):
    pipe = HostScoringPipeline(model, chunk=chunk, fetch_valid_only=mode)
    out = None
    for _ in range(3):
        out = pipe.score(*host, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        pipe.score(*host, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"fetch_valid_only={mode}  chunk {chunk}  valid fraction {valid:.3f}  {ms:.2f} ms/step  {B / ms * 1e3:,.0f} outfits/s")
