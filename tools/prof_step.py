"""Small fixed workload for ncu: 2 CP passes (8192 outfits, d_model 512, bf16) and 2 searches
(8192 queries x --rows gallery).  Usage: ncu ... python tools/prof_step.py [--rows N] [--skip-cp]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from outfitx_b200.search import Gallery, local_search  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--queries", type=int, default=8192)
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--skip-cp", action="store_true")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--skip-cir", action="store_true")
ap.add_argument("--k", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda", 0)
if not a.skip_cp:
    model, _ = bench.make_model(dev)
    img, txt, mask, text, cand, _ = bench.make_cp_inputs(a.batch, dev, 1000)
    enc = {"image_embeddings": img, "text_embeddings": txt}
    for _ in range(2):
        model.score_cp(outfit_mask=mask, encoder_input_dict=enc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        model.score_cp(outfit_mask=mask, encoder_input_dict=enc)
    e1.record()
    torch.cuda.synchronize()
    print("cp pass ms", e0.elapsed_time(e1) / a.reps)
if not a.skip_cir:
    rows, lo = bench.make_gallery_shard(a.rows, 0, 1, dev)
    gal = Gallery.build(rows)
    q = torch.randn(a.queries, 1024, device=dev, generator=torch.Generator(device=dev).manual_seed(6)) * 0.05
    for _ in range(2):
        local_search(q, gal, a.k, "l2", True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        local_search(q, gal, a.k, "l2", True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    from outfitx_b200.search import SearchStats
    print("search ms", ms, "TFLOP/s", 2.0 * a.queries * a.rows * 1024 / ms / 1e9,
          "uncertified", SearchStats.uncertified, "of", SearchStats.queries)
print("ok")
