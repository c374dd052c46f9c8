#!/usr/bin/env python
"""Benchmark of the outfit-scoring hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on rank 0.  Primary metric: CP outfits/s on BASELINE.json configs[1]
(CP + FITB(4 candidates) scoring, 8192 outfits per GPU, bf16, mean fusion; weak scaling, no
collective).  The same line carries a `cir` object: queries/s of the exact top-10 search over a
10 M-item gallery sharded across the N GPUs with one NCCL all-gather (configs[3], strong
scaling).  `--impl reference` times the reference's CPU path (oracle/torch_port.py: the stock
torch modules the reference itself is built from -- the reference is pure Python and cannot
travel to the GPU box) on the host cores for the same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line: everything libraries print to fd 1 (e.g. NCCL's version
# banner) is sent to stderr instead, and the result line goes to the saved descriptor.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

from outfitx_b200 import synth  # noqa: E402

D_MODEL, D_EMBED, DPM, F_FFN, N_LAYERS = 512, 1024, 512, 2024, 6
N_CAND, TOPK = 4, 10
CIR_TRAFFIC_BYTES = 32338742069    # same for the main sweep of the search at 8192 queries x 10 M rows on one GPU
                                   # (profiles/r1_ncu_search_paced.txt; algorithmic: the packed gallery once = 21.76 GB)
FFN_TRAFFIC_BYTES = 380563456      # dram__bytes_read.sum + dram__bytes_write.sum of one ffn_block_kernel
                                   # launch (LN-emitting form) at 82158 rows (profiles/r1_ncu_ffn_block_ln.txt)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"burst": float(p["bf16_tflops"]), "sustained": float(p["bf16_tflops_sustained"]),
                "hbm": float(p["hbm_gbs"]), "source": "measured"}
    except Exception:
        return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


def flops_alg(n, dm=D_MODEL, f=F_FFN, de=D_EMBED, task="cp"):
    """Minimum exact work per outfit with n valid items: layers 0-4 dense over the
    1+n valid tokens, layer 5 pruned to the prefix-token query row; padding never counted."""
    n = np.asarray(n, np.float64)
    s = 1.0 + n
    dense = 5 * s * (8 * dm * dm + 4 * dm * f + 4 * s * dm)
    last = s * 4 * dm * dm + 2 * dm * dm + 4 * s * dm + 2 * dm * dm + 4 * dm * f
    head = 2 * dm if task == "cp" else 2 * dm * de + 3 * N_CAND * de
    return dense + last + head


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw)}


def timed(fn, steps, warmup, dist_ok, dev, sampler=None):
    """W untimed + exactly K timed steps, barrier + synchronize on both sides, CUDA events on the
    launching stream, MAX over ranks.  -> total ms."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    if dist_ok:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.__enter__()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    if dist_ok:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.__exit__()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist_ok:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def time_ffn_block(L, rows, dev, reps=20):
    """ofx_ffn_block_ln_bf16 -- the form layers 0-4 of the step launch: the fused FFN block that also emits
    norm1 of the next layer -- on `rows` token rows (d_model 512, d_ffn 2024 padded to 2048), inputs rotated
    over 4 buffers (4 x rows x 2 KB > L2).  Algorithmic flops = 4 * rows * 512 * 2024 (the padded columns and
    the LayerNorms are not counted); algorithmic HBM bytes = rows * 5 KB (fp32 row in, fp32 row out, bf16
    LayerNorm row out)."""
    from outfitx_b200 import _lib
    g = torch.Generator(device=dev).manual_seed(11)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    ln_w, ln_b = 1.0 + 0.1 * r(D_MODEL), 0.1 * r(D_MODEL)
    w1 = (r(2048, D_MODEL) / D_MODEL ** 0.5); w1[F_FFN:] = 0
    w2 = (r(D_MODEL, 2048) / 2048 ** 0.5); w2[:, F_FFN:] = 0
    b1 = 0.1 * r(2048); b1[F_FFN:] = 0
    b2 = 0.1 * r(D_MODEL)
    w1, w2 = w1.to(torch.bfloat16).contiguous(), w2.to(torch.bfloat16).contiguous()
    bufs = [r(rows, D_MODEL) for _ in range(4)]
    st = torch.cuda.current_stream(dev).cuda_stream

    h_next = torch.empty(rows, D_MODEL, device=dev, dtype=torch.bfloat16)

    def run(x):
        _lib.check(L.ofx_ffn_block_ln_bf16(x.data_ptr(), rows, D_MODEL, 2048, ln_w.data_ptr(), ln_b.data_ptr(),
                                           w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                           h_next.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), st))
    for x in bufs:
        run(x)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        run(bufs[i % 4])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    flops = 4.0 * rows * D_MODEL * F_FFN
    return {"kernel": "ffn_block_kernel", "rows": rows, "us_per_launch": ms * 1e3, "flops_per_launch": flops,
            "tflops": flops / (ms * 1e-3) / 1e12}


def make_cp_inputs(batch, dev, seed):
    """configs[1] inputs: raw (un-normalised) 512-d CLIP image / text embeddings per item,
    n ~ U{2..16} valid items left-aligned, 256-d target text, 4 FITB candidates of 1024-d."""
    g = torch.Generator(device=dev).manual_seed(seed)
    img = torch.randn(batch, 16, DPM, device=dev, generator=g)
    txt = torch.randn(batch, 16, DPM, device=dev, generator=g)
    lengths = synth.make_lengths(batch, seed + 1)
    mask = torch.from_numpy(synth.make_mask(lengths)).to(dev)
    text = torch.nn.functional.normalize(torch.randn(batch, D_MODEL // 2, device=dev, generator=g), dim=-1)
    cand = torch.randn(batch, N_CAND, 2, DPM, device=dev, generator=g)
    cand = torch.nn.functional.normalize(cand, dim=-1).reshape(batch, N_CAND, 2 * DPM).contiguous()
    return img, txt, mask, text, cand, lengths


def make_model(dev):
    import outfitx_b200 as o
    cfg = o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type="clip", aggregation_method="mean"))
    m = o.OutfitX(cfg, precision="bf16")
    sd = synth.make_state_dict(D_MODEL, D_EMBED, seed=0)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m.to(dev), sd


def make_gallery_shard(n_total, rank, world, dev, chunk=500_000):
    """Rows [lo, hi) of the synthetic gallery: per-modality-normalised halves (|g|^2 = 2)."""
    from outfitx_b200.search import shard_rows
    lo, hi = shard_rows(n_total, rank, world)
    out = torch.empty(hi - lo, 2 * DPM, dtype=torch.float32, device=dev)
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        g = torch.Generator(device=dev).manual_seed(5 * 2 ** 32 + c0)
        x = torch.randn(c1 - c0, 2, DPM, device=dev, generator=g)
        out[c0 - lo:c1 - lo] = torch.nn.functional.normalize(x, dim=-1).reshape(c1 - c0, 2 * DPM)
    return out, lo


def cpu_cp_sample(sd, sample, seed, reps=2):
    """Reference CPU path (stock torch modules, fp32) on a bounded sample of the CP+FITB workload."""
    from oracle import torch_port
    torch.set_num_threads(os.cpu_count() or 1)
    port = torch_port.ReferencePort.from_numpy(sd)
    img, txt = synth.make_modalities(sample, DPM, seed)
    lengths = synth.make_lengths(sample, seed + 1)
    mask = synth.make_mask(lengths)
    emb = torch.from_numpy(synth.fuse(img, txt, "mean"))
    maskt = torch.from_numpy(mask)
    text = torch.from_numpy(synth.make_text_prefix(sample, D_MODEL // 2, seed + 2))
    cand = torch.from_numpy(synth.make_items(sample * N_CAND, DPM, seed + 3).reshape(sample, N_CAND, 2 * DPM))
    best = float("inf")
    for r in range(reps + 1):
        t0 = time.perf_counter()
        logits = port.cp(emb, maskt)
        torch.sigmoid(logits.float())
        q = port.cir(emb, maskt, text)
        torch_port.fitb(q, cand)
        dt = time.perf_counter() - t0
        if r > 0:
            best = min(best, dt)
    return sample / best, torch.get_num_threads()


def cpu_cir_sample(nq_s=256, n_s=200_000, reps=2):
    """Reference idiom topk(cdist(Q, G)) chunked, on a bounded slice; cost is linear in nq*N."""
    from oracle import torch_port
    torch.set_num_threads(os.cpu_count() or 1)
    gal = torch.from_numpy(synth.make_items(n_s, DPM, seed=5))
    q = torch.from_numpy(synth.make_queries(nq_s, 2 * DPM, seed=6))
    best = float("inf")
    for r in range(reps + 1):
        t0 = time.perf_counter()
        torch_port.search_cdist_topk(q, gal, k=TOPK)
        dt = time.perf_counter() - t0
        if r > 0:
            best = min(best, dt)
    return nq_s * n_s / best  # (query, item) pairs per second


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, all host threads."""
    if rank != 0:
        return
    sd = synth.make_state_dict(D_MODEL, D_EMBED, seed=0)
    sample = args.cpu_sample
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.warmup + args.steps):
        v, threads = cpu_cp_sample(sd, sample, seed=1, reps=1)
        vals.append(v)
        if time.perf_counter() - t0 > 150:
            break
    vals = vals[min(args.warmup, len(vals) - 1):]
    value = float(np.mean(vals))
    pairs = cpu_cir_sample()
    cir_q = pairs / args.cir_rows
    line = {
        "impl": "reference", "metric": "CP outfits/sec", "value": value, "unit": "outfits/s",
        "n_gpus": world, "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * sample / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "configs[1]: CP + FITB(4 cand) scoring, mean fusion, d_model 512",
                   "sample": f"{sample} outfits per step (bounded sample of the 8192-outfit batch)"},
        "cpu_baseline": {"value": value, "unit": "outfits/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} outfits, CP + FITB, fp32 stock-torch port of the reference"},
        "e2e": {"value": value, "unit": "outfits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cir": {"metric": f"CIR queries/sec top-{TOPK} over {args.cir_rows} items", "value": cir_q,
                "unit": "queries/s", "kind": "port",
                "sample": "topk(cdist) on 256 q x 200k items, scaled linearly in nq*N"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8192, help="outfits per GPU per step (configs[1])")
    ap.add_argument("--cir-rows", type=int, default=10_000_000)
    ap.add_argument("--cir-queries", type=int, default=8192)
    ap.add_argument("--cir-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--no-cir", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=512)
    ap.add_argument("--e2e-chunk", type=int, default=2048, help="outfits per device chunk of the host pipeline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch.distributed as dist
    from outfitx_b200 import _lib
    from outfitx_b200.search import Gallery, ShardedSearch
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_ok = world > 1
    if dist_ok:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    pk = peaks()

    # ------------------------------------------------------------------ CP + FITB (primary)
    model, sd = make_model(dev)
    B = args.batch
    img, txt, mask, text, cand, lengths = make_cp_inputs(B, dev, seed=1000 + rank)
    enc = {"image_embeddings": img, "text_embeddings": txt}
    state = {}

    def cp_step():
        state["probs"] = model.score_cp(outfit_mask=mask, encoder_input_dict=enc)
        state["fitb"] = model.score_fitb(outfit_mask=mask, target_item_text_embedding=text,
                                         candidate_item_embedding=cand, encoder_input_dict=enc)

    cp_step()
    torch.cuda.synchronize(dev)
    n0 = L.ofx_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    cp_ms = timed(cp_step, args.steps, args.warmup, dist_ok, dev, sampler)
    launches_cp = (L.ofx_launch_count() - n0) * args.steps // (args.steps + args.warmup)
    cp_value = world * B * args.steps / (cp_ms * 1e-3)
    flops_step = float(flops_alg(lengths, task="cp").sum() + flops_alg(lengths, task="fitb").sum())
    cp_tflops = flops_step * args.steps / (cp_ms * 1e-3) / 1e12     # per GPU (every rank does B outfits)
    clocks = sampler.summary() if sampler else None

    # dominant kernel, timed alone with CUDA events on its launching stream: the fused FFN block
    # (ffn_block_kernel, 46 % of the step) on this batch's valid-token count
    dom = time_ffn_block(L, int(B + lengths.sum()), dev)

    # end to end through the public API with HOST buffers (pinned), copies inside the timed region
    host = {k: v.cpu().pin_memory() for k, v in
            dict(img=img, txt=txt, mask=mask, text=text, cand=cand).items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    res_host = {"probs": torch.empty(B, dtype=torch.float32).pin_memory(),
                "pred": torch.empty(B, dtype=torch.int64).pin_memory()}
    d2h = sum(v.numel() * v.element_size() for v in res_host.values())

    from outfitx_b200.pipeline import HostScoringPipeline
    pipe = HostScoringPipeline(model, chunk=args.e2e_chunk)
    res_host = {"probs": res_host["probs"], "pred": res_host["pred"]}

    def cp_e2e_step():
        # public API on HOST buffers: chunked H2D on a copy stream overlapped with scoring, results
        # back in pinned host memory; returns when they are valid (outfitx_b200/pipeline.py)
        pipe.score(host["img"], host["txt"], host["mask"], host["text"], host["cand"], out=res_host)

    e2e_ms = timed(cp_e2e_step, args.steps, 2, dist_ok, dev)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)

    # the same step with the collate on the device: item tables resident in HBM, only
    # item ids / masks / text prefixes / candidate ids cross PCIe.  Reported beside `e2e`, not as it.
    n_table = 200_000
    g2 = torch.Generator(device=dev).manual_seed(77 + rank)
    tab_img = torch.randn(n_table, DPM, device=dev, generator=g2)
    tab_txt = torch.randn(n_table, DPM, device=dev, generator=g2)
    tab_cand = torch.nn.functional.normalize(torch.randn(n_table, 2, DPM, device=dev, generator=g2), dim=-1)
    tab_cand = tab_cand.reshape(n_table, 2 * DPM).contiguous()
    ids_host = torch.randint(0, n_table, (B, 16), dtype=torch.int32).pin_memory()
    cids_host = torch.randint(0, n_table, (B, N_CAND), dtype=torch.int32).pin_memory()
    ids_h2d = ids_host.numel() * 4 + cids_host.numel() * 4 + host["mask"].numel() + host["text"].numel() * 4

    pipe_ids = HostScoringPipeline(model, chunk=B)   # ids are 64 B per outfit: nothing to overlap, one chunk

    def cp_e2e_ids_step():
        pipe_ids.score_ids(ids_host, host["mask"], tab_img, tab_txt, host["text"], cids_host, tab_cand, out=res_host)

    e2e_ids_ms = timed(cp_e2e_ids_step, args.steps, 2, dist_ok, dev)
    e2e_ids_value = world * B * args.steps / (e2e_ids_ms * 1e-3)
    del tab_img, tab_txt, tab_cand

    # ------------------------------------------------------------------ CIR (secondary, sharded)
    cir = None
    if not args.no_cir:
        k_steps = args.cir_steps or min(args.steps, 5)
        rows, lo = make_gallery_shard(args.cir_rows, rank, world, dev)
        gal = Gallery.build(rows, id_offset=lo, keep_fp32=True)
        g = torch.Generator(device=dev).manual_seed(6)
        queries = torch.randn(args.cir_queries, D_EMBED, device=dev, generator=g) * 0.05
        searcher = ShardedSearch()

        def cir_step():
            state["cir"] = searcher.search(queries, gal, TOPK, "l2", True)

        cir_step()
        torch.cuda.synchronize(dev)
        n1 = L.ofx_launch_count()
        cir_ms = timed(cir_step, k_steps, 3, dist_ok, dev)
        launches_cir = (L.ofx_launch_count() - n1) * k_steps // (k_steps + 3)
        cir_value = args.cir_queries * k_steps / (cir_ms * 1e-3)
        cir_flops = 2.0 * args.cir_queries * gal.n_rows * D_EMBED           # this rank's shard
        cir_tf = cir_flops * k_steps / (cir_ms * 1e-3) / 1e12

        q_host = queries.cpu().pin_memory()
        idx_host = torch.empty(args.cir_queries, TOPK, dtype=torch.int64).pin_memory()

        def cir_e2e_step():
            q = q_host.to(dev, non_blocking=True)
            idx, _ = searcher.search(q, gal, TOPK, "l2", True)
            idx_host.copy_(idx, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()

        cir_e2e_ms = timed(cir_e2e_step, k_steps, 1, dist_ok, dev)
        cir = {
            "metric": f"CIR queries/sec top-{TOPK} over {args.cir_rows} items", "value": cir_value,
            "unit": "queries/s", "scaling": "strong", "steps": k_steps, "ms_per_step": cir_ms / k_steps,
            "config": {"workload": f"configs[3]: {args.cir_queries} queries, exact top-{TOPK} (L2) over a "
                                   f"{args.cir_rows}-item 1024-d gallery row-sharded over {world} GPU(s), "
                                   "one NCCL all-gather + merge", "rows_per_gpu": gal.n_rows,
                       "l2": "gallery shard (bf16) larger than L2"},
            "roofline": {"bound": "tensor", "achieved": cir_tf, "peak": pk["sustained"], "unit": "TFLOP/s",
                         "frac": cir_tf / pk["sustained"],
                         "traffic": CIR_TRAFFIC_BYTES if (world == 1 and args.cir_rows == 10_000_000 and args.cir_queries == 8192) else None,
                         "algorithmic_bytes_per_launch": gal.n_rows * (1024 + 64) * 2,
                         "kernel": "tc_kernel<256,6,2,SchedSearch,EpiTopK<32>,pair> (+ EpiBlockMax seeding launch, merge_rerank), per GPU",
                         "flops_per_launch": cir_flops, "peak_kind": f"sustained bf16, {pk['source']}"},
            "e2e": {"value": args.cir_queries * k_steps / (cir_e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": q_host.numel() * 4, "d2h_bytes_per_step": idx_host.numel() * 8},
            "gpu_launches": int(launches_cir),
        }
        del rows, gal

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            v, threads = cpu_cp_sample(sd, args.cpu_sample, seed=1)
            cpu = {"value": v, "unit": "outfits/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_sample} outfits (of the 8192-outfit batch), CP + FITB, fp32, "
                             "stock-torch port of the reference (oracle/torch_port.py)"}
            if cir is not None:
                pairs = cpu_cir_sample()
                cir["cpu_baseline"] = {"value": pairs / args.cir_rows, "unit": "queries/s",
                                       "cores": threads, "kind": "port",
                                       "sample": "topk(cdist(Q,G)) on 256 q x 200k items, scaled linearly in nq*N"}
        line = {
            "metric": "CP outfits/sec", "value": cp_value, "unit": "outfits/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cp_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "configs[1]: CP + FITB(4 candidates) scoring, 8192 outfits per GPU, bf16, "
                                   "mean-aggregation fusion (d_model 512, 6 layers, 16 heads, d_ffn 2024), "
                                   "n ~ U{2..16} items per outfit", "batch_per_gpu": B,
                       "l2": "inputs (805 MB per step) larger than L2", "parallelism": f"dp{world} by outfit, no collective"},
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": dom["tflops"], "peak": pk["burst"], "unit": "TFLOP/s",
                         "frac": dom["tflops"] / pk["burst"], "traffic": FFN_TRAFFIC_BYTES,
                         "kernel": "ffn_block_kernel (fused LN2 + linear1 + mish + linear2 + residual + norm1 of the "
                                   "next layer; dominant kernel of the step), timed alone with CUDA events",
                         "rows_per_launch": dom["rows"], "us_per_launch": dom["us_per_launch"],
                         "flops_per_launch": dom["flops_per_launch"],
                         "algorithmic_bytes_per_launch": dom["rows"] * 5120,
                         "peak_kind": f"burst bf16, {pk['source']}",
                         "traffic_note": "dram read+write of one launch at 82k rows from profiles/ (ncu --set full)"},
            "roofline_step": {"bound": "tensor", "achieved": cp_tflops, "peak": pk["burst"], "unit": "TFLOP/s",
                              "frac": cp_tflops / pk["burst"], "flops_per_step": flops_step,
                              "kernel": "whole step, all launches (flops_alg: minimum exact work, SURVEY 8d)"},
            "e2e": {"value": e2e_value, "unit": "outfits/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "outfitx_b200.pipeline.HostScoringPipeline.score: fp32 (B,16,512) image + text embeddings "
                           "from pinned host memory, chunked H2D overlapped with scoring"},
            "e2e_device_collate": {"value": e2e_ids_value, "unit": "outfits/s", "h2d_bytes_per_step": ids_h2d,
                                   "d2h_bytes_per_step": d2h,
                                   "api": "HostScoringPipeline.score_ids: item ids from host, 200k-item embedding "
                                          "tables resident in HBM (SURVEY.md N2)"},
            "gpu_launches": int(launches_cp),
            "cpu_baseline": cpu,
            "cir": cir,
        }
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if dist_ok:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
