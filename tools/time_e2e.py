"""Times HostScoringPipeline on configs[1] from pinned host buffers: padded (B,16,dpm) tensors against the packed
valid-rows layout, and where the step's time goes (copies alone, scoring alone)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from outfitx_b200.pipeline import HostScoringPipeline, pack_valid_rows
dev = torch.device("cuda", 0)
model, sd = bench.make_model(dev)
B = 8192
img, txt, mask, text, cand, lengths = bench.make_cp_inputs(B, dev, seed=1000)
host = {k: v.cpu().pin_memory() for k, v in dict(img=img, txt=txt, mask=mask, text=text, cand=cand).items()}
ir, tr, lens = pack_valid_rows(host["img"], host["txt"], host["mask"])
ir, tr = ir.pin_memory(), tr.pin_memory()
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for chunk, tail in ((2048, 2048), (2048, 1024), (1536, 1536), (2731, 2731), (2560, 2560)):
    pipe = HostScoringPipeline(model, chunk=chunk); pipe.tail_min = tail
    b = t(lambda: pipe.score_packed(ir, tr, lens, host["text"], host["cand"]))
    print(f"chunk {chunk} tail_min {tail}: plan {[hi - lo for lo, hi in pipe._plan(B, True)]}  packed {b:.2f} ms ({B/b*1e3:,.0f}/s)", flush=True)
for chunk in (2048,):
    pipe = HostScoringPipeline(model, chunk=chunk)
    a = t(lambda: pipe.score(host["img"], host["txt"], host["mask"], host["text"], host["cand"]))
    b = t(lambda: pipe.score_packed(ir, tr, lens, host["text"], host["cand"]))
    print(f"chunk {chunk}: padded {a:.2f} ms ({B/a*1e3:,.0f}/s)   packed {b:.2f} ms ({B/b*1e3:,.0f}/s)", flush=True)
# pieces
dst = torch.empty_like(ir, device=dev)
print(f"H2D of the packed rows of one modality ({ir.numel()*4/1e6:.0f} MB): {t(lambda: dst.copy_(ir, non_blocking=True)):.2f} ms")
dst2 = torch.empty_like(host['img'], device=dev)
print(f"H2D of the padded tensor of one modality ({host['img'].numel()*4/1e6:.0f} MB): {t(lambda: dst2.copy_(host['img'], non_blocking=True)):.2f} ms")
t0 = time.perf_counter()
for _ in range(20):
    lens64 = lens.to(torch.int64); off = torch.zeros(B + 1, dtype=torch.int64); torch.cumsum(lens64, 0, out=off[1:])
print(f"host-side collate arithmetic: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
# where the packed step's time goes: the copies alone (same chunking, no scoring) and the scoring alone (chunks
# already on the device)
pipe = HostScoringPipeline(model, chunk=2048)
pipe.score_packed(ir, tr, lens, host["text"], host["cand"])
off = torch.zeros(B + 1, dtype=torch.int64); torch.cumsum(lens.to(torch.int64), 0, out=off[1:])
def copies_only():
    with torch.cuda.stream(pipe.copy_stream):
        for i, lo in enumerate(range(0, B, 2048)):
            hi = lo + 2048; s = i % pipe.n_slots; r0, r1 = int(off[lo]), int(off[hi])
            pipe._stage_rows(s, "pimg", ir[r0:r1], 2048 * 16); pipe._stage_rows(s, "ptxt", tr[r0:r1], 2048 * 16)
            pipe._stage(s, "text", host["text"][lo:hi]); pipe._stage(s, "cand", host["cand"][lo:hi])
    pipe.copy_stream.synchronize()
print(f"the four chunks' H2D copies alone: {t(copies_only):.2f} ms")
dev_chunks = [dict(img=img[lo:lo + 2048].contiguous(), txt=txt[lo:lo + 2048].contiguous(), mask=mask[lo:lo + 2048].contiguous(),
                   text=text[lo:lo + 2048].contiguous(), cand=cand[lo:lo + 2048].contiguous()) for lo in range(0, B, 2048)]
def compute_only():
    for c in dev_chunks:
        enc = {"image_embeddings": c["img"], "text_embeddings": c["txt"]}
        model.score_cp(outfit_mask=c["mask"], encoder_input_dict=enc)
        model.score_fitb(outfit_mask=c["mask"], target_item_text_embedding=c["text"], candidate_item_embedding=c["cand"],
                         encoder_input_dict=enc)
print(f"scoring of four device-resident chunks of 2048 (eager launches): {t(compute_only):.2f} ms")
enc = {"image_embeddings": img, "text_embeddings": txt}
def compute_whole():
    model.score_cp(outfit_mask=mask, encoder_input_dict=enc)
    model.score_fitb(outfit_mask=mask, target_item_text_embedding=text, candidate_item_embedding=cand, encoder_input_dict=enc)
print(f"scoring of the whole 8192-outfit batch: {t(compute_whole):.2f} ms")
