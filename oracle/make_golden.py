"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the reference itself.

Run in the build container (where /root/reference exists):

    python -m oracle.make_golden

Every fixture stores the seeds that regenerate its inputs through ``outfitx_b200.synth``
(weights are never stored: 51 M parameters), a checksum of those inputs, and the outputs the
UNMODIFIED reference produced for them on CPU in fp32 eval mode with autocast off -- the
setting of the reference demo (``src/demo/app.py:128,176,212``).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from outfitx_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, fusion, d_model, batch): lengths are drawn by case_inputs below
CASES = [("concat1024", "concat", 1024, 8), ("mean512", "mean", 512, 8)]


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def case_inputs(method: str, batch: int, seed: int = 100):
    """Inputs of a golden case: ragged lengths incl. the edge cases 0, 1 and 16 items, plus
    one outfit whose valid items are NOT left-aligned (the model must not care)."""
    img, txt = synth.make_modalities(batch, 512, seed)
    lengths = synth.make_lengths(batch, seed + 1)
    lengths[:3] = (0, 1, 16)
    mask = synth.make_mask(lengths)
    mask[3] = np.array([1, 0, 1, 1, 0, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0], bool)
    emb = synth.fuse(img, txt, method)
    emb[mask] = 0.0
    d_model = emb.shape[-1]
    text = synth.make_text_prefix(batch, d_model // 2, seed + 2)
    cand = synth.make_items(batch * 4, 512, seed + 3).reshape(batch, 4, 1024)
    return img, txt, emb, mask, text, cand


def main():
    assert ref_shim.available(), "reference not mounted"
    os.makedirs(OUT, exist_ok=True)
    ox, cfgs, dts = ref_shim.load_reference()
    torch.set_grad_enabled(False)

    for name, method, d_model, batch in CASES:
        sd = synth.make_state_dict(d_model, 1024, seed=0)
        model = ref_shim.build_reference_model(method, sd)
        assert model.item_encoder.d_embed == d_model
        img, txt, emb, mask, text, cand = case_inputs(method, batch)
        t = torch.from_numpy
        logits = model(task=dts.OutfitCompatibilityPredictionTask,
                       outfit_embedding=t(emb), outfit_mask=t(mask))
        query = model(task=dts.OutfitComplementaryItemRetrievalTask, outfit_embedding=t(emb),
                      outfit_mask=t(mask), target_item_text_embedding=t(text))
        query_fitb = model(task=dts.OutfitFillInTheBlankTask, outfit_embedding=t(emb),
                           outfit_mask=t(mask), target_item_text_embedding=t(text))
        assert torch.equal(query, query_fitb)
        # caller idioms (trainers / demo)
        probs = torch.sigmoid(logits.float())                      # cp trainer :408
        dists = torch.cdist(query.unsqueeze(1), t(cand), p=2).squeeze(1)  # fitb trainer :52
        fitb_idx = torch.argmin(dists, dim=-1)                      # :53
        np.savez_compressed(
            os.path.join(OUT, f"model_{name}.npz"),
            method=method, d_model=d_model, batch=batch, weight_seed=0, input_seed=100,
            input_digest=digest(emb, mask, text, cand), weight_digest=digest(*sd.values()),
            mask=mask, logits=logits.numpy(), probs=probs.numpy(), query=query.numpy(),
            fitb_dists=dists.numpy(), fitb_argmin=fitb_idx.numpy())
        print(name, "logits", logits.flatten()[:4].tolist())

    # fusion contract: F.normalize per modality then aggregate_embeddings (concat)
    import torch.nn.functional as F
    from src.utils.model_utils import aggregate_embeddings
    img, txt = synth.make_modalities(4, 512, 200)
    fi, ft = F.normalize(torch.from_numpy(img), p=2, dim=-1), F.normalize(torch.from_numpy(txt), p=2, dim=-1)
    cat = aggregate_embeddings(fi, ft, "concat")
    # literal 'mean' only on 1-D inputs (the only shape for which the reference code is right)
    mean1d = torch.stack([aggregate_embeddings(fi[0, j], ft[0, j], "mean") for j in range(16)])
    np.savez_compressed(os.path.join(OUT, "fusion.npz"), seed=200, input_digest=digest(img, txt),
                        concat=cat.numpy(), mean_row0=mean1d.numpy())

    # search idiom: topk(cdist(Q, G), largest=False) over a 3000-item pool (trainer :240-242)
    pool = synth.make_items(3000, 512, seed=300)
    q = synth.make_queries(64, 1024, seed=301) * np.float32(0.05)
    d = torch.cdist(torch.from_numpy(q), torch.from_numpy(pool))
    tk = torch.topk(d, k=50, largest=False)
    np.savez_compressed(os.path.join(OUT, "search_pool3000.npz"), pool_seed=300, query_seed=301,
                        input_digest=digest(pool, q), indices=tk.indices.numpy(),
                        dists=tk.values.numpy())
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
