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

# (name, fusion, d_model, batch): lengths are drawn by case_inputs below.  slip1536 is the reference's DEFAULT
# configuration (ItemEncoderConfig.type = 'slip': 768 per modality, d_model 1536, head_dim 96,
# src/models/configs/item_encoder_config.py:9,24-26).
CASES = [("concat1024", "concat", 1024, 8), ("mean512", "mean", 512, 8), ("slip1536", "concat", 1536, 8)]


def case_dims(method: str, d_model: int):
    """-> (dim_per_modality, cfg.d_embed = 2 * dim_per_modality, encoder type)."""
    dpm = d_model // 2 if method == "concat" else d_model
    return dpm, 2 * dpm, ("slip" if dpm == 768 else "clip")


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def case_inputs(method: str, batch: int, seed: int = 100, dpm: int = 512):
    """Inputs of a golden case: ragged lengths incl. the edge cases 0, 1 and 16 items, plus
    one outfit whose valid items are NOT left-aligned (the model must not care)."""
    img, txt = synth.make_modalities(batch, dpm, seed)
    lengths = synth.make_lengths(batch, seed + 1)
    lengths[:3] = (0, 1, 16)
    mask = synth.make_mask(lengths)
    mask[3] = np.array([1, 0, 1, 1, 0, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0], bool)
    emb = synth.fuse(img, txt, method)
    emb[mask] = 0.0
    d_model = emb.shape[-1]
    text = synth.make_text_prefix(batch, d_model // 2, seed + 2)
    cand = synth.make_items(batch * 4, dpm, seed + 3).reshape(batch, 4, 2 * dpm)
    return img, txt, emb, mask, text, cand


def processor_items(seed: int = 500):
    """Raw material of the collate fixture: per outfit a list of item embeddings (ragged: 1, 3, 16, 20 -> truncated
    to 16, 7, 2 items), the target item's embedding and 4 FITB candidates.  1024-d fused items (clip + concat)."""
    lengths = [1, 3, 16, 20, 7, 2]
    pool = synth.make_items(sum(lengths) + len(lengths) * 5, 512, seed)
    outfits, targets, cands, at = [], [], [], 0
    for n in lengths:
        outfits.append([pool[at + i] for i in range(n)])
        at += n
        targets.append(pool[at])
        cands.append(np.stack([pool[at + 1 + j] for j in range(4)]))
        at += 5
    return outfits, targets, cands


def write_processor_golden(out_dir: str = OUT):
    """The reference's OWN collate (outfit_x_base_processor.py:20-81, CP processor :7-22, CIR processor :95-114,
    FITB processor :9-40) run on lists of its own pydantic task objects; the tensors it emits and what the
    reference model answers to them."""
    ox, cfgs, dts = ref_shim.load_reference()
    from src.models.processor.outfit_x.outfit_x_compatibility_prediction_task_processor import (
        OutfitXCompatibilityPredictionTaskProcessor)
    from src.models.processor.outfit_x.outfit_x_complementary_item_retrieval_processor import (
        OutfitXComplementaryItemRetrievalTaskProcessor)
    from src.models.processor.outfit_x.outfit_x_fill_in_the_blank_task_processor import (
        OutfitXFillInTheBlankTaskProcessor)
    cfg = cfgs.OutfitXConfig(item_encoder=cfgs.ItemEncoderConfig(type="clip", aggregation_method="concat"))
    outfits, targets, cands = processor_items()
    item = lambda e, i: dts.FashionItem(item_id=i, embedding=e, text_embedding=e[512:])   # text = second half
    # (polyvore_item_dataset.py:75)
    cp_batch = [(dts.OutfitCompatibilityPredictionTask(outfit=[item(e, i) for i, e in enumerate(o)]), float(b % 2))
                for b, o in enumerate(outfits)]
    cir_batch = [(dts.OutfitComplementaryItemRetrievalTask(outfit=[item(e, i) for i, e in enumerate(o)],
                                                           target_item=item(t, 1000 + b)), None)
                 for b, (o, t) in enumerate(zip(outfits, targets))]
    fitb_batch = [(dts.OutfitFillInTheBlankTask(outfit=[item(e, i) for i, e in enumerate(o)], target_item=item(t, 1000 + b)),
                   torch.from_numpy(c), b % 4)
                  for b, (o, t, c) in enumerate(zip(outfits, targets, cands))]
    cp = OutfitXCompatibilityPredictionTaskProcessor(cfg)(cp_batch)
    cir = OutfitXComplementaryItemRetrievalTaskProcessor("test", cfg)(cir_batch)
    fitb = OutfitXFillInTheBlankTaskProcessor(cfg)(fitb_batch)
    sd = synth.make_state_dict(1024, 1024, seed=0)
    model = ref_shim.build_reference_model("concat", sd)
    logits = model(**cp["input_dict"])
    query = model(**cir["input_dict"])
    q_fitb = model(**fitb["input_dict"])
    d = torch.cdist(q_fitb.unsqueeze(1), fitb["candidate_item_embedding"], p=2).squeeze(1)   # fitb trainer :52-53
    np.savez_compressed(
        os.path.join(out_dir, "processor_clip1024.npz"), seed=500, weight_seed=0,
        cp_outfit_embedding=cp["input_dict"]["outfit_embedding"].numpy(), cp_outfit_mask=cp["input_dict"]["outfit_mask"].numpy(),
        cp_task=cp["input_dict"]["task"].__name__, cp_label=cp["label"].numpy(),
        cir_outfit_embedding=cir["input_dict"]["outfit_embedding"].numpy(), cir_outfit_mask=cir["input_dict"]["outfit_mask"].numpy(),
        cir_text=cir["input_dict"]["target_item_text_embedding"].numpy(), cir_task=cir["input_dict"]["task"].__name__,
        cir_pos_item_id=np.asarray(cir["pos_item_id"]),
        fitb_task=fitb["input_dict"]["task"].__name__, fitb_cand=fitb["candidate_item_embedding"].numpy(),
        fitb_answer=fitb["answer_index"].numpy(),
        logits=logits.numpy(), query=query.numpy(), fitb_dists=d.numpy(), fitb_argmin=torch.argmin(d, -1).numpy())
    print("processor_clip1024.npz: logits", logits.flatten()[:3].tolist(), "mask rows", cp["input_dict"]["outfit_mask"].sum(-1).tolist())


def loss_inputs(seed: int = 400):
    """Seeded inputs of the N4 fixtures (regenerated by the tests, digest stored)."""
    r = np.random.default_rng(seed)
    n = 4096
    logits = (r.standard_normal(n) * 3).astype(np.float32)
    logits[:6] = (0.0, 30.0, -30.0, 1e-8, -1e-8, 88.0)            # saturation / threshold edge cases
    labels = (r.random(n) < 0.4).astype(np.float32)
    logits[100:110] = logits[110:120]                              # exact score ties across classes
    labels[100:110], labels[110:120] = 1.0, 0.0
    b, k, d = 48, 10, 1024
    y = r.standard_normal((b, d)).astype(np.float32) * np.float32(0.05)
    y_hat = (y + r.standard_normal((b, d)).astype(np.float32) * np.float32(0.03)).astype(np.float32)
    neg = (y[:, None, :] + r.standard_normal((b, k, d)).astype(np.float32) * np.float32(0.04)).astype(np.float32)
    n_valid = r.integers(0, k + 1, size=b)
    n_valid[:3] = (0, 1, k)                                        # an outfit with no negatives at all
    neg_mask = np.arange(k)[None, :] >= n_valid[:, None]           # True = padding
    return logits, labels, y, y_hat, neg, neg_mask


def write_loss_golden(out_dir: str = OUT):
    """FocalLoss / SetWiseRankingLoss from the reference's own classes, CP metrics as
    compute_cp_metrics computes them (compatibility_prediction_trainer.py:406-436)."""
    if ref_shim.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    import importlib.util

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref_shim.REFERENCE_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    fl = load("_ref_focal_loss", "src/losses/focal_loss.py")
    rl = load("_ref_set_wise_ranking_loss", "src/losses/set_wise_ranking_loss.py")
    from sklearn.metrics import roc_auc_score
    logits, labels, y, y_hat, neg, neg_mask = loss_inputs()
    t = torch.from_numpy
    out = {"seed": 400, "input_digest": digest(logits, labels, y, y_hat, neg, neg_mask)}
    for gamma, alpha in ((2, 0.5), (0, 0.25), (1.5, 1.0)):
        for red in ("none", "sum", "mean"):
            v = fl.FocalLoss(gamma=gamma, alpha=alpha, reduction=red)(t(logits), t(labels))
            out[f"focal_g{gamma}_a{alpha}_{red}"] = v.numpy()
    for margin in (2.0, 0.1):
        out[f"rank_m{margin}"] = rl.SetWiseRankingLoss(margin=margin)(t(y), t(y_hat), t(neg), t(neg_mask)).numpy()
    # compute_cp_metrics, statement by statement (:406-436)
    probs = torch.sigmoid(t(logits).float()).detach().cpu()
    lab = t(labels).int().detach().cpu()
    auc = roc_auc_score(lab.numpy(), probs.numpy())
    pred = (probs > 0.5).int()
    tp = torch.sum((pred == 1) & (lab == 1)).item()
    fp = torch.sum((pred == 1) & (lab == 0)).item()
    fn = torch.sum((pred == 0) & (lab == 1)).item()
    acc = torch.mean((pred == lab).float()).item()
    out.update(cp_tp=tp, cp_fp=fp, cp_fn=fn, cp_accuracy=acc, cp_auc=auc, cp_probs=probs.numpy())
    np.savez_compressed(os.path.join(out_dir, "losses.npz"), **out)
    print("losses.npz: focal mean", float(out["focal_g2_a0.5_mean"]), "rank", float(out["rank_m2.0"]), "auc", auc)


def main(only=None):
    """only: names of the fixtures to (re)write (model cases by name, 'processor', 'fusion', 'search', 'losses');
    None = all.  Existing fixtures are byte-stable under regeneration on the same torch build."""
    assert ref_shim.available(), "reference not mounted"
    os.makedirs(OUT, exist_ok=True)
    ox, cfgs, dts = ref_shim.load_reference()
    torch.set_grad_enabled(False)
    want = lambda n: only is None or n in only

    for name, method, d_model, batch in CASES:
        if not want(name):
            continue
        dpm, d_embed, enc_type = case_dims(method, d_model)
        sd = synth.make_state_dict(d_model, d_embed, seed=0)
        model = ref_shim.build_reference_model(method, sd, encoder_type=enc_type)
        assert model.item_encoder.d_embed == d_model and model.cfg.d_embed == d_embed
        img, txt, emb, mask, text, cand = case_inputs(method, batch, dpm=dpm)
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

    if want("processor"):
        write_processor_golden()
    if want("losses"):
        write_loss_golden()
    if not want("fusion") and not want("search"):
        return
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
    main(sys.argv[1:] or None)
