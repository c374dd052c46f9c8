"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's outfit-scoring arithmetic.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``outfitx_b200``) never does.

Parity status: the reference ships NO golden vectors or known-answer tests for this path
("parity unpinned" by its own tests).  This restatement is therefore
pinned against outputs of the reference itself, run in the build container through
``oracle/ref_shim.py`` and committed as ``tests/golden/*.npz`` by ``oracle/make_golden.py``.

The arithmetic lives in a third-party dependency of the reference, PyTorch (pinned
``torch==2.5.0`` in ``/root/reference/environment.yml:9``; this image has 2.11.0), reached
from these reference call sites:

* token assembly  ``src/models/outfit_x.py:129-136`` (CP) and ``:154-163`` (CIR / FITB)
* encoder         ``src/models/outfit_x.py:32-45`` (construction) / ``:137-140``, ``:165-168``
                  -> ``torch.nn.TransformerEncoderLayer`` slow path (pre-LN, mish, float mask)
* heads           ``src/models/outfit_x.py:142-143`` (cp_ffn), ``:170-171`` (cir_ffn)
* fusion          ``src/utils/model_utils.py:26-45`` after the per-modality normalisation of
                  ``src/models/encoders/image/base_image_encoder.py:46-47``
* CP probability  ``src/trains/trainers/compatibility_prediction_trainer.py:408``
* FITB            ``src/trains/trainers/fill_in_the_blank_trainer.py:50-53``
* CIR search      ``src/trains/trainers/complementary_item_retrieval_trainer.py:240-242``

Plain numpy, explicit equations; ``dtype`` selects fp32 or fp64.
"""
from __future__ import annotations

import os

import numpy as np

LN_EPS = 1e-5
N_HEAD = 16


def l2_normalize(x, eps=1e-12):
    """F.normalize(p=2, dim=-1): x / max(||x||, eps)  (base_image_encoder.py:46-47)."""
    n = np.sqrt((x * x).sum(-1, keepdims=True))
    return x / np.maximum(n, eps)


def fuse(img, txt, method="concat", normalize=True):
    """aggregate_embeddings (model_utils.py:26-45).  'mean' implements the intended
    elementwise (img+txt)/2 (the literal torch.mean(dim=-2) is only right
    for 1-D inputs)."""
    if normalize:
        img, txt = l2_normalize(img), l2_normalize(txt)
    if method == "concat":
        return np.concatenate([img, txt], axis=-1)
    if method == "mean":
        return (img + txt) * img.dtype.type(0.5)
    raise ValueError(f"Unsupported aggregation method: {method}. Use 'concat' or 'mean'.")


def layer_norm(x, w, b):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)  # biased variance
    return (x - mu) / np.sqrt(var + x.dtype.type(LN_EPS)) * w + b


def mish(u):
    """F.mish: u * tanh(softplus(u)), softplus threshold 20 as in torch."""
    sp = np.where(u > 20, u, np.log1p(np.exp(np.minimum(u, 20))))
    return u * np.tanh(sp)


def encoder(x, pad, sd, n_layers=6, n_head=N_HEAD):
    """nn.TransformerEncoder(norm_first=True, activation=mish, no final norm), eval mode.

    x: (B,S,Dm); pad: (B,S) bool True = key is padding (additive -inf, every head / row).
    """
    B, S, Dm = x.shape
    hd = Dm // n_head
    dt = x.dtype
    neg = np.where(pad, -np.inf, 0.0).astype(dt)[:, None, None, :]  # (B,1,1,S)
    for l in range(n_layers):
        p = f"transformer_encoder.layers.{l}."
        W = lambda k: sd[p + k].astype(dt)
        h = layer_norm(x, W("norm1.weight"), W("norm1.bias"))
        qkv = h @ W("self_attn.in_proj_weight").T + W("self_attn.in_proj_bias")
        q, k, v = (qkv[..., i * Dm:(i + 1) * Dm].reshape(B, S, n_head, hd).transpose(0, 2, 1, 3)
                   for i in range(3))
        s = (q @ k.transpose(0, 1, 3, 2)) / dt.type(np.sqrt(hd)) + neg
        s = s - s.max(-1, keepdims=True)
        e = np.exp(s)
        a = (e / e.sum(-1, keepdims=True)) @ v  # (B,H,S,hd)
        a = a.transpose(0, 2, 1, 3).reshape(B, S, Dm)
        x = x + a @ W("self_attn.out_proj.weight").T + W("self_attn.out_proj.bias")
        h = layer_norm(x, W("norm2.weight"), W("norm2.bias"))
        u = mish(h @ W("linear1.weight").T + W("linear1.bias"))
        x = x + u @ W("linear2.weight").T + W("linear2.bias")
    return x


def _assemble(prefix, emb, mask):
    B = emb.shape[0]
    x = np.concatenate([np.broadcast_to(prefix, (B, 1, emb.shape[-1])), emb], axis=1)
    pad = np.concatenate([np.zeros((B, 1), bool), np.asarray(mask, bool)], axis=1)
    return x, pad


def cp_forward(sd, emb, mask, dtype=np.float32, n_layers=6):
    """OutfitX._cp_forward (outfit_x.py:120-144) -> logits (B,1)."""
    emb = np.asarray(emb, dtype)
    x, pad = _assemble(sd["outfit_token"].astype(dtype)[None, None, :], emb, mask)
    h0 = encoder(x, pad, sd, n_layers)[:, 0]
    return h0 @ sd["cp_ffn.1.weight"].astype(dtype).T + sd["cp_ffn.1.bias"].astype(dtype)


def cir_forward(sd, emb, mask, text, dtype=np.float32, n_layers=6):
    """OutfitX._cir_forward (outfit_x.py:147-172) -> query embeddings (B, d_embed)."""
    emb = np.asarray(emb, dtype)
    B = emb.shape[0]
    img = np.broadcast_to(sd["target_item_image_emb"].astype(dtype), (B, emb.shape[-1] // 2))
    prefix = np.concatenate([img, np.asarray(text, dtype)], axis=-1)[:, None, :]
    x, pad = _assemble(prefix, emb, mask)
    h0 = encoder(x, pad, sd, n_layers)[:, 0]
    return h0 @ sd["cir_ffn.0.weight"].astype(dtype).T


def sigmoid(z):
    """compatibility_prediction_trainer.py:408 / demo/app.py:130."""
    return 1.0 / (1.0 + np.exp(-z))


def fitb(query, cand):
    """cdist(q[:,None], cand, p=2).squeeze(1) -> argmin (fill_in_the_blank_trainer.py:50-53).

    Returns (argmin (B,) int64 -- first minimum on ties, dists (B,4))."""
    d = np.sqrt(((cand - query[:, None, :]) ** 2).sum(-1))
    return d.argmin(-1).astype(np.int64), d


_LIB = None


def _search_lib():
    """oracle/search_oracle.c through ctypes (built on first use by oracle/Makefile)."""
    global _LIB
    if _LIB is None:
        import ctypes
        import os
        import subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        so = os.path.join(here, "_build", "liboracle_search.so")
        src = os.path.join(here, "search_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", here])
        lib = ctypes.CDLL(so)
        i64, vp = ctypes.c_int64, ctypes.c_void_p
        lib.ofx_oracle_scores.argtypes = [vp, i64, vp, i64, i64, ctypes.c_int, vp]
        lib.ofx_oracle_search.argtypes = [vp, i64, vp, i64, i64, ctypes.c_int, i64, i64, ctypes.c_int, vp, vp]
        lib.ofx_oracle_scores.restype = lib.ofx_oracle_search.restype = None
        _LIB = lib
    return _LIB


_METRIC = {"dot": 0, "l2": 1}


def search_scores(queries, gallery, metric="l2"):
    """fp64 ranking score, larger = better.  'l2': q.g - 0.5|g|^2 (same order as -cdist,
    SURVEY.md D8); 'dot': q.g.  Position-independent sequential fp64 sums (search_oracle.c)."""
    q = np.ascontiguousarray(queries, np.float32)
    g = np.ascontiguousarray(gallery, np.float32)
    out = np.empty((len(q), len(g)), np.float64)
    _search_lib().ofx_oracle_scores(q.ctypes.data, len(q), g.ctypes.data, len(g), q.shape[1],
                                    _METRIC[metric], out.ctypes.data)
    return out


def topk_lex(scores, k, id_offset=0):
    """Exact top-k by the lexicographic key (-score, index): ties -> lowest index first
    (torch.topk does not guarantee this)."""
    order = np.argsort(-scores, axis=-1, kind="stable")[:, :k]
    return order.astype(np.int64) + id_offset, np.take_along_axis(scores, order, -1)


def search(queries, gallery, k=10, metric="l2", id_offset=0):
    """Exact k-NN of complementary_item_retrieval_trainer.py:240-242 restated as an fp64
    max-score search with deterministic tie-break.  Returns (idx (nq,k) i64, score f64)."""
    q = np.ascontiguousarray(queries, np.float32)
    g = np.ascontiguousarray(gallery, np.float32)
    idx = np.empty((len(q), k), np.int64)
    score = np.empty((len(q), k), np.float64)
    _search_lib().ofx_oracle_search(q.ctypes.data, len(q), g.ctypes.data, len(g), q.shape[1],
                                    _METRIC[metric], k, id_offset, os.cpu_count() or 1,
                                    idx.ctypes.data,
                                    score.ctypes.data)
    return idx, score


def merge_topk(idx, score, k):
    """Merge candidate lists by (-score, idx) -- the all-gather merge of the sharded search."""
    order = np.lexsort((idx, -score), axis=-1)[:, :k]
    return np.take_along_axis(idx, order, -1), np.take_along_axis(score, order, -1)
