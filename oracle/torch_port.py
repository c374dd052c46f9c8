"""TEST INFRASTRUCTURE ONLY -- the reference's CPU path rebuilt from stock PyTorch modules.

The reference is pure Python over ``torch.nn`` and cannot travel to the GPU box
(``/root/reference`` is absent there), so the CPU baseline that ``bench.py`` times beside the
CUDA path is this port: the same stock ``nn.TransformerEncoder`` stack the reference
constructs at ``src/models/outfit_x.py:32-45`` (pre-LN, batch-first, mish, d_ffn 2024, 16 heads,
6 layers, dropout inert in eval, ``enable_nested_tensor=False``) with the token assembly of
``:129-136`` / ``:154-163`` and the heads of ``:142-143`` / ``:170-171``, i.e. the identical
library kernels (MKL / oneDNN GEMMs, ATen SDPA / LayerNorm / mish) run on the host cores.
Pinned against the reference itself through ``tests/golden`` (tests/test_oracle.py).
The search baseline is the literal trainer idiom ``topk(cdist(Q, G), largest=False)``
(``complementary_item_retrieval_trainer.py:240-242``), chunked over the gallery because the
full distance matrix cannot be materialised at 1 M+ rows.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


class ReferencePort(nn.Module):
    def __init__(self, d_model: int, d_embed: int = 1024, n_head: int = 16, d_ffn: int = 2024,
                 n_layers: int = 6, dropout: float = 0.3):
        super().__init__()
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=n_head, dim_feedforward=d_ffn,
                                           dropout=dropout, batch_first=True, norm_first=True,
                                           activation=F.mish)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=n_layers,
                                                         enable_nested_tensor=False)
        self.outfit_token = nn.Parameter(torch.zeros(d_model))
        self.cp_ffn = nn.Sequential(nn.Dropout(dropout), nn.Linear(d_model, 1))
        self.cir_ffn = nn.Sequential(nn.Linear(d_model, d_embed, bias=False))
        self.target_item_image_emb = nn.Parameter(torch.zeros(d_model // 2))

    @classmethod
    def from_numpy(cls, sd: dict, **kw):
        d_model = sd["outfit_token"].shape[0]
        m = cls(d_model, d_embed=sd["cir_ffn.0.weight"].shape[0], **kw)
        m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()})
        return m.eval()

    def _encode(self, prefix, emb, mask):
        B = emb.shape[0]
        x = torch.cat([prefix, emb], dim=1)
        pad = torch.cat([torch.zeros(B, 1, dtype=torch.bool), mask], dim=1)
        return self.transformer_encoder(src=x, src_key_padding_mask=pad)[:, 0, :]

    @torch.no_grad()
    def cp(self, emb, mask):
        B = emb.shape[0]
        return self.cp_ffn(self._encode(self.outfit_token.expand(B, 1, -1), emb, mask))

    @torch.no_grad()
    def cir(self, emb, mask, text):
        B = emb.shape[0]
        prefix = torch.cat([self.target_item_image_emb.expand(B, -1), text], dim=-1).unsqueeze(1)
        return self.cir_ffn(self._encode(prefix, emb, mask))


@torch.no_grad()
def fitb(query, cand):
    d = torch.cdist(query.unsqueeze(1), cand, p=2).squeeze(1)
    return torch.argmin(d, dim=-1), d


@torch.no_grad()
def search_cdist_topk(queries, gallery, k=10, chunk=100_000):
    """Trainer idiom, chunked: per chunk topk(cdist), then merge by (dist, idx)."""
    best_d = torch.empty(queries.shape[0], 0)
    best_i = torch.empty(queries.shape[0], 0, dtype=torch.long)
    for lo in range(0, gallery.shape[0], chunk):
        d = torch.cdist(queries, gallery[lo:lo + chunk])
        t = torch.topk(d, k=min(k, d.shape[1]), largest=False)
        best_d = torch.cat([best_d, t.values], 1)
        best_i = torch.cat([best_i, t.indices + lo], 1)
        if best_d.shape[1] > k:
            t = torch.topk(best_d, k=k, largest=False)
            best_d, best_i = t.values, torch.gather(best_i, 1, t.indices)
    return best_i, best_d
