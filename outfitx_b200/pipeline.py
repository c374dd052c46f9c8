"""Host-to-result scoring pipeline: the H2D boundary of the reference's callers, overlapped.

The reference moves every collated batch to the GPU and then runs the model on it
(``v.to(local_rank)`` followed by ``model(**input_dict)``,
``src/trains/trainers/compatibility_prediction_trainer.py:140-145``; the demo does the same,
``src/demo/app.py:124-130``), so copy and compute are serialised.  At 64 KB of fp32 embeddings
per outfit (16 slots x (512 + 512) floats) the copy is the longer of the two on a PCIe-attached
B200, so this module splits a host batch into chunks and double-buffers them: chunk i+1 crosses
PCIe on a copy stream while chunk i is scored on the compute stream, and results return through
pinned host buffers.  Everything it calls is the public ``OutfitX`` API; it adds no arithmetic.

``fetch_valid_only=True`` replaces the DMA copies of the padded per-modality tensors by a kernel that
reads the pinned host tensors in place and fetches only the valid slots (``ofx_fetch_valid_items``,
44 % fewer bytes at n ~ U{2..16}).  Measured on configs[1] it is SLOWER while the scoring kernels run
(16.97 vs 14.95 ms per 8192-outfit step): they are persistent and hold every SM's registers / shared
memory, so the fetch kernel only advances in the gaps, whereas the copy engines need no SM.  It is
therefore off by default and kept for hosts whose batches are mostly padding.  (A shrinking tail of
small chunks, meant to shorten the un-overlapped scoring of the last chunk, was measured too: every
extra chunk costs ~0.8 ms, 16.6 ms per step with chunks 2048 x 3 + 1024 + 512 x 2 -- uniform chunks stay.)
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib


class HostScoringPipeline:
    """Scores host-resident batches with ``model`` (an ``outfitx_b200.OutfitX`` on a CUDA device).

    ``chunk``: outfits per device chunk.  Inputs should be pinned (``tensor.pin_memory()``) for
    the copies to be asynchronous; pageable tensors work but serialise.
    """

    def __init__(self, model, chunk: int = 2048, fetch_valid_only: bool = False):
        self.fetch_valid_only = fetch_valid_only
        if chunk < 1:
            raise ValueError("chunk must be >= 1")
        self.model, self.chunk = model, chunk
        dev = model.device
        if dev.type != "cuda":
            raise RuntimeError("the model must be on a CUDA device: outfitx_b200 has no CPU path")
        self.dev = dev
        self.copy_stream = torch.cuda.Stream(dev)
        self.compute_stream = torch.cuda.Stream(dev)
        self._slots = [dict(), dict()]          # device staging buffers, two in flight
        self._free = [torch.cuda.Event(), torch.cuda.Event()]   # slot may be overwritten
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]  # slot's copies have landed

    def _stage(self, slot: int, name: str, src: torch.Tensor) -> torch.Tensor:
        buf = self._slots[slot].get(name)
        if buf is None or buf.shape[1:] != src.shape[1:] or buf.dtype != src.dtype or buf.shape[0] < src.shape[0]:
            buf = torch.empty((self.chunk,) + tuple(src.shape[1:]), dtype=src.dtype, device=self.dev)
            self._slots[slot][name] = buf
        view = buf[: src.shape[0]]
        view.copy_(src, non_blocking=True)
        return view

    def _stage_items(self, slot: int, img: torch.Tensor, txt: torch.Tensor, mask_dev: torch.Tensor):
        """Device copies of one chunk of the per-modality tensors.  Pinned fp32 sources: the valid slots are
        fetched by a kernel reading host memory in place (padded slots stay stale -- nothing reads them);
        anything else: plain copies of the whole chunk."""
        direct = (self.fetch_valid_only and img.is_pinned() and txt.is_pinned() and img.dtype == torch.float32
                  and txt.dtype == torch.float32 and img.is_contiguous() and txt.is_contiguous() and img.dim() == 3
                  and img.shape == txt.shape and img.shape[-1] % 4 == 0)
        if not direct:
            return self._stage(slot, "img", img), self._stage(slot, "txt", txt)
        out = []
        for name in ("img", "txt"):
            buf = self._slots[slot].get(name)
            if buf is None or buf.shape[1:] != img.shape[1:] or buf.dtype != img.dtype:
                buf = torch.zeros((self.chunk,) + tuple(img.shape[1:]), dtype=img.dtype, device=self.dev)
                self._slots[slot][name] = buf
            out.append(buf[: img.shape[0]])
        n, items, dpm = img.shape
        _lib.check(_lib.lib().ofx_fetch_valid_items(
            img.data_ptr(), txt.data_ptr(), mask_dev.view(torch.uint8).data_ptr(), n, items, dpm,
            out[0].data_ptr(), out[1].data_ptr(), torch.cuda.current_stream(self.dev).cuda_stream))
        return out[0], out[1]

    @torch.no_grad()
    def score(self, image_embeddings: torch.Tensor, text_embeddings: torch.Tensor, outfit_mask: torch.Tensor,
              target_item_text_embedding: Optional[torch.Tensor] = None,
              candidate_item_embedding: Optional[torch.Tensor] = None,
              out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """CP probabilities (and, when ``target_item_text_embedding`` + ``candidate_item_embedding``
        are given, FITB predictions) for a HOST batch of raw per-modality embeddings
        ``(B, L, dim_per_modality)``.  Returns pinned host tensors ``probs (B,)`` [, ``pred (B,)``];
        the call returns once they are valid."""
        B = image_embeddings.shape[0]
        fitb = candidate_item_embedding is not None
        if fitb and target_item_text_embedding is None:
            raise ValueError("FITB scoring needs target_item_text_embedding")
        if out is None:
            out = {"probs": torch.empty(B, dtype=torch.float32).pin_memory()}
            if fitb:
                out["pred"] = torch.empty(B, dtype=torch.int64).pin_memory()
        cur = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_stream(cur)
        self.compute_stream.wait_stream(cur)
        for i, lo in enumerate(range(0, B, self.chunk)):
            hi = min(B, lo + self.chunk)
            s = i & 1
            with torch.cuda.stream(self.copy_stream):
                if i >= 2:
                    self.copy_stream.wait_event(self._free[s])
                d = {"mask": self._stage(s, "mask", outfit_mask[lo:hi])}
                d["img"], d["txt"] = self._stage_items(s, image_embeddings[lo:hi], text_embeddings[lo:hi], d["mask"])
                if fitb:
                    d["text"] = self._stage(s, "text", target_item_text_embedding[lo:hi])
                    d["cand"] = self._stage(s, "cand", candidate_item_embedding[lo:hi])
                self._ready[s].record(self.copy_stream)
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(self._ready[s])
                enc = {"image_embeddings": d["img"], "text_embeddings": d["txt"]}
                probs = self.model.score_cp(outfit_mask=d["mask"], encoder_input_dict=enc)
                out["probs"][lo:hi].copy_(probs, non_blocking=True)
                if fitb:
                    pred, _, _ = self.model.score_fitb(outfit_mask=d["mask"], target_item_text_embedding=d["text"],
                                                       candidate_item_embedding=d["cand"], encoder_input_dict=enc)
                    out["pred"][lo:hi].copy_(pred, non_blocking=True)
                self._free[s].record(self.compute_stream)
        cur.wait_stream(self.compute_stream)
        self.compute_stream.synchronize()
        return out

    @torch.no_grad()
    def score_ids(self, item_ids: torch.Tensor, outfit_mask: torch.Tensor, image_table: torch.Tensor,
                  text_table: torch.Tensor, target_item_text_embedding: Optional[torch.Tensor] = None,
                  candidate_ids: Optional[torch.Tensor] = None, candidate_table: Optional[torch.Tensor] = None,
                  out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """As ``score`` but with the collate on the device: the item embedding tables
        ``(n_items, dim_per_modality)`` [and the fused candidate table ``(n_items, d_embed)``] are
        DEVICE tensors resident in HBM -- the counterpart of the embedding dict the reference's
        trainers keep in host memory (``compatibility_prediction_trainer.py:329-349``) -- and a
        batch is ``item_ids (B, L)`` int32 + ``outfit_mask`` [+ ``candidate_ids (B, n_cand)``] on the
        HOST.  Only ids cross PCIe (64 B per outfit instead of 64 KB)."""
        B = item_ids.shape[0]
        fitb = candidate_ids is not None
        if fitb and (target_item_text_embedding is None or candidate_table is None):
            raise ValueError("FITB scoring needs target_item_text_embedding and candidate_table")
        if out is None:
            out = {"probs": torch.empty(B, dtype=torch.float32).pin_memory()}
            if fitb:
                out["pred"] = torch.empty(B, dtype=torch.int64).pin_memory()
        cur = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_stream(cur)
        self.compute_stream.wait_stream(cur)
        for i, lo in enumerate(range(0, B, self.chunk)):
            hi = min(B, lo + self.chunk)
            s = i & 1
            with torch.cuda.stream(self.copy_stream):
                if i >= 2:
                    self.copy_stream.wait_event(self._free[s])
                d = {"ids": self._stage(s, "ids", item_ids[lo:hi]), "mask": self._stage(s, "idmask", outfit_mask[lo:hi])}
                if fitb:
                    d["text"] = self._stage(s, "text", target_item_text_embedding[lo:hi])
                    d["cids"] = self._stage(s, "cids", candidate_ids[lo:hi])
                self._ready[s].record(self.copy_stream)
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(self._ready[s])
                enc = {"image_embeddings": image_table, "text_embeddings": text_table, "item_ids": d["ids"]}
                probs = self.model.score_cp(outfit_mask=d["mask"], encoder_input_dict=enc)
                out["probs"][lo:hi].copy_(probs, non_blocking=True)
                if fitb:
                    pred, _, _ = self.model.score_fitb(outfit_mask=d["mask"], target_item_text_embedding=d["text"],
                                                       candidate_item_embedding=(candidate_table, d["cids"]),
                                                       encoder_input_dict=enc)
                    out["pred"][lo:hi].copy_(pred, non_blocking=True)
                self._free[s].record(self.compute_stream)
        cur.wait_stream(self.compute_stream)
        self.compute_stream.synchronize()
        return out
