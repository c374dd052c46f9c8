"""Evaluation-side scoring of the reference's trainers, forward only, on the GPU.

Same names, constructor arguments and ``forward`` signatures as the reference modules, so a
validation loop can swap them in:

* ``FocalLoss``            ``src/losses/focal_loss.py:8-41``
* ``SetWiseRankingLoss``   ``src/losses/set_wise_ranking_loss.py:5-37``
* ``compute_cp_metrics``   ``src/trains/trainers/compatibility_prediction_trainer.py:406-436``
* ``gather_cp_eval``       the DDP metric all-gather of ``compatibility_prediction_trainer.py:385-399``

Forward only: the returned scalars carry no autograd graph (the backward pass is a separate
project).  No CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib


def _need_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("outfitx_b200.losses has no CPU path: tensors must live on a CUDA device")


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _ws(batch: int, dev) -> torch.Tensor:
    return torch.empty(_lib.lib().ofx_loss_workspace_bytes(batch), dtype=torch.uint8, device=dev)


class FocalLoss(torch.nn.Module):
    """``FocalLoss(gamma=2, alpha=0.5, reduction='mean')`` -- focal_loss.py:8-41."""

    def __init__(self, gamma=2, alpha=0.5, reduction="mean"):
        super().__init__()
        # same argument checks (and exception type) as the reference constructor, focal_loss.py:9-18
        assert gamma >= 0, f"gamma must be >= 0, got {gamma}"
        assert 0 <= alpha <= 1, f"alpha must lie in [0, 1], got {alpha}"
        assert reduction in ("none", "mean", "sum"), f"reduction must be one of none / mean / sum, got {reduction!r}"
        self.gamma, self.alpha, self.reduction = gamma, alpha, reduction

    def forward(self, y_hat: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
        _need_cuda(y_hat, y_true)
        if y_hat.shape != y_true.shape:
            raise ValueError(f"Target size ({tuple(y_true.shape)}) must be the same as input size ({tuple(y_hat.shape)})")
        x, y = _f32(y_hat), _f32(y_true)
        n = x.numel()
        dev = x.device
        per = torch.empty_like(x) if self.reduction == "none" else None
        out = torch.empty(2, dtype=torch.float64, device=dev)
        ws = _ws(0, dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ofx_focal_loss(
                x.data_ptr(), y.data_ptr(), n, float(self.gamma), float(self.alpha),
                per.data_ptr() if per is not None else None, out.data_ptr(), ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream(dev).cuda_stream))
        if self.reduction == "none":
            return per
        return out[0 if self.reduction == "sum" else 1].to(torch.float32)


class SetWiseRankingLoss(torch.nn.Module):
    """``SetWiseRankingLoss(margin=2.0)`` -- set_wise_ranking_loss.py:5-37."""

    def __init__(self, margin: float = 2.0):
        super().__init__()
        self.margin = margin

    def forward(self, batch_y: torch.Tensor, batch_y_hat: torch.Tensor, batch_negative_samples: torch.Tensor,
                batch_negative_mask: torch.Tensor) -> torch.Tensor:
        return self.terms(batch_y, batch_y_hat, batch_negative_samples, batch_negative_mask)[0].to(torch.float32)

    def terms(self, batch_y, batch_y_hat, batch_negative_samples, batch_negative_mask) -> torch.Tensor:
        """(L_all + L_hard, L_all, L_hard) as one fp64 device tensor."""
        _need_cuda(batch_y, batch_y_hat, batch_negative_samples, batch_negative_mask)
        y, yh, neg = _f32(batch_y), _f32(batch_y_hat), _f32(batch_negative_samples)
        if yh.dim() != 2 or y.shape != yh.shape or neg.dim() != 3 or neg.shape[0] != yh.shape[0] or neg.shape[2] != yh.shape[1]:
            raise ValueError(f"shapes: y {tuple(y.shape)}, y_hat {tuple(yh.shape)}, negatives {tuple(neg.shape)}")
        if tuple(batch_negative_mask.shape) != tuple(neg.shape[:2]):
            raise ValueError(f"negative mask {tuple(batch_negative_mask.shape)} != {tuple(neg.shape[:2])}")
        mask = batch_negative_mask.detach().to(torch.bool).contiguous().view(torch.uint8)
        b, k, d = neg.shape
        dev = yh.device
        out = torch.empty(3, dtype=torch.float64, device=dev)
        ws = _ws(b, dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ofx_set_wise_ranking_loss(
                y.data_ptr(), yh.data_ptr(), neg.data_ptr(), mask.data_ptr(), b, k, d, float(self.margin),
                out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        return out


def cp_counts(y_hats: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(probs (n) fp32, counts (7) int64) on the device -- see ``ofx_cp_metrics`` in ofx.h."""
    _need_cuda(y_hats, labels)
    x, y = _f32(y_hats).reshape(-1), _f32(labels).reshape(-1)
    if x.numel() != y.numel():
        raise ValueError(f"{x.numel()} scores but {y.numel()} labels")
    dev = x.device
    probs = torch.empty_like(x)
    counts = torch.empty(7, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ofx_cp_metrics(x.data_ptr(), y.data_ptr(), x.numel(), probs.data_ptr(),
                                             counts.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return probs, counts


def metrics_from_counts(counts) -> Dict[str, float]:
    """The dictionary of compute_cp_metrics (:430-436) from the seven integer counts."""
    tp, fp, fn, ok, n_pos, n_neg, auc2 = (int(c) for c in counts)
    n = n_pos + n_neg
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
    f1 = (2 * precision * recall) / (precision + recall) if (precision + recall) > 0 else 0.0
    # roc_auc_score is only called when both classes are present (:412), else 0.0
    auc = auc2 / (2.0 * n_pos * n_neg) if n_pos > 0 and n_neg > 0 else 0.0
    return {"Accuracy": ok / n if n else float("nan"), "Precision": precision, "Recall": recall, "F1": f1, "AUC": auc}


def compute_cp_metrics(y_hats: torch.Tensor, labels: torch.Tensor) -> Dict[str, float]:
    """compatibility_prediction_trainer.py:406-436 with the counting done on the GPU (one 56-byte
    device->host read instead of copying every score)."""
    _, counts = cp_counts(y_hats, labels)
    return metrics_from_counts(counts.cpu().tolist())


def gather_cp_eval(local_y_hats: torch.Tensor, local_labels: torch.Tensor, local_loss: torch.Tensor,
                   batch_count: int = 1, group=None) -> Tuple[torch.Tensor, torch.Tensor, float]:
    """The metric all-gather of compatibility_prediction_trainer.py:385-399: every rank ends up with the
    rank-ordered concatenation of all scores and labels and with mean(loss over ranks) / batch_count.
    Pure ``torch.distributed`` plumbing (NCCL on the GPU box, gloo in the CPU tests); like the reference
    it requires equally sized local tensors."""
    import torch.distributed as dist

    y, lab, loss = local_y_hats.detach(), local_labels.detach(), local_loss.detach()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        w = dist.get_world_size(group)
        ys = [torch.empty_like(y) for _ in range(w)]
        ls = [torch.empty_like(lab) for _ in range(w)]
        los = [torch.empty_like(loss) for _ in range(w)]
        dist.all_gather(ys, y, group=group)
        dist.all_gather(ls, lab, group=group)
        dist.all_gather(los, loss, group=group)
    else:
        ys, ls, los = [y], [lab], [loss]
    return torch.cat(ys, dim=0), torch.cat(ls, dim=0), (torch.stack(los).mean() / batch_count).item()


def cp_eval_metrics(local_y_hats, local_labels, local_loss, batch_count: int = 1, group=None) -> Dict[str, float]:
    """build_metrics (:371-404): all-gather, then {'loss', Accuracy, Precision, Recall, F1, AUC}."""
    all_y, all_lab, loss = gather_cp_eval(local_y_hats, local_labels, local_loss, batch_count, group)
    return {"loss": loss, **compute_cp_metrics(all_y, all_lab)}
