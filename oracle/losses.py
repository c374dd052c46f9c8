"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's evaluation-side scoring.

Only ``tests/`` may import this module; the product (``outfitx_b200.losses``) never does.

Pinned against the UNMODIFIED reference classes (``src/losses/focal_loss.py``,
``src/losses/set_wise_ranking_loss.py``) and against ``sklearn.metrics.roc_auc_score`` as
``compute_cp_metrics`` calls it (``compatibility_prediction_trainer.py:406-436``): the outputs they
produced in the build container are committed as ``tests/golden/losses.npz`` by
``oracle/make_golden.py``.
"""
from __future__ import annotations

import numpy as np


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def focal_loss(y_hat, y_true, gamma=2.0, alpha=0.5, reduction="mean", dtype=np.float64):
    """focal_loss.py:23-41.  BCE-with-logits = max(x,0) - x y + log1p(exp(-|x|))."""
    x, y = np.asarray(y_hat, dtype), np.asarray(y_true, dtype)
    ce = np.maximum(x, 0) - x * y + np.log1p(np.exp(-np.abs(x)))
    p = sigmoid(x)
    p_t = p * y + (1 - p) * (1 - y)
    loss = ce * (1 - p_t) ** gamma
    if alpha >= 0:
        loss = (alpha * y + (1 - alpha) * (1 - y)) * loss
    if reduction == "none":
        return loss
    return loss.sum() if reduction == "sum" else loss.mean()


def set_wise_ranking_loss(y, y_hat, negatives, negative_mask, margin=2.0, dtype=np.float64):
    """set_wise_ranking_loss.py:14-37; returns (L_all + L_hard, L_all, L_hard)."""
    y, y_hat, neg = (np.asarray(a, dtype) for a in (y, y_hat, negatives))
    mask = np.asarray(negative_mask, bool)
    pos = np.sqrt(((y_hat - y + dtype(1e-6)) ** 2).sum(-1))            # F.pairwise_distance eps
    nd = np.sqrt(((y_hat[:, None, :] - neg) ** 2).sum(-1))             # (B, K)
    valid = (~mask).astype(dtype)
    hinge = np.maximum(pos[:, None] - nd + margin, 0) * valid
    l_all = hinge.sum() / max(valid.sum(), 1.0)
    hardest = np.where(mask, np.inf, nd).min(axis=1) if nd.shape[1] else np.full(len(pos), np.inf)
    l_hard = np.maximum(pos - hardest + margin, 0).mean()
    return l_all + l_hard, l_all, l_hard


def cp_counts(logits, labels):
    """(TP, FP, FN, correct, n_pos, n_neg, 2 * #(neg < pos) + #(neg == pos)) on fp32 probabilities,
    the integers behind compute_cp_metrics (:406-436)."""
    x = np.asarray(logits, np.float32).reshape(-1)
    p = (np.float32(1) / (np.float32(1) + np.exp(-x))).astype(np.float32)
    y = np.asarray(labels).astype(np.int32).reshape(-1)
    pred = (p > np.float32(0.5)).astype(np.int32)
    tp = int(((pred == 1) & (y == 1)).sum()); fp = int(((pred == 1) & (y == 0)).sum())
    fn = int(((pred == 0) & (y == 1)).sum()); ok = int((pred == y).sum())
    return tp, fp, fn, ok, int((y == 1).sum()), int((y == 0).sum()), cp_counts_from_probs(p, y)


def cp_counts_from_probs(probs, labels):
    """2 * #{pos i, neg j: p_j < p_i} + #{p_j == p_i}: twice the Mann-Whitney U roc_auc_score ranks by."""
    p = np.asarray(probs, np.float32).reshape(-1)
    y = np.asarray(labels).astype(np.int32).reshape(-1)
    pos, neg = np.sort(p[y == 1]), np.sort(p[y == 0])
    less = np.searchsorted(neg, pos, side="left")          # negatives strictly below each positive
    leq = np.searchsorted(neg, pos, side="right")
    return int((2 * less + (leq - less)).sum())


def cp_metrics(logits, labels):
    tp, fp, fn, ok, n_pos, n_neg, auc2 = cp_counts(logits, labels)
    n = n_pos + n_neg
    precision = tp / (tp + fp) if tp + fp else 0.0
    recall = tp / (tp + fn) if tp + fn else 0.0
    f1 = 2 * precision * recall / (precision + recall) if precision + recall else 0.0
    auc = auc2 / (2.0 * n_pos * n_neg) if n_pos and n_neg else 0.0
    return {"Accuracy": ok / n, "Precision": precision, "Recall": recall, "F1": f1, "AUC": auc}
