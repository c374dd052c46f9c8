"""GPU: FocalLoss / SetWiseRankingLoss forward and the CP metric counts through the
C ABI, against the reference's own outputs (tests/golden/losses.npz) and the numpy oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as OL
from oracle.make_golden import loss_inputs

pytestmark = pytest.mark.gpu

FOCAL = ((2, 0.5), (0, 0.25), (1.5, 1.0))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "losses.npz"))


@pytest.mark.parametrize("gamma,alpha", FOCAL)
def test_focal_loss_matches_reference(gold, gamma, alpha):
    from outfitx_b200.losses import FocalLoss
    logits, labels = loss_inputs()[:2]
    x, y = torch.from_numpy(logits).cuda(), torch.from_numpy(labels).cuda()
    for red in ("none", "sum", "mean"):
        got = FocalLoss(gamma=gamma, alpha=alpha, reduction=red)(x, y).cpu().numpy()
        want = gold[f"focal_g{gamma}_a{alpha}_{red}"]
        assert got.shape == want.shape and got.dtype == np.float32
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=5e-7)   # fp32 rounding of the reference on O(1) losses
    # (B,1) logits as the model returns them, large ragged n, determinism
    r = np.random.default_rng(7)
    big = (r.standard_normal((300_001, 1)) * 4).astype(np.float32)
    lab = (r.random((300_001, 1)) < 0.5).astype(np.float32)
    f = FocalLoss(gamma=gamma, alpha=alpha)
    a = f(torch.from_numpy(big).cuda(), torch.from_numpy(lab).cuda())
    b = f(torch.from_numpy(big).cuda(), torch.from_numpy(lab).cuda())
    assert a.item() == b.item()
    np.testing.assert_allclose(a.item(), OL.focal_loss(big, lab, gamma, alpha), rtol=1e-5)


def test_focal_loss_empty_and_shape_errors():
    from outfitx_b200.losses import FocalLoss
    e = torch.zeros(0, device="cuda")
    assert FocalLoss(reduction="sum")(e, e).item() == 0.0
    assert np.isnan(FocalLoss()(e, e).item())                      # torch: mean of nothing
    with pytest.raises(ValueError):
        FocalLoss()(torch.zeros(3, device="cuda"), torch.zeros(4, device="cuda"))


def test_set_wise_ranking_loss_matches_reference(gold):
    from outfitx_b200.losses import SetWiseRankingLoss
    _, _, y, y_hat, neg, mask = loss_inputs()
    c = lambda a: torch.from_numpy(a).cuda()
    for margin in (2.0, 0.1):
        got = SetWiseRankingLoss(margin=margin)(c(y), c(y_hat), c(neg), c(mask))
        assert got.dtype == torch.float32 and got.dim() == 0
        np.testing.assert_allclose(got.item(), float(gold[f"rank_m{margin}"]), rtol=5e-6)
        terms = SetWiseRankingLoss(margin=margin).terms(c(y), c(y_hat), c(neg), c(mask)).cpu().numpy()
        np.testing.assert_allclose(terms, OL.set_wise_ranking_loss(y, y_hat, neg, mask, margin), rtol=5e-6)
    # every negative padded: 0 / clamp(0, 1) + relu(pos - inf + m) = 0
    assert SetWiseRankingLoss()(c(y), c(y_hat), c(neg), c(np.ones_like(mask))).item() == 0.0
    # other shapes: dim 512, 1 and 33 negatives, batch 1
    r = np.random.default_rng(3)
    for b, k, d in ((1, 1, 512), (37, 33, 512), (5, 0, 1024)):
        y2 = r.standard_normal((b, d)).astype(np.float32)
        yh2 = r.standard_normal((b, d)).astype(np.float32)
        n2 = r.standard_normal((b, k, d)).astype(np.float32)
        m2 = r.random((b, k)) < 0.3
        got = SetWiseRankingLoss(1.0)(c(y2), c(yh2), c(n2), c(m2)).item()
        np.testing.assert_allclose(got, OL.set_wise_ranking_loss(y2, yh2, n2, m2, 1.0)[0], rtol=5e-6, atol=1e-7)


def test_cp_metrics_match_reference(gold):
    from outfitx_b200.losses import compute_cp_metrics, cp_counts
    logits, labels = loss_inputs()[:2]
    x, y = torch.from_numpy(logits).cuda(), torch.from_numpy(labels).cuda()
    probs, counts = cp_counts(x.view(-1, 1), y)                    # (B,1) logits as the model returns them
    counts = counts.cpu().tolist()
    np.testing.assert_allclose(probs.cpu().numpy(), gold["cp_probs"], rtol=0, atol=1.2e-7)
    assert counts[:3] == [int(gold["cp_tp"]), int(gold["cp_fp"]), int(gold["cp_fn"])]
    assert counts[4] + counts[5] == len(logits) and counts[4] == int(labels.sum())
    # the AUC pair count is exact on the device's own probabilities
    assert counts[6] == OL.cp_counts_from_probs(probs.cpu().numpy(), labels)
    m = compute_cp_metrics(x, y)
    ref = OL.cp_metrics(logits, labels)
    assert abs(m["AUC"] - float(gold["cp_auc"])) < 1e-6
    for key in ("Accuracy", "Precision", "Recall", "F1"):
        assert abs(m[key] - ref[key]) < 1e-12
    assert abs(m["Accuracy"] - float(gold["cp_accuracy"])) < 1e-7


def test_cp_metrics_sizes_and_degenerate():
    from outfitx_b200.losses import compute_cp_metrics, cp_counts
    r = np.random.default_rng(11)
    for n in (1, 255, 2049, 30_011):
        lg = (r.standard_normal(n) * 2).astype(np.float32)
        lab = (r.random(n) < 0.3).astype(np.float32)
        lg[: n // 3] = np.round(lg[: n // 3])                       # many exact ties
        probs, counts = cp_counts(torch.from_numpy(lg).cuda(), torch.from_numpy(lab).cuda())
        want = OL.cp_counts(lg, lab)
        got = counts.cpu().tolist()
        assert got[:6] == list(want[:6])
        assert got[6] == OL.cp_counts_from_probs(probs.cpu().numpy(), lab)
    one = compute_cp_metrics(torch.ones(10, device="cuda"), torch.ones(10, device="cuda"))
    assert one["AUC"] == 0.0 and one["Accuracy"] == 1.0 and one["Recall"] == 1.0
    _, c0 = cp_counts(torch.zeros(0, device="cuda"), torch.zeros(0, device="cuda"))
    assert c0.cpu().tolist() == [0] * 7
