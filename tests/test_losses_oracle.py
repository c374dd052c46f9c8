"""CPU: pin the N4 oracle (oracle/losses.py) to what the reference's own FocalLoss / SetWiseRankingLoss
classes and compute_cp_metrics' statements (with sklearn's roc_auc_score) produced
(tests/golden/losses.npz, written by oracle/make_golden.py), plus the host side of the metric
all-gather (world_size 2 over gloo) and the no-CPU-path rule."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import losses as OL
from oracle.make_golden import digest, loss_inputs

FOCAL = ((2, 0.5), (0, 0.25), (1.5, 1.0))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "losses.npz"))


def test_inputs_regenerate(gold):
    assert digest(*loss_inputs()) == str(gold["input_digest"])


@pytest.mark.parametrize("gamma,alpha", FOCAL)
def test_focal_restatement_matches_reference(gold, gamma, alpha):
    logits, labels = loss_inputs()[:2]
    for red in ("none", "sum", "mean"):
        want = gold[f"focal_g{gamma}_a{alpha}_{red}"]
        got = OL.focal_loss(logits, labels, gamma, alpha, red)
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=5e-7)   # fp32 rounding of the reference on O(1) losses


def test_ranking_restatement_matches_reference(gold):
    _, _, y, y_hat, neg, mask = loss_inputs()
    for margin in (2.0, 0.1):
        got = OL.set_wise_ranking_loss(y, y_hat, neg, mask, margin)[0]
        np.testing.assert_allclose(got, gold[f"rank_m{margin}"], rtol=2e-6)
    # all negatives padded everywhere: L_all = 0 / clamp(0, 1) and L_hard = relu(-inf) = 0
    assert OL.set_wise_ranking_loss(y, y_hat, neg, np.ones_like(mask))[0] == 0.0


def test_cp_metrics_restatement_matches_reference(gold):
    logits, labels = loss_inputs()[:2]
    tp, fp, fn, ok, n_pos, n_neg, auc2 = OL.cp_counts(logits, labels)
    assert (tp, fp, fn) == (int(gold["cp_tp"]), int(gold["cp_fp"]), int(gold["cp_fn"]))
    assert abs(ok / len(logits) - float(gold["cp_accuracy"])) < 1e-7
    m = OL.cp_metrics(logits, labels)
    assert abs(m["AUC"] - float(gold["cp_auc"])) < 1e-12          # Mann-Whitney with ties at 1/2 == roc_auc_score
    # brute force on a prefix: the pair-count definition itself
    p = gold["cp_probs"][:300]
    y = labels[:300].astype(int)
    brute = sum(2 * (pj < pi) + (pj == pi) for pi in p[y == 1] for pj in p[y == 0])
    assert OL.cp_counts(logits[:300], labels[:300])[6] == brute


def test_metrics_from_counts_degenerate_cases():
    from outfitx_b200.losses import metrics_from_counts
    m = metrics_from_counts([0, 0, 0, 5, 0, 5, 0])     # one class only -> AUC 0.0 (:412), no division by zero
    assert m == {"Accuracy": 1.0, "Precision": 0.0, "Recall": 0.0, "F1": 0.0, "AUC": 0.0}
    m = metrics_from_counts([3, 1, 2, 8, 5, 5, 40])
    assert m["Precision"] == 0.75 and m["Recall"] == 0.6 and m["AUC"] == 0.8 and m["Accuracy"] == 0.8


def test_losses_have_no_cpu_path():
    from outfitx_b200.losses import FocalLoss, SetWiseRankingLoss, compute_cp_metrics
    x = torch.zeros(4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        FocalLoss()(x, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        SetWiseRankingLoss()(torch.zeros(2, 8), torch.zeros(2, 8), torch.zeros(2, 3, 8), torch.zeros(2, 3, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="no CPU path"):
        compute_cp_metrics(x, x)
    with pytest.raises(AssertionError):
        FocalLoss(gamma=-1)
    with pytest.raises(AssertionError):
        FocalLoss(reduction="avg")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from outfitx_b200.losses import gather_cp_eval
        y = torch.arange(5, dtype=torch.float32) + 10 * rank
        lab = torch.full((5,), float(rank))
        loss = torch.tensor(3.0 * (rank + 1))
        ay, al, l = gather_cp_eval(y, lab, loss, batch_count=3)
        if rank == 1:
            np.savez(out, y=ay.numpy(), lab=al.numpy(), loss=l)
    finally:
        dist.destroy_process_group()


def test_metric_all_gather_host_logic(tmp_path):
    """compatibility_prediction_trainer.py:385-399: rank-ordered concatenation, loss = mean over ranks / batch_count."""
    out = str(tmp_path / "r1.npz")
    mp.spawn(_gather_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    g = np.load(out)
    assert np.array_equal(g["y"], np.array([0, 1, 2, 3, 4, 10, 11, 12, 13, 14], np.float32))
    assert np.array_equal(g["lab"], np.array([0] * 5 + [1] * 5, np.float32))
    assert abs(float(g["loss"]) - (3.0 + 6.0) / 2 / 3) < 1e-7
    # single process (no process group): identity
    from outfitx_b200.losses import gather_cp_eval
    ay, al, l = gather_cp_eval(torch.ones(3), torch.zeros(3), torch.tensor(4.0), batch_count=2)
    assert ay.tolist() == [1, 1, 1] and l == 2.0
