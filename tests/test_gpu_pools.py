"""GPU: per-category candidate pools + Recall@k against the fp64 oracle and the
reference's own idiom topk(cdist(q, pool), largest=False)
(/root/reference/src/trains/trainers/complementary_item_retrieval_trainer.py:192-249)."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from outfitx_b200 import synth

pytestmark = pytest.mark.gpu


def _setup(n_pools, seed, sizes=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    sizes = sizes or [int(rng.integers(40, 3001)) for _ in range(n_pools)]
    pools = [synth.make_items(n, 512, seed=seed * 100 + c, dup=min(20, n // 4)) for c, n in enumerate(sizes)]
    return pools, sizes, rng


def test_pool_search_matches_fp64_oracle_and_reference_idiom():
    from outfitx_b200.search import PoolSet, pool_search
    pools, sizes, rng = _setup(7, 3, sizes=[3000, 40, 1, 777, 2999, 64, 1500])
    nq = 300
    qp = rng.integers(0, len(pools), size=nq).astype(np.int32)
    q = synth.make_queries(nq, 1024, seed=5) * np.float32(0.05)
    ps = PoolSet.build([torch.from_numpy(p).cuda() for p in pools])
    for k in (1, 10, 50, 64):
        idx, score = pool_search(torch.from_numpy(q).cuda(), torch.from_numpy(qp).cuda(), ps, k=k)
        idx, score = idx.cpu().numpy(), score.cpu().numpy()
        for i in range(nq):
            want_i, want_s = R.search(q[i:i + 1], pools[qp[i]], k=k)
            assert np.array_equal(idx[i], want_i[0])
            np.testing.assert_allclose(score[i], want_s[0], rtol=1e-13, atol=1e-11)
    # the trainer's idiom on one pool (no duplicates among the compared ranks -> identical indices)
    c = 3
    sel = np.nonzero(qp == c)[0]
    d = torch.cdist(torch.from_numpy(q[sel]), torch.from_numpy(pools[c]))
    ref = torch.topk(d, k=10, largest=False).indices.numpy()
    got, _ = pool_search(torch.from_numpy(q[sel]).cuda(), torch.full((len(sel),), c, dtype=torch.int32).cuda(), ps, k=10)
    dsort = np.sort(d.numpy(), -1)
    clear = (np.diff(dsort[:, :11], axis=-1) > 1e-5).all(-1)        # rows whose top-11 distances are distinct in fp32
    assert clear.mean() > 0.3
    assert np.array_equal(got.cpu().numpy()[clear], ref[clear])


def test_recall_at_k_planted_ground_truth():
    from outfitx_b200.search import PoolSet, pool_search, recall_at_k
    pools, sizes, rng = _setup(5, 11)
    nq = 200
    qp = rng.integers(0, len(pools), size=nq).astype(np.int32)
    gt = np.array([rng.integers(0, sizes[c]) for c in qp], dtype=np.int64)
    # queries = ground-truth item + noise: small noise -> rank 0, large noise -> somewhere below
    noise = rng.standard_normal((nq, 1024)).astype(np.float32) * np.where(np.arange(nq) % 2 == 0, 0.01, 0.6)[:, None].astype(np.float32)
    q = np.stack([pools[c][g] for c, g in zip(qp, gt)]) + noise
    ps = PoolSet.build([torch.from_numpy(p).cuda() for p in pools])
    idx, _ = pool_search(torch.from_numpy(q).cuda(), torch.from_numpy(qp).cuda(), ps, k=50)
    got = recall_at_k(idx, torch.from_numpy(gt))
    # the metric computed the reference's way, from the oracle's lists
    want = {}
    lists = np.stack([R.search(q[i:i + 1], pools[qp[i]], k=50)[0][0] for i in range(nq)])
    for k in (1, 5, 10, 15, 30, 50):
        want[f"Recall@{k}"] = float((lists[:, :k] == gt[:, None]).any(-1).mean())
    assert got == pytest.approx(want)
    assert got["Recall@1"] >= 0.45 and got["Recall@50"] >= got["Recall@1"]


def test_pool_search_rejects_bad_arguments():
    from outfitx_b200.search import PoolSet, pool_search
    ps = PoolSet.build([torch.zeros(10, 1024, device="cuda")])
    q = torch.zeros(2, 1024, device="cuda")
    with pytest.raises(ValueError):
        pool_search(q, torch.tensor([0, 1]), ps, k=5)           # pool index out of range
    with pytest.raises(ValueError):
        pool_search(q, torch.tensor([0, 0]), ps, k=65)
    with pytest.raises(ValueError):
        PoolSet.build([torch.zeros(5000, 1024, device="cuda")])
    idx, score = pool_search(q, torch.tensor([0, 0]), ps, k=12)  # pool smaller than k: -1 padding
    assert (idx[:, 10:] == -1).all() and torch.isinf(score[:, 10:]).all()
