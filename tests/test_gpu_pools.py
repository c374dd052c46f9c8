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
    # rows whose top-11 distances are clearly distinct in fp32 (the host's cdist rounds differently from CPU to CPU:
    # a 1e-5 margin held on most boxes of the pool but not on all)
    clear = (np.diff(dsort[:, :11], axis=-1) > 1e-4).all(-1)
    assert clear.mean() > 0.05
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


def test_embedding_pickles_flow_into_gallery_and_search(tmp_path):
    """SURVEY.md N3 end to end: files in the reference's precomputed-embedding wire format
    (precompute_embedding_script.py:47-53: pickle {'ids': [int], 'embeddings': float32 (N, 2*dpm)}, one file per
    rank, named {model_name}_embedding_subset_{rank}.pkl) -> load_embedding_pickles -> Gallery.build -> CIR query
    embedding (text prefix = SECOND half of the target item's row, polyvore_item_dataset.py:75) -> cir_search.
    The retrieved ITEM IDS must be those of the fp64 oracle run on the same arrays."""
    import pickle
    import outfitx_b200 as o
    from oracle import torch_port
    from outfitx_b200.search import Gallery, cir_search, load_embedding_pickles
    n = 9000
    items = synth.make_items(n, 512, seed=71, dup=40)
    rng = np.random.Generator(np.random.PCG64(72))
    item_ids = rng.permutation(np.arange(10_000, 10_000 + 3 * n, 3))[:n].astype(np.int64)      # non-contiguous catalogue ids
    cuts, paths = (0, 4000, 6500, n), []
    for r in range(3):
        d = {"ids": [int(i) for i in item_ids[cuts[r]:cuts[r + 1]]], "embeddings": items[cuts[r]:cuts[r + 1]]}
        p = tmp_path / f"fashion-clip_embedding_subset_{r}.pkl"
        with open(p, "wb") as f:
            pickle.dump(d, f)
        paths.append(str(p))
    ids, emb, index = load_embedding_pickles(paths, device="cuda")
    assert emb.is_cuda and emb.shape == (n, 1024) and torch.equal(ids, torch.from_numpy(item_ids))
    gallery = Gallery.build(emb)
    # outfits made of catalogue items, looked up by item id as the datasets do; the target item's text half is the prefix
    B, dev = 48, "cuda"
    lengths = synth.make_lengths(B, 73)
    rows = rng.integers(0, n, size=(B, 16))
    mask = synth.make_mask(lengths)
    outfit = emb[torch.from_numpy(rows).to(dev)]
    outfit[torch.from_numpy(mask).to(dev)] = 0.0
    targets = rng.integers(0, n, size=B)
    text = emb[torch.from_numpy(targets).to(dev)][:, 512:].contiguous()
    sd = synth.make_state_dict(1024, 1024, seed=0)
    m = o.OutfitX(o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type="clip")), precision="fp32")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(dev)
    query = m(o.OutfitComplementaryItemRetrievalTask, outfit_embedding=outfit, outfit_mask=torch.from_numpy(mask).to(dev),
              target_item_text_embedding=text)
    # the query embeddings agree with the reference stack ...
    port = torch_port.ReferencePort.from_numpy(sd)
    want_q = port.cir(outfit.cpu(), torch.from_numpy(mask), text.cpu()).numpy()
    assert float(np.abs(query.cpu().numpy() - want_q).max() / np.abs(want_q).max()) <= 1e-3
    # ... and searching with them returns the oracle's rows, reported as catalogue item ids
    idx, score = cir_search(query, gallery, k=10, metric="l2")
    want_i, want_s = R.search(query.cpu().numpy(), items, k=10)
    assert np.array_equal(idx.cpu().numpy(), want_i)
    got_ids = ids.to(dev)[idx]
    assert np.array_equal(got_ids.cpu().numpy(), item_ids[want_i])
    assert all(index[int(i)] == int(r) for i, r in zip(got_ids[:, 0].cpu(), idx[:, 0].cpu()))
