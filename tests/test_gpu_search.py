"""GPU parity of the CIR search (ofx_topk_search / ofx_topk_merge) against the fp64 oracle
(oracle/search_oracle.c, itself pinned to the reference idiom topk(cdist) in tests/golden).

Bar (BASELINE.json north_star): exact search -> indices bit-exact vs the fp64 (-score, idx)
oracle; bf16 pass alone -> recall@10 >= 0.999.
"""
import os

import numpy as np
import pytest
import torch

from oracle import restatement as R
from outfitx_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def _search(q, gal, k, metric="l2", exact=True, id_offset=0):
    from outfitx_b200.search import Gallery, local_search
    G = Gallery.build(_t(gal), id_offset=id_offset)
    idx, score = local_search(_t(q), G, k, metric, exact)
    torch.cuda.synchronize()
    return idx.cpu().numpy(), score.cpu().numpy()


def test_reference_pool_golden(golden_dir):
    """The trainer's own setting (complementary_item_retrieval_trainer.py:240-242): a 3000-item category
    pool ranked with k = max(top_k_list) = 50."""
    g = np.load(os.path.join(golden_dir, "search_pool3000.npz"))
    pool = synth.make_items(3000, 512, seed=int(g["pool_seed"]))
    q = synth.make_queries(64, 1024, seed=int(g["query_seed"])) * np.float32(0.05)
    for k in (50, 64):
        idx, score = _search(q, pool, k)
        want_i, want_s = R.search(q, pool, k=k)
        assert np.array_equal(idx, want_i)
        np.testing.assert_allclose(score, want_s, rtol=0, atol=1e-12)
    # top-10 equal to the reference's fp32 topk(cdist) output itself
    assert np.array_equal(idx[:, :10], g["indices"][:, :10])


@pytest.mark.parametrize("n,nq,k,metric", [
    (20000, 300, 10, "l2"), (20000, 300, 10, "dot"), (5000, 7, 1, "l2"), (100, 130, 16, "dot"),
    (257, 5, 10, "l2"), (70001, 129, 32, "l2"), (40_000, 33, 50, "l2"), (9000, 20, 64, "dot"), (30_000, 65, 20, "l2"),
    # large enough for the block-maxima threshold seeding (>= 16 x 8192 rows at k <= 16, 16 x 16384 above)
    (150_001, 64, 10, "l2"), (270_000, 16, 32, "dot")])
def test_exact_indices_with_duplicates(n, nq, k, metric):
    gal = synth.make_items(n, 512, seed=n, dup=min(1000, n // 4))
    if metric == "l2":  # break the |g|^2 == 2 degeneracy so that the bias term matters
        gal = gal * (1.0 + 0.2 * synth.make_queries(n, 1, seed=n + 1)).astype(np.float32)
    q = synth.make_queries(nq, 1024, seed=n + 2)
    idx, score = _search(q, gal, k, metric, id_offset=12345)
    want_i, want_s = R.search(q, gal, k=k, metric=metric, id_offset=12345)
    assert np.array_equal(idx, want_i)
    np.testing.assert_allclose(score, want_s, rtol=1e-13, atol=1e-11)


def test_small_and_empty_galleries():
    q = synth.make_queries(4, 1024, seed=1)
    gal = synth.make_items(6, 512, seed=2)
    idx, score = _search(q, gal, 10)
    want_i, want_s = R.search(q, gal, k=10)
    assert np.array_equal(idx, want_i)          # 6 real entries then -1 padding
    assert np.all(idx[:, 6:] == -1) and np.all(np.isneginf(score[:, 6:]))
    idx, score = _search(q, gal[:0], 3)
    assert np.all(idx == -1) and np.all(np.isneginf(score))
    idx, _ = _search(q[:0], gal, 3)
    assert idx.shape == (0, 3)


def test_bf16_pass_recall():
    """The bf16 path = bf16 tensor-core contraction + fused top-K' + re-rank of the K' candidates:
    recall@10 >= 0.999 (north_star).  Ranking by the raw bf16 scores alone (exact=False, no fp32
    rows resident) is the documented ~0.993 of SURVEY.md H6, which is why the path over-fetches."""
    n, nq = 200_000, 512
    gal = synth.make_items(n, 512, seed=77)
    q = synth.make_queries(nq, 1024, seed=78)
    want_i, _ = R.search(q, gal, k=10)

    def recall(idx):
        return sum(len(set(a) & set(b)) for a, b in zip(idx, want_i)) / want_i.size

    idx, _ = _search(q, gal, 10, "l2", exact=True)
    assert recall(idx) >= 0.999
    assert np.array_equal(idx, want_i)
    raw, _ = _search(q, gal, 10, "l2", exact=False)
    assert recall(raw) >= 0.985


def test_sharded_search_equals_single(golden_dir):
    """Gallery sharding emulated on one GPU: per-shard local search with global ids, then the
    merge kernel -- must equal the unsharded search and the oracle, for every world size."""
    from outfitx_b200.search import Gallery, local_search, merge_lists, shard_rows
    n, nq, k = 30_001, 200, 10
    gal = synth.make_items(n, 512, seed=5, dup=1000)
    q = synth.make_queries(nq, 1024, seed=6)
    want_i, want_s = R.search(q, gal, k=k)
    full = _t(gal)
    for world in (1, 2, 4, 8):
        parts = []
        for r in range(world):
            lo, hi = shard_rows(n, r, world)
            G = Gallery.build(full[lo:hi], id_offset=lo)
            parts.append(local_search(_t(q), G, k))
        idx = torch.stack([p[0] for p in parts])
        score = torch.stack([p[1] for p in parts])
        mi, ms = merge_lists(idx, score, k)
        assert np.array_equal(mi.cpu().numpy(), want_i)
        np.testing.assert_allclose(ms.cpu().numpy(), want_s, rtol=1e-13, atol=1e-11)


def test_queries_from_the_encoder_find_planted_items():
    """End to end CIR: encoder-produced queries, gallery with each query's own embedding planted
    at a known row -> rank 0 must be the planted row (distance 0)."""
    import outfitx_b200 as o
    sd = synth.make_state_dict(1024, 1024, seed=0)
    m = o.OutfitX(o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type="clip")), precision="bf16")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(DEV)
    B = 64
    emb, mask, _ = synth.make_outfits(B, "concat", seed=91)
    text = synth.make_text_prefix(B, 512, seed=92)
    qv = m.cir_embed(_t(emb), _t(mask), _t(text))
    gal = _t(synth.make_items(10_000, 512, seed=93))
    rows = torch.arange(B, device=DEV) * 150 + 7
    gal[rows] = qv
    from outfitx_b200.search import Gallery, cir_search
    idx, score = cir_search(qv, Gallery.build(gal), k=10, metric="l2")
    assert torch.equal(idx[:, 0], rows)


def test_full_size_properties_config3():
    """BASELINE.json configs[2] size (4096 queries x 1 M rows, top-10): the fp64 oracle cannot sweep
    that in seconds, so check size-independent properties: planted rows come back at rank 0 with
    distance 0, every list is sorted by (-score, idx), all ids are valid and distinct, an oracle
    re-score of the returned rows reproduces the returned fp64 scores, a 4-way row-sharded search +
    merge equals the unsharded one, and a query subset is answered identically (batch invariance)."""
    from outfitx_b200.search import Gallery, local_search, merge_lists, shard_rows
    n, nq, k = 1_000_000, 4096, 10
    g = torch.Generator(device=DEV).manual_seed(123)
    gal = torch.nn.functional.normalize(torch.randn(n, 2, 512, device=DEV, generator=g), dim=-1).reshape(n, 1024)
    q = torch.randn(nq, 1024, device=DEV, generator=g) * 0.05
    planted = torch.arange(0, 256, device=DEV) * 3907 + 11
    q[:256] = gal[planted]
    G = Gallery.build(gal)
    idx, score = local_search(q, G, k)
    assert torch.equal(idx[:256, 0], planted)
    assert torch.all(score[:256, 0] >= -1e-6 + 0.5 * (gal[planted].double() ** 2).sum(-1) - 1e-6)   # q.g - |g|^2/2 = |g|^2/2
    assert int(idx.min()) >= 0 and int(idx.max()) < n
    assert all(len(set(row)) == k for row in idx[:512].cpu().tolist())
    ds = score[:, 1:] - score[:, :-1]
    assert torch.all((ds < 0) | ((ds == 0) & (idx[:, 1:] > idx[:, :-1])))
    # returned scores are the exact fp64 scores of the returned rows
    sel = slice(1000, 1032)
    rows = gal[idx[sel]].double()                                                    # (32, k, 1024)
    want = (rows * q[sel].double()[:, None, :]).sum(-1) - 0.5 * (rows * rows).sum(-1)
    torch.testing.assert_close(score[sel], want, rtol=1e-12, atol=1e-10)
    # nothing better was missed: a brute-force fp32 sweep for a few queries
    few = q[2000:2008]
    full = few @ gal.T - 0.5 * (gal * gal).sum(-1)
    top = torch.topk(full, k, dim=-1)
    assert torch.equal(torch.sort(top.indices, -1).values, torch.sort(idx[2000:2008], -1).values)
    # sharded == unsharded
    parts = []
    for r in range(4):
        lo, hi = shard_rows(n, r, 4)
        parts.append(local_search(q, Gallery.build(gal[lo:hi], id_offset=lo), k))
    mi, ms = merge_lists(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), k)
    assert torch.equal(mi, idx) and torch.equal(ms, score)
    # batch invariance
    sub_i, sub_s = local_search(q[512:1024], G, k)
    assert torch.equal(sub_i, idx[512:1024]) and torch.equal(sub_s, score[512:1024])


def _near_duplicate_gallery(n, n_dup, seed):
    """A catalogue with a cluster of n_dup near-duplicates: rows that differ from a base row by ~1e-4, far
    below what a bf16 copy resolves (2^-9 relative), so the tensor-core pass sees them as (nearly) tied."""
    gal = synth.make_items(n, 512, seed=seed)
    r = np.random.Generator(np.random.PCG64(seed + 1))
    rows = r.choice(n, size=n_dup, replace=False)
    base = gal[rows[0]].copy()
    gal[rows] = base[None, :] + (r.standard_normal((n_dup, 1024)) * 1e-4).astype(np.float32)
    return gal, base, rows


@pytest.mark.parametrize("n,n_dup,k", [(50_000, 24, 10), (50_000, 400, 10), (200_000, 3000, 10), (20_000, 300, 50)])
def test_near_duplicates_are_ranked_exactly(n, n_dup, k):
    """ADVICE r1 / VERDICT 9: near-duplicate gallery rows whose scores differ by less than bf16 rounding.  With a
    handful of them the certificate's extension re-scores the whole band; with more of them than the lists hold
    (400 / 3000 > kcap) the query must come back UNcertified and be answered by the exhaustive fp64 fallback.
    Either way the indices are those of the fp64 oracle."""
    from outfitx_b200.search import Gallery, SearchStats, local_search
    gal, base, rows = _near_duplicate_gallery(n, n_dup, seed=n + n_dup)
    q = synth.make_queries(40, 1024, seed=3) * np.float32(0.05)
    q[:8] = base[None, :] * np.float32(0.7) + q[:8] * np.float32(0.1)     # queries that look straight at the cluster
    G = Gallery.build(_t(gal))
    SearchStats.reset()
    idx, score, proved = local_search(_t(q), G, k, "l2", True, return_certified=True)
    want_i, want_s = R.search(q, gal, k=k)
    assert np.array_equal(idx.cpu().numpy(), want_i)
    np.testing.assert_allclose(score.cpu().numpy(), want_s, rtol=1e-13, atol=1e-11)
    proved = proved.cpu().numpy()
    assert proved[8:].all()                      # ordinary queries are proved by the bound alone
    if n_dup > 2500:
        # more near-ties than ALL the per-segment lists together can hold (<= 64 segments x 32): only the exhaustive
        # fallback can know.  (With a few hundred, the lists of the gallery's segments happen to keep them all and the
        # extension proves the result without the fallback -- equally exact, asserted above.)
        assert not proved[:8].any()
        assert SearchStats.uncertified == 8
    # the bf16 ranking alone really is wrong here (otherwise this test tests nothing)
    raw, _ = local_search(_t(q), G, k, "l2", False)
    assert not np.array_equal(raw.cpu().numpy()[:8], want_i[:8])


def test_exhaustive_fallback_alone():
    """ofx_exact_search on its own: every query through the fp64 brute force, ragged sizes, both metrics."""
    from outfitx_b200 import _lib
    L = _lib.lib()
    for n, nq, k, metric in ((5000, 19, 10, "l2"), (300, 9, 64, "dot"), (70_000, 3, 1, "l2"), (40, 5, 50, "l2")):
        gal = synth.make_items(n, 512, seed=n, dup=n // 5)
        q = synth.make_queries(nq, 1024, seed=n + 1)
        g, qq = _t(gal), _t(q)
        score = torch.full((nq, k), 7.0, dtype=torch.float64, device=DEV)
        idx = torch.full((nq, k), -7, dtype=torch.int64, device=DEV)
        sel = torch.arange(nq - 1, -1, -2, dtype=torch.int32, device=DEV)       # every other query, reversed
        cert = torch.zeros(nq, dtype=torch.uint8, device=DEV)
        ws = torch.empty(L.ofx_exact_search_workspace_bytes(n, sel.numel(), k), dtype=torch.uint8, device=DEV)
        _lib.check(L.ofx_exact_search(g.data_ptr(), n, 1024, 1000, qq.data_ptr(), sel.data_ptr(), sel.numel(), k,
                                      1 if metric == "l2" else 0, score.data_ptr(), idx.data_ptr(), cert.data_ptr(),
                                      ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
        want_i, want_s = R.search(q, gal, k=k, metric=metric, id_offset=1000)
        s = sel.cpu().numpy()
        assert np.array_equal(idx.cpu().numpy()[s], want_i[s])
        got_s = score.cpu().numpy()[s]
        fin = np.isfinite(want_s[s])
        np.testing.assert_allclose(got_s[fin], want_s[s][fin], rtol=1e-13, atol=1e-11)
        untouched = np.setdiff1d(np.arange(nq), s)
        assert np.all(idx.cpu().numpy()[untouched] == -7) and np.all(cert.cpu().numpy()[untouched] == 0)
        assert np.all(cert.cpu().numpy()[s] == 1)


def test_certificate_on_ordinary_data():
    """On the benchmark's kind of data (random 1024-d items) every query is proved by the bound alone, at the
    headline k = 10 and at the reference's k = 50."""
    from outfitx_b200.search import Gallery, local_search
    gal = synth.make_items(300_000, 512, seed=31)
    q = synth.make_queries(512, 1024, seed=32) * np.float32(0.05)
    G = Gallery.build(_t(gal))
    for k in (10, 50):
        _, _, proved = local_search(_t(q), G, k, "l2", True, return_certified=True)
        assert bool(proved.all())


@pytest.mark.parametrize("case", ["flat_sample", "outlier_in_sample", "negative_dot", "few_queries_k50"])
def test_counting_bound_on_awkward_samples(case):
    """The seeded bound and the counting histogram are derived from the FIRST rows of the shard (up to four sample
    segments of pooled block maxima).  Samples that say nothing about the rest -- all rows identical (zero bucket
    width), one huge outlier (one bucket holds everything), all scores negative (dot metric) -- must only cost
    speed: indices stay those of the fp64 oracle and the scores stay exact."""
    from outfitx_b200.search import Gallery, local_search
    n, k, metric, nq = 560_000, 10, "l2", 40
    gal = synth.make_items(n, 512, seed=77)
    q = synth.make_queries(nq, 1024, seed=78) * np.float32(0.05)
    if case == "flat_sample":
        gal[:40_000] = gal[0]
    elif case == "outlier_in_sample":
        gal[1234] *= np.float32(40.0)
        q[:4] = gal[1234] * np.float32(0.02) + q[:4]
    elif case == "negative_dot":
        metric = "dot"
        gal = -np.abs(gal)
        q = np.abs(q)
    elif case == "few_queries_k50":
        k, nq = 50, 3
        q = q[:3]
    G = Gallery.build(_t(gal))
    idx, score, proved = local_search(_t(q), G, k, metric, True, return_certified=True)
    want_i, want_s = R.search(q, gal, k=k, metric=metric)
    assert np.array_equal(idx.cpu().numpy(), want_i)
    np.testing.assert_allclose(score.cpu().numpy(), want_s, rtol=1e-13, atol=1e-11)
