"""CPU: pin the oracle (numpy restatement + stock-torch port) to the reference's own outputs.

The golden files were produced by the UNMODIFIED reference (oracle/make_golden.py).
"""
import os

import numpy as np
import pytest
import torch

from oracle import restatement as R
from oracle import torch_port
from oracle.make_golden import CASES, case_dims, case_inputs, digest, processor_items
from outfitx_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name,method,d_model,batch", CASES)
def test_inputs_regenerate_bit_exactly(golden_dir, name, method, d_model, batch):
    g = _load(golden_dir, f"model_{name}.npz")
    img, txt, emb, mask, text, cand = case_inputs(method, batch, dpm=case_dims(method, d_model)[0])
    assert digest(emb, mask, text, cand) == str(g["input_digest"])
    sd = synth.make_state_dict(d_model, case_dims(method, d_model)[1], seed=0)
    assert digest(*sd.values()) == str(g["weight_digest"])
    assert np.array_equal(mask, g["mask"])


@pytest.mark.parametrize("name,method,d_model,batch", CASES)
def test_restatement_matches_reference(golden_dir, name, method, d_model, batch):
    g = _load(golden_dir, f"model_{name}.npz")
    sd = synth.make_state_dict(d_model, case_dims(method, d_model)[1], seed=0)
    _, _, emb, mask, text, cand = case_inputs(method, batch, dpm=case_dims(method, d_model)[0])
    for dt, tol in ((np.float32, 2e-5), (np.float64, 1e-5)):
        logits = R.cp_forward(sd, emb, mask, dt)
        query = R.cir_forward(sd, emb, mask, text, dt)
        np.testing.assert_allclose(logits, g["logits"], rtol=0, atol=tol)
        np.testing.assert_allclose(query, g["query"], rtol=0, atol=tol)
    np.testing.assert_allclose(R.sigmoid(g["logits"]), g["probs"], atol=1e-6)
    idx, d = R.fitb(g["query"].astype(np.float64), cand.astype(np.float64))
    np.testing.assert_allclose(d, g["fitb_dists"], atol=1e-4)
    assert np.array_equal(idx, g["fitb_argmin"])


@pytest.mark.parametrize("name,method,d_model,batch", CASES)
def test_torch_port_matches_reference(golden_dir, name, method, d_model, batch):
    g = _load(golden_dir, f"model_{name}.npz")
    port = torch_port.ReferencePort.from_numpy(synth.make_state_dict(d_model, case_dims(method, d_model)[1], seed=0))
    _, _, emb, mask, text, cand = case_inputs(method, batch, dpm=case_dims(method, d_model)[0])
    t = torch.from_numpy
    logits = port.cp(t(emb), t(mask)).numpy()
    query = port.cir(t(emb), t(mask), t(text))
    # same library kernels as the reference -> bitwise on the same host, tiny slack elsewhere
    np.testing.assert_allclose(logits, g["logits"], atol=1e-5)
    np.testing.assert_allclose(query.numpy(), g["query"], atol=1e-5)
    idx, d = torch_port.fitb(query, t(cand))
    assert np.array_equal(idx.numpy(), g["fitb_argmin"])


def test_fusion_matches_reference(golden_dir):
    g = _load(golden_dir, "fusion.npz")
    img, txt = synth.make_modalities(4, 512, int(g["seed"]))
    assert digest(img, txt) == str(g["input_digest"])
    np.testing.assert_allclose(R.fuse(img, txt, "concat"), g["concat"], atol=1e-6)
    np.testing.assert_allclose(synth.fuse(img, txt, "concat"), g["concat"], atol=1e-6)
    # the intended elementwise mean equals the literal reference code on 1-D inputs (D5)
    np.testing.assert_allclose(R.fuse(img, txt, "mean")[0], g["mean_row0"], atol=1e-6)
    with pytest.raises(ValueError):
        R.fuse(img, txt, "sum")  # in the reference's Literal but unimplemented there too


def test_search_oracle_matches_reference_idiom(golden_dir):
    g = _load(golden_dir, "search_pool3000.npz")
    pool = synth.make_items(3000, 512, seed=int(g["pool_seed"]))
    q = synth.make_queries(64, 1024, seed=int(g["query_seed"])) * np.float32(0.05)
    assert digest(pool, q) == str(g["input_digest"])
    idx, score = R.search(q, pool, k=50, metric="l2")
    # |g|^2 == 2 for every item, so dot and l2 rankings coincide
    idx_dot, _ = R.search(q, pool, k=50, metric="dot")
    assert np.array_equal(idx, idx_dot)
    # reference: topk(cdist) in fp32.  It may only differ from the fp64 oracle where two
    # neighbours are closer than fp32 cdist noise (a swapped near-tie).
    ref = g["indices"]
    mism = idx != ref
    assert mism.mean() < 0.005 and not mism[:, :10].any()
    full = R.search_scores(q, pool)
    rows = np.nonzero(mism)[0]
    assert np.all(np.abs(full[rows, idx[mism]] - full[rows, ref[mism]]) < 1e-6)
    d = np.sqrt(np.maximum((q.astype(np.float64) ** 2).sum(-1)[:, None] - 2 * score, 0))
    np.testing.assert_allclose(d, g["dists"], atol=1e-4)


def test_topk_tie_break_lowest_index():
    s = np.array([[1.0, 3.0, 3.0, 3.0, 2.0, 3.0]])
    idx, val = R.topk_lex(s, 3)
    assert idx.tolist() == [[1, 2, 3]]  # torch.topk gives [3,5,1] here
    gal = synth.make_items(500, 512, seed=9, dup=50)
    q = synth.make_queries(8, 1024, seed=10)
    idx, score = R.search(q, gal, k=20)
    full_i, full_s = R.topk_lex(R.search_scores(q, gal), 20)
    assert np.array_equal(idx, full_i)
    for r in range(8):  # equal scores -> ascending index
        for a in range(19):
            if score[r, a] == score[r, a + 1]:
                assert idx[r, a] < idx[r, a + 1]


def test_merge_is_shard_invariant():
    gal = synth.make_items(4000, 512, seed=12, dup=100)
    q = synth.make_queries(16, 1024, seed=13)
    want_i, want_s = R.search(q, gal, k=10)
    for world in (2, 4, 8):
        per = len(gal) // world
        parts = [R.topk_lex(R.search_scores(q, gal[r * per:(r + 1) * per]), 10, r * per)
                 for r in range(world)]
        i, s = R.merge_topk(np.concatenate([p[0] for p in parts], 1),
                            np.concatenate([p[1] for p in parts], 1), 10)
        assert np.array_equal(i, want_i) and np.array_equal(s, want_s)


def test_collate_contract_matches_reference_processors(golden_dir):
    """tests/golden/processor_clip1024.npz holds what the reference's OWN processors emitted for lists of its
    task objects (oracle/make_golden.py: write_processor_golden).  The contract the CUDA path is built on
    (truncate to 16, zero pad rows, mask True on pad, valid items left-aligned, text = second half
    of the target item's embedding) must regenerate those tensors bit for bit, and the oracle must reproduce the
    reference model's answers to them."""
    g = _load(golden_dir, "processor_clip1024.npz")
    outfits, targets, cands = processor_items(int(g["seed"]))
    B = len(outfits)
    emb = np.zeros((B, 16, 1024), np.float32)
    mask = np.ones((B, 16), bool)
    for b, o in enumerate(outfits):
        n = min(len(o), 16)
        emb[b, :n] = np.stack(o[:n])
        mask[b, :n] = False
    for pre in ("cp", "cir"):
        assert np.array_equal(g[pre + "_outfit_embedding"], emb)
        assert np.array_equal(g[pre + "_outfit_mask"], mask)
    assert np.array_equal(g["cir_text"], np.stack(targets)[:, 512:])
    assert np.array_equal(g["fitb_cand"], np.stack(cands))
    assert str(g["cp_task"]) == "OutfitCompatibilityPredictionTask"
    assert str(g["cir_task"]) == str(g["fitb_task"]) == "OutfitComplementaryItemRetrievalTask"   # FITB dispatches as CIR
    sd = synth.make_state_dict(1024, 1024, seed=int(g["weight_seed"]))
    np.testing.assert_allclose(R.cp_forward(sd, emb, mask), g["logits"], atol=2e-5)
    q = R.cir_forward(sd, emb, mask, g["cir_text"])
    np.testing.assert_allclose(q, g["query"], atol=2e-5)
    idx, d = R.fitb(g["query"].astype(np.float64), g["fitb_cand"].astype(np.float64))
    assert np.array_equal(idx, g["fitb_argmin"])
