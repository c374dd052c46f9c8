"""CPU: the host-side mirror of the reference's src/models interface -- config rules, parameter
names / shapes, task dispatch, error behaviour -- compared with the reference itself when it is
mounted (build container), and with the committed fixtures otherwise."""
import numpy as np
import pytest
import torch

import outfitx_b200 as o
from oracle import ref_shim
from outfitx_b200 import synth
from outfitx_b200.search import shard_rows


@pytest.mark.parametrize("enc_type,method,d_model,d_embed", [
    ("clip", "concat", 1024, 1024), ("clip", "mean", 512, 1024), ("slip", "concat", 1536, 1536)])
def test_config_rules(enc_type, method, d_model, d_embed):
    cfg = o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type=enc_type, aggregation_method=method))
    assert cfg.item_encoder.d_embed == d_model and cfg.d_embed == d_embed
    assert (cfg.transformer.n_head, cfg.transformer.d_ffn, cfg.transformer.n_layers) == (16, 2024, 6)
    assert cfg.max_length == 16 and cfg.padding == "max_length" and cfg.truncation is True
    with pytest.raises(ValueError):
        o.ItemEncoderConfig(type="vgg")


@pytest.mark.parametrize("method,d_model", [("concat", 1024), ("mean", 512)])
def test_state_dict_matches_reference_layout(method, d_model):
    cfg = o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type="clip", aggregation_method=method))
    m = o.OutfitX(cfg)
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    want = {k: v.shape for k, v in synth.make_state_dict(d_model, 1024).items()}
    assert ours == want
    assert sum(p.numel() for p in m.parameters()) == {1024: 51_155_313, 512: 19_292_273}[d_model]
    if ref_shim.available():
        ref = ref_shim.build_reference_model(method)
        assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == ours
        rc = ref.cfg
        assert (rc.d_embed, rc.max_length, rc.transformer.d_ffn, rc.transformer.n_head,
                rc.transformer.n_layers, rc.item_encoder.dim_per_modality, rc.model_name) == (
            cfg.d_embed, cfg.max_length, cfg.transformer.d_ffn, cfg.transformer.n_head,
            cfg.transformer.n_layers, cfg.item_encoder.dim_per_modality, cfg.model_name)
        # a reference checkpoint (with frozen-encoder keys) loads strictly
        sd = dict(ref.state_dict())
        sd["item_encoder.image_enc.model.weight"] = torch.zeros(3)
        m.load_state_dict(sd)
        assert torch.equal(m.outfit_token, ref.outfit_token)


def test_dispatch_and_errors():
    m = o.OutfitX(o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type="clip")))
    assert m.device == torch.device("cpu") and not m.training
    with pytest.raises(KeyError):
        m(dict, outfit_embedding=None, outfit_mask=None)          # unknown task: reference KeyError
    with pytest.raises(NotImplementedError):
        m(o.OutfitPrecomputeEmbeddingTask, images=[[None]], texts=[[""]])   # upstream encoders
    with pytest.raises(NotImplementedError):
        m.train()
    with pytest.raises(ValueError):
        o.OutfitX(precision="fp8")
    if ref_shim.available():   # the reference's own task classes work as dispatch keys
        _, _, dts = ref_shim.load_reference()
        assert m.forward_[dts.OutfitFillInTheBlankTask.__name__] == m._cir_forward
        with pytest.raises(RuntimeError, match="no CPU path"):
            m(dts.OutfitCompatibilityPredictionTask, outfit_embedding=torch.zeros(1, 16, 1024),
              outfit_mask=torch.zeros(1, 16, dtype=torch.bool))


def test_shard_rows_partition():
    for n in (0, 1, 7, 10_000_000, 30_001):
        for w in (1, 2, 4, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= max(1, w)


def test_embedding_pickle_loader(tmp_path):
    """the reference's precomputed-embedding files (precompute_embedding_script.py:47-53)."""
    import pickle
    from outfitx_b200.search import load_embedding_pickles
    rng = np.random.default_rng(0)
    paths = []
    for r, n in enumerate((5, 3)):
        d = {"ids": [100 * r + i for i in range(n)], "embeddings": rng.standard_normal((n, 1024)).astype(np.float32)}
        p = tmp_path / f"fashion-clip_embedding_subset_{r}.pkl"
        with open(p, "wb") as f:
            pickle.dump(d, f)
        paths.append(str(p))
    ids, emb, index = load_embedding_pickles(paths)
    assert ids.tolist() == [0, 1, 2, 3, 4, 100, 101, 102] and emb.shape == (8, 1024) and emb.dtype == torch.float32
    assert index[101] == 6
    text = emb[:, 512:]            # polyvore_item_dataset.py:75: text embedding = second half
    assert text.shape == (8, 512)


def test_packed_host_layout_arithmetic():
    """The host side of HostScoringPipeline.score_packed (pure CPU): mask, chunk-local row ids and offsets of a
    ragged batch, and the round trip padded -> packed -> gathered-by-ids == padded (valid slots)."""
    import numpy as np
    import torch
    from outfitx_b200.pipeline import pack_valid_rows, packed_layout
    rng = np.random.Generator(np.random.PCG64(7))
    B, L, d = 37, 16, 8
    lengths = torch.from_numpy(rng.integers(0, L + 1, size=B)).to(torch.int64)
    lengths[0], lengths[1] = 0, L                                # the edge cases
    mask = torch.arange(L)[None, :] >= lengths[:, None]          # left-aligned valid items
    img = torch.from_numpy(rng.standard_normal((B, L, d)).astype(np.float32))
    txt = torch.from_numpy(rng.standard_normal((B, L, d)).astype(np.float32))
    ir, tr, lens = pack_valid_rows(img, txt, mask)
    assert torch.equal(lens.to(torch.int64), lengths) and ir.shape == (int(lengths.sum()), d)
    plan = [(0, 10), (10, 11), (11, 30), (30, 37)]                # ragged chunks
    mask_out = torch.empty(B, L, dtype=torch.bool)
    ids = torch.empty(B, L, dtype=torch.int32)
    off = packed_layout(lengths, plan, L, mask_out, ids)
    assert torch.equal(mask_out, mask)
    assert off[0] == 0 and off[-1] == lengths.sum() and torch.equal(off[1:] - off[:-1], lengths)
    for lo, hi in plan:
        rows_i, rows_t = ir[int(off[lo]):int(off[hi])], tr[int(off[lo]):int(off[hi])]      # what the chunk stages
        for b in range(lo, hi):
            n = int(lengths[b])
            sel = ids[b, :n].long()
            assert n == 0 or (int(sel.min()) >= 0 and int(sel.max()) < rows_i.shape[0])
            assert torch.equal(rows_i[sel], img[b, :n]) and torch.equal(rows_t[sel], txt[b, :n])
    # a mask that is not left-aligned keeps the slot order of the valid items
    m2 = mask.clone(); m2[5] = torch.tensor([1, 0, 1, 1, 0, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0], dtype=torch.bool)
    ir2, _, lens2 = pack_valid_rows(img, txt, m2)
    o2 = int(lens2[:5].sum())
    n5 = int((~m2[5]).sum())
    assert n5 == 5 and int(lens2[5]) == n5 and torch.equal(ir2[o2:o2 + n5], img[5][~m2[5]])
