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
