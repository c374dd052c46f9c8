"""GPU: the host scoring pipeline (chunked, double-buffered H2D overlapped with scoring) returns
exactly what the direct device-resident calls return, for ragged chunk counts."""
import numpy as np
import pytest
import torch

from outfitx_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch,chunk", [(100, 32), (64, 64), (7, 16), (300, 128)])
def test_pipeline_equals_direct_calls(batch, chunk):
    import outfitx_b200 as o
    from outfitx_b200.pipeline import HostScoringPipeline
    sd = synth.make_state_dict(512, 1024, seed=0)
    m = o.OutfitX(o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type="clip", aggregation_method="mean")))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to("cuda")
    img, txt = synth.make_modalities(batch, 512, seed=batch)
    mask = synth.make_mask(synth.make_lengths(batch, batch + 1))
    text = synth.make_text_prefix(batch, 256, batch + 2)
    cand = synth.make_items(batch * 4, 512, batch + 3).reshape(batch, 4, 1024)
    host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (img, txt, mask, text, cand)]
    out = HostScoringPipeline(m, chunk=chunk).score(*host)
    dev = [h.cuda() for h in host]
    enc = {"image_embeddings": dev[0], "text_embeddings": dev[1]}
    probs = m.score_cp(outfit_mask=dev[2], encoder_input_dict=enc)
    pred, _, _ = m.score_fitb(outfit_mask=dev[2], target_item_text_embedding=dev[3],
                              candidate_item_embedding=dev[4], encoder_input_dict=enc)
    # chunking changes which rows share a GEMM tile, not the arithmetic of a row
    torch.testing.assert_close(out["probs"], probs.cpu(), rtol=0, atol=1e-6)
    assert torch.equal(out["pred"], pred.cpu())
    cp_only = HostScoringPipeline(m, chunk=chunk).score(host[0], host[1], host[2])
    torch.testing.assert_close(cp_only["probs"], probs.cpu(), rtol=0, atol=1e-6)
    assert "pred" not in cp_only


def test_device_side_collate_equals_host_gather():
    """item-id gather from HBM-resident tables == the reference-style padded batch
    built on the host (outfit_x_base_processor.py:20-43), through the model API and the pipeline."""
    import outfitx_b200 as o
    from outfitx_b200.pipeline import HostScoringPipeline
    sd = synth.make_state_dict(512, 1024, seed=0)
    m = o.OutfitX(o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type="clip", aggregation_method="mean")))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to("cuda")
    n_items, B = 5000, 333
    rng = np.random.Generator(np.random.PCG64(7))
    img_t = rng.standard_normal((n_items, 512)).astype(np.float32)
    txt_t = rng.standard_normal((n_items, 512)).astype(np.float32)
    cand_t = synth.make_items(n_items, 512, seed=8)                       # fused 1024-d candidates
    ids = rng.integers(0, n_items, size=(B, 16)).astype(np.int32)
    cids = rng.integers(0, n_items, size=(B, 4)).astype(np.int32)
    mask = synth.make_mask(synth.make_lengths(B, 9))
    ids[mask] = -1                                                          # padded slots: any id
    text = synth.make_text_prefix(B, 256, 10)
    safe = np.where(ids < 0, 0, ids)
    dense = [torch.from_numpy(a).cuda() for a in (img_t[safe], txt_t[safe], mask, text, cand_t[cids])]
    enc = {"image_embeddings": dense[0], "text_embeddings": dense[1]}
    want_p = m.score_cp(outfit_mask=dense[2], encoder_input_dict=enc)
    want_pred, want_d, _ = m.score_fitb(outfit_mask=dense[2], target_item_text_embedding=dense[3],
                                        candidate_item_embedding=dense[4], encoder_input_dict=enc)
    tabs = [torch.from_numpy(a).cuda() for a in (img_t, txt_t, cand_t)]
    enc_ids = {"image_embeddings": tabs[0], "text_embeddings": tabs[1], "item_ids": torch.from_numpy(ids).cuda()}
    got_p = m.score_cp(outfit_mask=dense[2], encoder_input_dict=enc_ids)
    got_pred, got_d, _ = m.score_fitb(outfit_mask=dense[2], target_item_text_embedding=dense[3],
                                      candidate_item_embedding=(tabs[2], torch.from_numpy(cids).cuda()),
                                      encoder_input_dict=enc_ids)
    assert torch.equal(got_p, want_p) and torch.equal(got_pred, want_pred) and torch.equal(got_d, want_d)
    pipe = HostScoringPipeline(m, chunk=128)
    host = [torch.from_numpy(a).pin_memory() for a in (ids, mask, text, cids)]
    out = pipe.score_ids(host[0], host[1], tabs[0], tabs[1], host[2], host[3], tabs[2])
    torch.testing.assert_close(out["probs"], want_p.cpu(), rtol=0, atol=1e-6)
    assert torch.equal(out["pred"], want_pred.cpu())


def test_packed_host_layout_equals_padded():
    """a collate that skips zero padding hands over the VALID item rows only -- (sum n_i, dpm) per
    modality + lengths -- and the pipeline moves each chunk with one DMA per modality (44 % fewer PCIe bytes at
    n ~ U{2..16}).  Results must be bit-identical to the padded (B, 16, dpm) path, pinned or pageable, for chunk
    sizes that do and do not divide the batch, including outfits with 0 and 16 items."""
    import outfitx_b200 as o
    from outfitx_b200.pipeline import HostScoringPipeline, pack_valid_rows
    sd = synth.make_state_dict(512, 1024, seed=0)
    m = o.OutfitX(o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type="clip", aggregation_method="mean")))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to("cuda")
    B = 257
    img, txt = synth.make_modalities(B, 512, seed=41)
    lengths = synth.make_lengths(B, 42)
    lengths[:3] = (0, 16, 1)
    mask = synth.make_mask(lengths)
    text = synth.make_text_prefix(B, 256, 43)
    cand = synth.make_items(B * 4, 512, 44).reshape(B, 4, 1024)
    pageable = [torch.from_numpy(np.ascontiguousarray(a)) for a in (img, txt, mask, text, cand)]
    pinned = [t.pin_memory() for t in pageable]
    want = HostScoringPipeline(m, chunk=64).score(*pinned)
    img_rows, txt_rows, lens = pack_valid_rows(pageable[0], pageable[1], pageable[2])
    assert img_rows.shape == (int(lengths.sum()), 512) and torch.equal(lens, torch.from_numpy(lengths).to(lens.dtype))
    assert torch.equal(img_rows[:16], pageable[0][1])            # outfit 0 is empty, outfit 1 owns the first 16 rows
    for chunk, pin in ((64, True), (100, False), (1000, True)):
        rows = [img_rows.pin_memory(), txt_rows.pin_memory()] if pin else [img_rows, txt_rows]
        got = HostScoringPipeline(m, chunk=chunk).score_packed(rows[0], rows[1], lens, pinned[3] if pin else pageable[3],
                                                               pinned[4] if pin else pageable[4])
        assert torch.equal(got["probs"], want["probs"]) and torch.equal(got["pred"], want["pred"])
    # the same pipeline object again and again: the first call runs eagerly and captures one CUDA graph per staging
    # slot, the following calls replay them -- on different data each time (the graphs must not have baked values in)
    pipe = HostScoringPipeline(m, chunk=64)
    eager = HostScoringPipeline(m, chunk=64, use_graphs=False)
    for rep in range(3):
        perm = torch.randperm(B, generator=torch.Generator().manual_seed(rep))
        pg = [t[perm].contiguous() for t in pageable]
        ir, tr, ln = pack_valid_rows(pg[0], pg[1], pg[2])
        a = pipe.score_packed(ir, tr, ln, pg[3], pg[4])
        b = eager.score_packed(ir, tr, ln, pg[3], pg[4])
        assert torch.equal(a["probs"], b["probs"]) and torch.equal(a["pred"], b["pred"])
        assert torch.equal(a["probs"], want["probs"][perm]) and torch.equal(a["pred"], want["pred"][perm])
    assert any(e[0] is not None for e in pipe._graphs.values())        # graphs really were captured and replayed
    cp_only = HostScoringPipeline(m, chunk=128).score_packed(img_rows, txt_rows, lens)
    assert torch.equal(cp_only["probs"], want["probs"]) and "pred" not in cp_only
    with pytest.raises(ValueError):
        HostScoringPipeline(m, chunk=64).score_packed(img_rows[:-1], txt_rows, lens)


def test_pipeline_pinned_and_pageable_agree():
    import outfitx_b200 as o
    from outfitx_b200.pipeline import HostScoringPipeline
    sd = synth.make_state_dict(512, 1024, seed=0)
    m = o.OutfitX(o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type="clip", aggregation_method="mean")))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to("cuda")
    B = 130
    img, txt = synth.make_modalities(B, 512, seed=41)
    mask = synth.make_mask(synth.make_lengths(B, 42))
    mask[3] = np.array([1, 0, 1, 1, 0, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0], bool)      # not left-aligned
    text = synth.make_text_prefix(B, 256, 43)
    cand = synth.make_items(B * 4, 512, 44).reshape(B, 4, 1024)
    pageable = [torch.from_numpy(np.ascontiguousarray(a)) for a in (img, txt, mask, text, cand)]
    pinned = [t.pin_memory() for t in pageable]
    a = HostScoringPipeline(m, chunk=64).score(*pinned)
    b = HostScoringPipeline(m, chunk=64).score(*pageable)
    assert torch.equal(a["probs"], b["probs"]) and torch.equal(a["pred"], b["pred"])
