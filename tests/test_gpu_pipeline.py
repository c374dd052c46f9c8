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
