"""TEST INFRASTRUCTURE ONLY -- runs the reference's own ``OutfitX`` on CPU in the build container.

``/root/reference`` exists only in the build container (never on the GPU box), so this
module is used solely by ``oracle/make_golden.py`` and by CPU tests that skip when the
reference is absent.  Recipe: SURVEY.md App. B.  ``import src.models.outfit_x`` fails only
because ``src/models/encoders/item_encoder.py:1`` imports ``open_clip`` (not installed) and
``OutfitX.__init__`` builds an ``ItemEncoder`` that downloads pretrained weights
(``item_encoder.py:20-37``); both are replaced by inert stand-ins, everything else --
``nn.TransformerEncoder`` construction (``outfit_x.py:32-45``), ``_cp_forward`` /
``_cir_forward`` (``:120-172``) -- is the reference's code, unmodified.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("OFX_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models", "outfit_x.py"))


def load_reference():
    """Returns (module src.models.outfit_x, configs module, datatypes module)."""
    import torch
    from torch import nn

    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.modules.setdefault("open_clip", types.ModuleType("open_clip"))
    import src.models.outfit_x as ox  # noqa: E402  (the reference's module)
    import src.models.configs as cfgs
    import src.models.datatypes as dts

    class _NoEncoder(nn.Module):
        """Stand-in for ItemEncoder: keeps cfg and the d_embed rule of item_encoder.py:38-40."""

        def __init__(self, cfg):
            super().__init__()
            self.cfg = cfg

        @property
        def d_embed(self):
            d = self.cfg.dim_per_modality
            return d * 2 if self.cfg.aggregation_method == "concat" else d

    ox.ItemEncoder = _NoEncoder
    return ox, cfgs, dts


def build_reference_model(method: str = "concat", state_dict=None, encoder_type: str = "clip"):
    """Reference OutfitX (eval, fp32, CPU) with ``state_dict`` (numpy arrays) loaded."""
    import torch

    ox, cfgs, _ = load_reference()
    cfg = cfgs.OutfitXConfig(item_encoder=cfgs.ItemEncoderConfig(type=encoder_type,
                                                                   aggregation_method=method))
    torch.manual_seed(0)
    model = ox.OutfitX(cfg).eval()
    if state_dict is not None:
        model.load_state_dict({k: torch.from_numpy(v) for k, v in state_dict.items()}, strict=True)
    return model
