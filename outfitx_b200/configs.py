"""Model configuration dataclasses with the reference's field names and derived values.

Mirrors ``src/models/configs/{outfit_x_config,transformer_config,item_encoder_config}.py``
of the reference (same names, same defaults, same ``__post_init__`` rules) so that code
written against ``OutfitXConfig`` keeps working.  Two deliberate normalisations:
``batch_first`` / ``norm_first`` are real booleans here (the reference's trailing commas make
them the truthy tuples ``(True,)``, transformer_config.py:20-21), and
``activation`` is the string ``'mish'`` (the reference stores ``F.mish``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Literal

_DIM_PER_MODALITY = {"clip": 512, "resnet_hf_sentence_bert": 64, "slip": 768}
_MODEL_NAME = {"clip": "patrickjohncyh/fashion-clip",
               "resnet_hf_sentence_bert": "sentence-transformers/all-MiniLM-L6-v2",
               "slip": "hf-hub:Marqo/marqo-fashionSigLIP"}


@dataclass
class ItemEncoderConfig:
    """item_encoder_config.py:5-29."""
    type: Literal["clip", "resnet_hf_sentence_bert", "slip"] = "slip"
    norm_out: bool = True
    aggregation_method: Literal["concat", "sum", "mean"] = "concat"

    def __post_init__(self):
        if self.type not in _DIM_PER_MODALITY:
            raise ValueError(f"Unsupported type: {self.type}")
        self.dim_per_modality: int = _DIM_PER_MODALITY[self.type]
        name = _MODEL_NAME[self.type]
        if self.type == "clip":
            self.clip_model_name = name
        elif self.type == "slip":
            self.slip_model_name = name
        else:
            self.text_model_name = name

    @property
    def d_embed(self) -> int:
        """ItemEncoder.d_embed (item_encoder.py:38-40): the encoder's d_model."""
        d = self.dim_per_modality
        return d * 2 if self.aggregation_method == "concat" else d


@dataclass
class TransformerConfig:
    """transformer_config.py:7-23."""
    n_head: int = 16
    d_ffn: int = 2024
    n_layers: int = 6
    dropout: float = 0.3
    norm_out: bool = False
    batch_first: bool = True
    norm_first: bool = True
    activation: str = "mish"
    enable_nested_tensor: bool = False


@dataclass
class OutfitXConfig:
    """outfit_x_config.py:8-30."""
    padding: Literal["longest", "max_length"] = "max_length"
    max_length: int = 16
    truncation: bool = True
    d_embed: int = 1024
    item_encoder: ItemEncoderConfig = field(default_factory=ItemEncoderConfig)
    transformer: TransformerConfig = field(default_factory=TransformerConfig)

    def __post_init__(self):
        self.d_embed = self.item_encoder.dim_per_modality * 2
        self.model_name = _MODEL_NAME[self.item_encoder.type].split("/")[-1]
