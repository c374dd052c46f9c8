"""Task classes: on the hot path they are only dispatch keys for ``OutfitX.forward``
(``src/models/outfit_x.py:84-90``; definitions in ``src/models/datatypes/*.py``).

The reference's classes are pydantic models describing whole outfits (images, text, ids);
none of that reaches the scoring path, so these are light stand-ins with the same names.
``OutfitX.forward`` also accepts the reference's own classes (matched by class name), so a
caller that imports ``src.models.datatypes`` keeps working after swapping the model class.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional


@dataclass
class FashionItem:
    item_id: Optional[int] = None
    category: Optional[str] = ""
    image: Any = None
    description: Optional[str] = ""
    metadata: dict = field(default_factory=dict)
    embedding: Any = None
    text_embedding: Any = None


@dataclass
class OutfitCompatibilityPredictionTask:
    outfit: List[FashionItem] = field(default_factory=list)

    def __len__(self):
        return len(self.outfit)


@dataclass
class OutfitComplementaryItemRetrievalTask:
    outfit: List[FashionItem] = field(default_factory=list)
    target_item: FashionItem = field(default_factory=FashionItem)

    def __len__(self):
        return len(self.outfit)


@dataclass
class OutfitFillInTheBlankTask:
    """Same fields as the CIR task; only the task type differs (outfit_fitb_task.py)."""
    outfit: List[FashionItem] = field(default_factory=list)
    target_item: FashionItem = field(default_factory=FashionItem)

    def __len__(self):
        return len(self.outfit)


@dataclass
class OutfitPrecomputeEmbeddingTask:
    """Runs the frozen image / text encoders -- upstream of this package's scope."""
    items: List[FashionItem] = field(default_factory=list)
