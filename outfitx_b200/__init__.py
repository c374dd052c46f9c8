"""outfitx_b200 -- B200-native (sm_100a) outfit-scoring hot path of Krual-T/OutfitX.

Host side mirrors the reference's ``src/models`` API; the arithmetic lives in ``libofx.so``
(``include/ofx.h``).  Importing this package does not need a GPU; computing anything does.
"""
from .configs import ItemEncoderConfig, OutfitXConfig, TransformerConfig  # noqa: F401
from .datatypes import (FashionItem, OutfitCompatibilityPredictionTask,  # noqa: F401
                        OutfitComplementaryItemRetrievalTask, OutfitFillInTheBlankTask,
                        OutfitPrecomputeEmbeddingTask)

__all__ = ["OutfitX", "aggregate_embeddings", "Gallery", "cir_search", "local_search", "merge_lists",
           "ShardedSearch", "shard_rows", "PoolSet", "pool_search", "recall_at_k", "load_embedding_pickles", "OutfitXConfig", "TransformerConfig", "ItemEncoderConfig",
           "OutfitCompatibilityPredictionTask", "OutfitComplementaryItemRetrievalTask",
           "OutfitFillInTheBlankTask", "OutfitPrecomputeEmbeddingTask", "FashionItem"]


def __getattr__(name):  # torch is imported lazily so that `import outfitx_b200.synth` stays light
    if name in ("OutfitX", "aggregate_embeddings"):
        from . import model
        return getattr(model, name)
    if name in ("Gallery", "cir_search", "local_search", "merge_lists", "ShardedSearch", "shard_rows",
                "PoolSet", "pool_search", "recall_at_k", "load_embedding_pickles"):
        from . import search
        return getattr(search, name)
    raise AttributeError(name)
