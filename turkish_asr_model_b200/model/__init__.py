from .conformer import (ConformerBlock, ConformerConvModule, SwiGLUFeedForward, TransposeGroupNorm,  # noqa: F401
                        TurkishASRModel)
from .attention import RelativeMultiHeadAttention, RotaryEmbedding  # noqa: F401
