"""Module path of the reference (flaxdiff/inputs/encoders.py)."""
from . import (CLIPTextEncoder, CONDITIONAL_ENCODERS_REGISTRY, ConditioningEncoder,  # noqa: F401
               RandomEmbeddingEncoder)

TextEncoder = ConditioningEncoder
