"""Mirror of the reference's `scOT` package surface for the hot path: `scOT.model`."""
from .model import ConditionalLayerNorm, LayerNorm, ScOT, ScOTConfig, ScOTOutput  # noqa: F401
