"""Deterministic "trained-like" weights for parity tests and benchmarks — TEST INFRASTRUCTURE.

The reference's default init (HF `_init_weights`, modeling_swinv2.py:884-902) draws every nn.Linear
~N(0, 0.02) *including* the two Linear(1, C) of each ConditionalLayerNorm, so a fresh model has LN scale
~0.02*t and output rms ~0.01: relative errors on that are meaningless. These weights are
a function of (parameter name, shape, seed) only, so the reference in the authoring container, the
oracle and the CUDA engine on the GPU box all see bit-identical fp32 values without shipping tensors.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Tuple

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def make_weight(name: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    g = _gen(name, seed)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32)
    leaf = name.split(".")
    # ConditionalLayerNorm: <norm>.weight.{weight,bias}, <norm>.bias.{weight,bias}
    if len(leaf) >= 3 and leaf[-2] in ("weight", "bias") and ("norm" in leaf[-3]):
        if leaf[-2] == "weight":
            return (1.0 + 0.1 * rn(*shape)) if leaf[-1] == "bias" else 0.2 * rn(*shape)
        return 0.05 * rn(*shape) if leaf[-1] == "bias" else 0.1 * rn(*shape)
    if "norm" in leaf[-2] and len(shape) == 1:  # plain LayerNorm
        return (1.0 + 0.1 * rn(*shape)) if leaf[-1] == "weight" else 0.05 * rn(*shape)
    if leaf[-1] == "logit_scale":
        return math.log(10.0) + 0.3 * rn(*shape)
    if leaf[0] == "residual_blocks" and len(leaf) == 4 and leaf[-1] == "weight":  # ConvNeXt layer scale
        return 0.5 + 0.1 * rn(*shape)
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        if "patch_recovery.projection" in name:  # ConvTranspose2d: weight [in, out, k, k]
            fan_in = shape[0]
        if "continuous_position_bias_mlp.0" in name:
            return rn(*shape)  # inputs are O(1) log-coordinates
        return rn(*shape) / math.sqrt(fan_in)
    return 0.02 * rn(*shape)  # biases


def make_weights(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: make_weight(k, tuple(v), seed) for k, v in shapes.items()}


def make_inputs(batch: int, cin: int, cout: int, size: int, seed: int = 0, mask_channels=()):
    """Synthetic z-normalised PDE-like batch: N(0,1) fields, time ~ U(0,1)."""
    g = _gen("inputs", seed)
    x = torch.randn(batch, cin, size, size, generator=g)
    y = torch.randn(batch, cout, size, size, generator=g)
    t = torch.rand(batch, generator=g)
    pm = torch.zeros(batch, cout, dtype=torch.bool)
    for c in mask_channels:
        pm[:, c] = True
    return x, t, y, pm
