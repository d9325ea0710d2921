"""Scalar formulas of the CUDA kernels restated in numpy float32 and checked on the CPU against exact references.

* erf-GELU of the GEMM epilogues (`gelu_parts` in csrc/common.cuh: Abramowitz-Stegun 7.1.26 erfc, one ex2 and one rcp):
  the approximation error must stay far below the bf16 precision the results are stored in (reference: ACT2FN["gelu"] =
  exact erf GELU, HF modeling_swinv2.py:571-580).
* log-spaced relative coordinates of the continuous position bias (`cpb_coord` in csrc/attention.cu, HF:491-506).
"""
import math

import numpy as np
import torch


def gelu_parts_f32(x):
    """literal float32 transcription of gelu_parts(): returns (cdf, pdf)"""
    x = x.astype(np.float32)
    ax = np.abs(x) * np.float32(0.70710678118654752440)
    e2 = np.exp2((x * x) * np.float32(-0.72134752044448170368)).astype(np.float32)
    t = (np.float32(1.0) / (np.float32(0.3275911) * ax + np.float32(1.0))).astype(np.float32)
    poly = np.float32(1.061405429) * t + np.float32(-1.453152027)
    poly = poly * t + np.float32(1.421413741)
    poly = poly * t + np.float32(-0.284496736)
    poly = poly * t + np.float32(0.254829592)
    half_erfc = np.float32(0.5) * (poly * t * e2)
    cdf = np.where(x >= 0, np.float32(1.0) - half_erfc, half_erfc).astype(np.float32)
    pdf = (np.float32(0.39894228040143267794) * e2).astype(np.float32)
    return cdf, pdf


def test_gelu_epilogue_formula_matches_exact_erf_gelu():
    x = np.concatenate([np.linspace(-12, 12, 200001), np.random.default_rng(0).normal(0, 2, 100000)]).astype(np.float32)
    cdf, pdf = gelu_parts_f32(x)
    xt = torch.from_numpy(x).double()
    phi_exact = 0.5 * (1 + torch.erf(xt / math.sqrt(2)))
    gelu_exact = (xt * phi_exact).numpy()
    grad_exact = (phi_exact + xt * torch.exp(-0.5 * xt * xt) / math.sqrt(2 * math.pi)).numpy()
    gelu = x * cdf
    grad = x * pdf + cdf
    assert np.max(np.abs(cdf - phi_exact.numpy())) < 5e-7          # A&S 7.1.26: 0.75e-7 on Phi, the rest is fp32 rounding
    assert np.max(np.abs(gelu - gelu_exact)) < 1.5e-6
    assert np.max(np.abs(grad - grad_exact)) < 1.5e-6
    # against what is stored: the worst approximation error is < 1/1000 of a bf16 ulp at |y| ~ 1
    assert np.max(np.abs(gelu - gelu_exact)) < 2.0 ** -9 / 1000


def cpb_coord_f32(i, ws):
    x = np.float32(i)
    if ws > 1:
        x = x / np.float32(ws - 1)
    x = x * np.float32(8.0)
    s = np.float32(1.0) if x > 0 else (np.float32(-1.0) if x < 0 else np.float32(0.0))
    return s * np.log2(np.abs(x) + np.float32(1.0)) / np.float32(3.0)


def test_cpb_coords_match_hf_table():
    """HF create_coords_table_and_index: coords / (ws-1) * 8 ; sign * log2(|x| + 1) / log2(8)"""
    for ws in (4, 8, 16):
        rel = torch.arange(-(ws - 1), ws, dtype=torch.float32)
        table = rel / (ws - 1) * 8
        ref = torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / math.log2(8)
        got = np.array([cpb_coord_f32(i, ws) for i in range(-(ws - 1), ws)], dtype=np.float32)
        assert np.max(np.abs(got - ref.numpy())) < 1e-6
