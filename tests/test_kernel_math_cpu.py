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


def gelu_and_grad_f32(x):
    """literal float32 transcription of gelu_and_grad() (csrc/common.cuh): the folded-constant sequence of the async GEMM
    epilogue; returns (gelu, gelu')"""
    x = x.astype(np.float32)
    e2 = np.exp2((x * x) * np.float32(-0.72134752044448170368)).astype(np.float32)
    k = np.float32(np.float32(0.3275911) * np.float32(0.70710678118654752440))
    t = (np.float32(1.0) / (k * np.abs(x) + np.float32(1.0))).astype(np.float32)
    h = np.float32(0.5)
    poly = (h * np.float32(1.061405429)) * t + (h * np.float32(-1.453152027))
    poly = poly * t + (h * np.float32(1.421413741))
    poly = poly * t + (h * np.float32(-0.284496736))
    poly = poly * t + (h * np.float32(0.254829592))
    half_erfc = ((poly * t) * e2).astype(np.float32)
    cdf = np.where(x >= 0, np.float32(1.0) - half_erfc, half_erfc).astype(np.float32)
    g = (x * cdf).astype(np.float32)
    dg = ((x * np.float32(0.39894228040143267794)) * e2 + cdf).astype(np.float32)
    return g, dg


def test_folded_gelu_sequence_matches_exact_erf_gelu_and_its_derivative():
    x = np.linspace(-12, 12, 200001).astype(np.float32)
    g, dg = gelu_and_grad_f32(x)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    ref = torch.nn.functional.gelu(xt)
    ref.sum().backward()
    # absolute errors far below the bf16 resolution (2^-9 relative) the two outputs are stored in
    assert np.abs(g - ref.detach().numpy()).max() < 2e-6 * 12
    assert np.abs(dg - xt.grad.numpy()).max() < 5e-6
    # and the same function as the unfolded form used by the other epilogues
    cdf, pdf = gelu_parts_f32(x)
    assert np.abs(g - x * cdf).max() < 2e-6 and np.abs(dg - (x * pdf + cdf)).max() < 2e-6


def test_epilogue_slab_swizzles_are_conflict_free_and_match_the_tma_box_layout():
    """Staging slabs of the async GEMM epilogue (csrc/gemm.cu): thread `lane` owns row `lane` of a 32-row slab and writes its
    16-byte chunks at chunk ^ f(row). (1) The address must be what the TMA unit expects for the box's swizzle mode:
    SWIZZLE_64B (64-byte rows) XORs address bits [4:5] with bits [7:8]; SWIZZLE_128B (128-byte rows) XORs bits [4:6] with
    bits [7:9]. (2) The eight lanes of a quarter-warp (one 128-bit store phase) must hit eight distinct 16-byte bank groups."""
    for row_bytes, nchunk, f in ((64, 4, lambda r: (r >> 1) & 3), (128, 8, lambda r: r & 7)):
        for lane in range(32):
            for j in range(nchunk):
                addr = lane * row_bytes + ((j ^ f(lane)) << 4)
                linear = lane * row_bytes + (j << 4)  # where the element sits in the unswizzled box
                if row_bytes == 64:
                    expect = linear ^ (((linear >> 7) & 3) << 4)
                else:
                    expect = linear ^ (((linear >> 7) & 7) << 4)
                assert addr == expect, (row_bytes, lane, j)
        for j in range(nchunk):
            for phase in range(4):
                groups = {((l * row_bytes + ((j ^ f(l)) << 4)) >> 4) & 7 for l in range(8 * phase, 8 * phase + 8)}
                assert len(groups) == 8, (row_bytes, j, phase)
    # bias-gradient column sums read the bf16 slab back: lane = (column pair p, row parity), 16 rows each; the two parities
    # must fall into different 16-bank halves for every step (64-byte rows: even rows start at bank 0, odd rows at bank 16)
    for i in range(16):
        banks = []
        for lane in range(32):
            p, par = lane & 15, lane >> 4
            r = 2 * i + par
            addr = r * 64 + (((p >> 2) ^ ((r >> 1) & 3)) << 4) + (p & 3) * 4
            banks.append((addr >> 2) & 31)
        assert len(set(banks)) == 32, i


def rows_kernel_rpb(rows, T, conditioned, sweep, unit, target):
    """Python mirror of rows_kernel_rpb() in csrc/norm.cu"""
    rpb = -(-rows // target)
    rpb = -(-rpb // sweep) * sweep
    if not conditioned:
        return rpb
    if T <= 0 or rows % T:
        return 0
    best = 0
    d = sweep
    while d <= T:
        if T % d == 0:
            best = d
            if d >= rpb:
                break
        d += sweep
    if best == 0:
        d = unit
        while d <= T:
            if T % d == 0:
                best = d
            d += unit
    return best


def test_layernorm_block_sizing_never_lets_a_block_span_two_samples():
    """one-sample-per-block LayerNorm kernels: for a conditioned norm the rows per block must divide the rows per sample
    (else the generic kernels run: 0); the Poseidon-B stage shapes must get about one resident wave"""
    sms = 148
    for lpr, v in ((4, 3), (8, 3), (16, 3), (32, 3), (32, 6), (8, 1), (32, 2)):
        rpw, r = 32 // lpr, (2 if v <= 3 else 1)
        sweep, unit, target = 8 * rpw * r, 8 * rpw, sms * (2 if v <= 3 else 1)
        for T in (1, 4, 16, 20, 64, 100, 256, 576, 1024, 4096):
            for B in (1, 3, 8, 64):
                rpb = rows_kernel_rpb(B * T, T, True, sweep, unit, target)
                assert rpb == 0 or (T % rpb == 0 and rpb % unit == 0), (lpr, v, T, B, rpb)
                assert rows_kernel_rpb(B * T, T, False, sweep, unit, target) % sweep == 0
    # Poseidon-B, batch 64: (rows, C) -> (LPR, V) as dispatched in norm.cu; 256 / 256 / 256 / 128 blocks
    for rows, T, lpr, v, blocks in ((65536, 1024, 8, 3, 256), (16384, 256, 16, 3, 256), (4096, 64, 32, 3, 256), (1024, 16, 32, 6, 128)):
        rpw, r = 32 // lpr, (2 if v <= 3 else 1)
        rpb = rows_kernel_rpb(rows, T, True, 8 * rpw * r, 8 * rpw, sms * (2 if v <= 3 else 1))
        assert rows // rpb == blocks, (rows, rpb)
