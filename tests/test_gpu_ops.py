"""Per-kernel parity tests (GPU): every C-ABI op against plain torch on identical (bf16-rounded) inputs.

The torch reference is evaluated in fp32/fp64 on the same quantised inputs, so the only difference is the
kernel's internal rounding / accumulation order; index maps (window partition, cyclic shift, pixel shuffle,
relative position index) are exercised through shapes where any permutation error gives O(1) mismatch.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import scot_oracle as O

pytestmark = pytest.mark.gpu

dev = "cuda"


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture(scope="module")
def L():
    from poseidon_b200 import _lib

    _lib.load()
    return _lib


# ------------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1000, 64, 48), (4096, 288, 96), (1024, 2304, 768), (512, 80, 96)])
def test_gemm_forward_bias(L, impl, M, N, K):
    torch.manual_seed(1)
    A = torch.randn(M, K, device=dev).bfloat16()
    B = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev)
    L.gemm(A, B, M, N, K, mode=L.EPI_F32, bias=bias, out0=out, impl=impl)
    ref = A.float() @ B.float().t() + bias
    assert rel(out, ref) < 2e-5


@pytest.mark.parametrize("impl", [0, 1])
def test_gemm_dgrad_wgrad(L, impl):
    torch.manual_seed(2)
    M, N, K = 4096, 96, 288
    dY = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(K, N, device=dev) / K ** 0.5).bfloat16()
    g = torch.randn(M, N, device=dev)
    g0 = g.clone()
    L.gemm(dY, W, M, N, K, b_mn=True, mode=L.EPI_RMW_F32, out0=g, impl=impl)
    assert rel(g, g0 + dY.float() @ W.float()) < 2e-5
    X = torch.randn(M, N, device=dev).bfloat16()
    dW = torch.zeros(K, N, device=dev)
    L.gemm(dY, X, K, N, M, a_mn=True, b_mn=True, mode=L.EPI_ATOMIC_F32, out0=dW, impl=impl)
    assert rel(dW, dY.float().t() @ X.float()) < 2e-5


@pytest.mark.parametrize("impl", [0, 1])
def test_gemm_gelu_and_backward(L, impl):
    torch.manual_seed(3)
    M, N, K = 2048, 384, 96
    A = torch.randn(M, K, device=dev).bfloat16()
    B = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    gp = torch.empty(M, N, device=dev, dtype=torch.bfloat16)  # gelu'(h), saved for backward
    g = torch.empty_like(gp)
    L.gemm(A, B, M, N, K, mode=L.EPI_GELU, bias=bias, out0=gp, out1=g, impl=impl)
    href = (A.float() @ B.float().t() + bias).requires_grad_(True)
    gref = F.gelu(href)
    gref.sum().backward()
    assert rel(g.float(), gref) < 4e-3
    assert rel(gp.float(), href.grad) < 4e-3
    dY = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(K, N, device=dev) / K ** 0.5).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    cs = torch.zeros(N, device=dev)
    L.gemm(dY, W, M, N, K, b_mn=True, mode=L.EPI_GELU_BWD, out0=out, aux=gp, colsum=cs, impl=impl)
    assert rel(out.float(), (dY.float() @ W.float()) * gp.float()) < 4e-3
    assert rel(cs, out.float().sum(0)) < 1e-4


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("tok,C", [(4096, 96), (1024, 768), (512, 48)])
def test_gemm_wgrad_group(L, impl, tok, C):
    """the four weight gradients of a transformer block in one launch (mixed tile widths fall back to four)"""
    torch.manual_seed(tok + C)
    H = 4 * C
    mk = lambda n: torch.randn(tok, n, device=dev).bfloat16()
    probs = [(mk(C), mk(H), torch.zeros(C, H, device=dev)), (mk(H), mk(C), torch.zeros(H, C, device=dev)),
             (mk(C), mk(C), torch.zeros(C, C, device=dev)), (mk(3 * C), mk(C), torch.zeros(3 * C, C, device=dev))]
    L.wgrad_group(probs, impl=impl)
    L.wgrad_group(probs[:2], impl=impl)  # accumulates
    for k, (dY, X, dW) in enumerate(probs):
        ref = dY.float().t() @ X.float() * (2 if k < 2 else 1)
        assert rel(dW, ref) < 2e-5, k


# ------------------------------------------------------------------------------------------------------
# (conditional) layer norm
# ------------------------------------------------------------------------------------------------------
def ref_cln(z, t, aw, ab, cw, cb, eps, T):
    mean = z.mean(-1, keepdim=True)
    var = (z ** 2).mean(-1, keepdim=True) - mean ** 2
    zh = (z - mean) / (var + eps).sqrt()
    tt = t.repeat_interleave(T).unsqueeze(1) if t is not None else 0.0
    sc = ab + (aw * tt if aw is not None else 0.0)
    sh = cb + (cw * tt if cw is not None else 0.0)
    return sc * zh + sh


@pytest.mark.parametrize("C", [16, 48, 96, 192, 384, 768])
@pytest.mark.parametrize("cond,Bn,T", [(True, 3, 64), (False, 3, 64), (True, 5, 20), (True, 2, 256)])
def test_cln_forward_backward(L, C, cond, Bn, T):
    """T = 64 / 256: one-sample-per-block kernels (rows per block divides T); T = 20 with a conditioned norm: no block size
    divides the sample, the generic kernels (per-row lead time, blocks spanning samples) run"""
    torch.manual_seed(C)
    rows = Bn * T
    z = torch.randn(rows, C, device=dev, dtype=torch.float64) * 2 + 0.5
    res = torch.randn(rows, C, device=dev, dtype=torch.float64)
    t = torch.rand(Bn, device=dev, dtype=torch.float64) if cond else None
    aw = torch.randn(C, device=dev, dtype=torch.float64) * 0.2 if cond else None
    cw = torch.randn(C, device=dev, dtype=torch.float64) * 0.2 if cond else None
    ab = 1 + 0.1 * torch.randn(C, device=dev, dtype=torch.float64)
    cb = 0.1 * torch.randn(C, device=dev, dtype=torch.float64)
    leaves = [v.requires_grad_(True) for v in ([z, ab, cb] + ([aw, cw] if cond else []))]
    y_ref = ref_cln(z, t, aw, ab, cw, cb, 1e-5, T) + res
    f = lambda v: None if v is None else v.detach().float().contiguous()
    x_out = torch.empty(rows, C, device=dev)
    xb = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
    zhat = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
    rstd = torch.empty(rows, device=dev)
    L.cln_fwd(f(z), f(res), f(t), f(aw), f(ab), f(cw), f(cb), x_out, xb, zhat, rstd, rows, C, T, 0, 1e-5)
    assert rel(x_out, y_ref) < 1e-5
    assert rel(xb.float(), y_ref) < 4e-3
    dy = torch.randn(rows, C, device=dev, dtype=torch.float64)
    y_ref.backward(dy)
    dz = torch.empty(rows, C, device=dev)
    gs = [torch.zeros(C, device=dev) for _ in range(5)]
    L.cln_bwd(f(dy), zhat, rstd, f(t), f(aw), f(ab), dz, True, gs[0] if cond else None, gs[1], gs[2] if cond else None,
              gs[3], gs[4], rows, C, T, 0)
    # zhat is stored in bf16 -> tolerance of a bf16-rounded intermediate
    assert rel(dz, z.grad) < 6e-3
    assert rel(gs[1], ab.grad) < 6e-3 and rel(gs[3], cb.grad) < 1e-5
    if cond:
        assert rel(gs[0], aw.grad) < 6e-3 and rel(gs[2], cw.grad) < 1e-5
    assert rel(gs[4], dz.sum(0)) < 1e-4


def test_cln_unmerge_permutation(L):
    """pixel-shuffle row permutation of ScOTPatchUnmerging (scOT/model.py:748-754) is an exact index map"""
    torch.manual_seed(0)
    Bn, res, Ch = 2, 4, 32  # Ch = C/2 of the coarse stage
    z = torch.randn(Bn * res * res, 4 * Ch, device=dev)
    ab = torch.ones(Ch, device=dev)
    cb = torch.zeros(Ch, device=dev)
    rows = Bn * res * res * 4
    xb = torch.empty(rows, Ch, device=dev, dtype=torch.bfloat16)
    zh = torch.empty(rows, Ch, device=dev, dtype=torch.bfloat16)
    rstd = torch.empty(rows, device=dev)
    x = torch.empty(rows, Ch, device=dev)
    L.cln_fwd(z.view(rows, Ch).contiguous(), None, None, None, ab, None, cb, x, xb, zh, rstd, rows, Ch, 4 * res * res, res, 1e-5)
    ref = z.reshape(Bn, res, res, 2, 2, Ch).permute(0, 1, 3, 2, 4, 5).reshape(Bn, 4 * res * res, Ch)
    ref = F.layer_norm(ref, (Ch,), ab, cb, 1e-5).reshape(rows, Ch)
    assert rel(x, ref) < 1e-5
    # backward un-permutes: with dy = one-hot rows the gradient rows land on the inverse positions
    dy = torch.randn(rows, Ch, device=dev)
    dz = torch.empty(rows, Ch, device=dev)
    g1, g2 = torch.zeros(Ch, device=dev), torch.zeros(Ch, device=dev)
    L.cln_bwd(dy, zh, rstd, None, None, ab, dz, True, None, g1, None, g2, None, rows, Ch, 4 * res * res, res)
    zz = z.clone().requires_grad_(True)
    r2 = zz.reshape(Bn, res, res, 2, 2, Ch).permute(0, 1, 3, 2, 4, 5).reshape(Bn, 4 * res * res, Ch)
    F.layer_norm(r2, (Ch,), ab, cb, 1e-5).reshape(rows, Ch).backward(dy)
    assert rel(dz.view_as(zz), zz.grad) < 6e-3


# ------------------------------------------------------------------------------------------------------
# continuous position bias + window attention
# ------------------------------------------------------------------------------------------------------
def ref_attention(qkv, cpb, ls, Bn, res, ws, shift, heads, hd):
    """torch restatement on token-major qkv [M, 3C] (float64), returns token-major out [M, C]"""
    C = heads * hd
    x = qkv.view(Bn, res, res, 3 * C)
    if shift:
        x = torch.roll(x, (-shift, -shift), (1, 2))
    xw = O.window_partition(x, ws).view(-1, ws * ws, 3 * C)
    bw, n = xw.shape[0], ws * ws
    q, k, v = [t.reshape(bw, n, heads, hd).transpose(1, 2) for t in xw.split(C, dim=-1)]
    attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
    attn = attn * torch.clamp(ls, max=math.log(100.0)).exp()
    w1, b1, w2 = cpb
    coords = O.relative_coords_table(ws).to(qkv)
    table = F.linear(F.relu(F.linear(coords, w1, b1)), w2)
    bias = table[O.relative_position_index(ws).view(-1).to(qkv.device)].view(n, n, heads).permute(2, 0, 1)
    attn = attn + 16 * torch.sigmoid(bias).unsqueeze(0)
    mask = O.shift_attn_mask(res, ws, shift, torch.float64)
    if mask is not None:
        nw = mask.shape[0]
        attn = attn.view(bw // nw, nw, heads, n, n) + 2 * mask.to(qkv).unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, n, n)
    out = (attn.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(bw, ws, ws, C)
    out = O.window_reverse(out, ws, res, res)
    if shift:
        out = torch.roll(out, (shift, shift), (1, 2))
    return out.reshape(-1, C)


ATTN_CASES = [
    # Bn, res, ws, shift, heads, hd
    (2, 32, 16, 8, 3, 32),   # Poseidon-B stage 0, shifted
    (3, 16, 16, 0, 6, 32),   # stage 1
    (3, 8, 8, 0, 12, 32),    # stage 2
    (5, 4, 4, 0, 24, 32),    # stage 3
    (2, 32, 16, 8, 3, 16),   # Poseidon-T stage 0
    (2, 16, 8, 4, 2, 16),    # 8x8 shifted windows (tiny golden model)
    (2, 32, 16, 0, 3, 64),   # Poseidon-L head_dim
]


@pytest.mark.parametrize("case", ATTN_CASES)
def test_window_attention_forward_backward(L, case):
    Bn, res, ws, shift, heads, hd = case
    torch.manual_seed(sum(case))
    C = heads * hd
    M = Bn * res * res
    qkv = (torch.randn(M, 3 * C, device=dev) * 1.5).bfloat16()
    w1 = torch.randn(512, 2, device=dev)
    b1 = torch.randn(512, device=dev) * 0.1
    w2 = torch.randn(heads, 512, device=dev) / 512 ** 0.5
    ls = math.log(10.0) + 0.3 * torch.randn(heads, 1, 1, device=dev)
    R = (2 * ws - 1) ** 2
    cpb = L.CpbLayerBuffers(w1, b1, w2, ls, ws, heads)
    cpb.forward()
    tab2, alpha = cpb.tab2, cpb.alpha
    coords = O.relative_coords_table(ws).to(dev)
    tab_ref = 16 * torch.sigmoid(F.linear(F.relu(F.linear(coords, w1, b1)), w2)) * math.log2(math.e)
    assert rel(tab2, tab_ref) < 1e-5
    assert rel(alpha, torch.clamp(ls, max=math.log(100.0)).exp().view(-1)) < 1e-6

    nwin = Bn * (res // ws) ** 2
    out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(nwin * heads, ws * ws, device=dev)
    L.attn_fwd(qkv, out, lse, tab2, alpha, Bn, res, ws, shift, heads, hd)
    leaves = [t.double().requires_grad_(True) for t in (qkv, w1, b1, w2, ls)]
    ref = ref_attention(leaves[0], leaves[1:4], leaves[4], Bn, res, ws, shift, heads, hd)
    assert rel(out.float(), ref) < 1e-2, "attention forward"

    d_o = torch.randn(M, C, device=dev).bfloat16()
    ref.backward(d_o.double())
    dqkv = torch.zeros(M, 3 * C, device=dev, dtype=torch.bfloat16)
    import ctypes
    pbytes = L.load().scot_attn_bwd_partial_bytes(ws, heads, nwin)
    partial = torch.zeros(pbytes // 4, device=dev)  # accumulation buffer: zero on entry, zero again on return
    dtab, dalpha = cpb.dtab, cpb.dalpha
    gq = torch.zeros(C, device=dev)
    gv = torch.zeros(C, device=dev)
    L.attn_bwd(qkv, out, d_o, lse, tab2, alpha, dqkv, partial, dtab, dalpha, gq, gv, Bn, res, ws, shift, heads, hd)
    g_ref = leaves[0].grad
    assert rel(dqkv[:, 2 * C:].float(), g_ref[:, 2 * C:]) < 2e-2, "dv"
    assert rel(dqkv[:, :C].float(), g_ref[:, :C]) < 3e-2, "dq"
    assert rel(dqkv[:, C:2 * C].float(), g_ref[:, C:2 * C]) < 3e-2, "dk"
    assert rel(gq, dqkv[:, :C].float().sum(0)) < 1e-3 and rel(gv, dqkv[:, 2 * C:].float().sum(0)) < 1e-3
    assert float(partial.abs().max()) == 0.0
    # bias-table / logit-scale gradients through the two-stage reduction + cpb backward
    cpb.backward()
    g1, gb, g2, gls = cpb.grad(0, w1.shape), cpb.grad(1, b1.shape), cpb.grad(2, w2.shape), cpb.grad(3, (heads,))
    assert rel(g2, leaves[3].grad) < 3e-2, "cpb w2 grad"
    assert rel(g1, leaves[1].grad) < 3e-2, "cpb w1 grad"
    assert rel(gb, leaves[2].grad) < 3e-2, "cpb b1 grad"
    assert rel(gls, leaves[4].grad.view(-1)) < 3e-2, "logit_scale grad"
