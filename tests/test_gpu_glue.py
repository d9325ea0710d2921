"""Glue ops of the hot path through their own C-ABI entry points (GPU): the index-only ones (patch im2col, merge gather /
scatter, pixel unshuffle, unmerge permutation, the attention token map) are EXACT permutations and are held to
torch.equal against the reference's tensor ops; the arithmetic ones (depthwise 7x7, 5x5 mixing conv, layer scale, loss)
to fp32 round-off against torch.

Reference ops: scOT/model.py:295-310 (embed), :694-704 (merge), :748-754 (unmerge), :198-217 (ConvNeXt), :639-647
(recovery), :1422-1484 (mask + loss); HF modeling_swinv2.py:146-166 (window partition / reverse)."""
import ctypes

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


def bf16_exact(*shape):
    """random values that are exactly representable in bf16 (so a cast is the identity and permutations compare equal)"""
    return torch.randn(*shape, device=dev).bfloat16().float()


def test_embed_im2col_is_the_exact_patch_gather(L):
    B, Cin, H, ps = 3, 5, 32, 4
    x = bf16_exact(B, Cin, H, H)
    out = torch.empty(B * (H // ps) ** 2, Cin * ps * ps, device=dev, dtype=torch.bfloat16)
    L.glue("scot_embed_im2col", x, out, B, Cin, H, H, ps)
    ref = F.unfold(x, kernel_size=ps, stride=ps).transpose(1, 2).reshape(-1, Cin * ps * ps)  # k = (c, di, dj)
    assert torch.equal(out.float(), ref)
    # and the GEMM on it equals the reference Conv2d
    w = bf16_exact(16, Cin, ps, ps)
    y = torch.empty(out.shape[0], 16, device=dev)
    L.gemm(out, w.reshape(16, -1).bfloat16().contiguous(), out.shape[0], 16, Cin * ps * ps, mode=L.EPI_F32, out0=y)
    yr = F.conv2d(x, w, stride=ps).flatten(2).transpose(1, 2).reshape(-1, 16)
    assert rel(y, yr) < 1e-5


@pytest.mark.parametrize("with_inp", [False, True])
def test_merge_gather_scatter_exact(L, with_inp):
    B, res, C = 2, 8, 32
    x = bf16_exact(B * res * res, C)
    inp = torch.zeros_like(x) if with_inp else None  # exactness needs a sum that is bf16-representable
    out = torch.empty(B * (res // 2) ** 2, 4 * C, device=dev, dtype=torch.bfloat16)
    L.glue("scot_merge_gather", x, inp, out, B, res, C)
    xi = x.view(B, res, res, C)
    ref = torch.cat([xi[:, 0::2, 0::2], xi[:, 1::2, 0::2], xi[:, 0::2, 1::2], xi[:, 1::2, 1::2]], -1).reshape(-1, 4 * C)
    assert torch.equal(out.float(), ref)  # order (0,0),(1,0),(0,1),(1,1): scOT/model.py:694-704
    # scatter = exact inverse permutation (+ optional accumulate)
    dG = torch.randn(B * (res // 2) ** 2, 4 * C, device=dev)
    g_in = torch.randn(B * res * res, C, device=dev)
    g_out = torch.empty_like(g_in)
    L.glue("scot_merge_scatter", dG, None, g_out, B, res, C)
    d = dG.view(B, res // 2, res // 2, 4, C)
    refg = torch.empty(B, res, res, C, device=dev)
    refg[:, 0::2, 0::2], refg[:, 1::2, 0::2], refg[:, 0::2, 1::2], refg[:, 1::2, 1::2] = d[..., 0, :], d[..., 1, :], d[..., 2, :], d[..., 3, :]
    assert torch.equal(g_out, refg.view(-1, C))
    L.glue("scot_merge_scatter", dG, g_in, g_out, B, res, C)
    assert torch.equal(g_out, refg.view(-1, C) + g_in)


def test_recovery_unshuffle_exact_and_conv5(L):
    B, OC, H, ps = 2, 5, 32, 4
    D = torch.randn(B * (H // ps) ** 2, OC * ps * ps, device=dev)
    P = torch.empty(B, OC, H, H, device=dev)
    L.glue("scot_recovery_unshuffle", D, P, B, OC, H, H, ps)
    ref = D.view(B, H // ps, H // ps, OC, ps, ps).permute(0, 3, 1, 4, 2, 5).reshape(B, OC, H, H)
    assert torch.equal(P, ref)  # ConvTranspose2d(k = s = ps) output layout, model.py:645
    w = torch.randn(OC, OC, 5, 5, device=dev) * 0.2
    labels = torch.randn(B, OC, H, H, device=dev)
    resid = torch.randn(B, OC, H, H, device=dev)
    pred = torch.empty_like(P)
    L.glue("scot_recovery_conv5_fwd", P, w, None, OC, None, None, 0, pred, B, OC, H, H)
    refc = F.conv2d(P, w, padding=2)
    assert rel(pred, refc) < 1e-5
    # learn_residual input + per-channel mask (mode 1) and per-pixel mask (mode 2, Airfoil: compressible.py:46-53)
    m1 = torch.zeros(B, OC, dtype=torch.uint8, device=dev)
    m1[:, 3] = 1
    L.glue("scot_recovery_conv5_fwd", P, w, resid, OC, labels, m1, 1, pred, B, OC, H, H)
    r1 = refc + resid
    r1[m1.bool()] = labels[m1.bool()]
    assert rel(pred, r1) < 1e-5 and torch.equal(pred[:, 3], labels[:, 3])
    m2 = (torch.rand(B, 1, H, H, device=dev) < 0.3).expand(B, OC, H, H).contiguous()
    L.glue("scot_recovery_conv5_fwd", P, w, None, OC, labels, m2.to(torch.uint8), 2, pred, B, OC, H, H)
    r2 = refc.clone()
    r2[m2] = labels[m2]
    assert rel(pred, r2) < 1e-5 and torch.equal(pred[m2], labels[m2])
    # backward: data gradient (token-major bf16), mixing-weight gradient, projection-bias gradient
    dpred = torch.randn(B, OC, H, H, device=dev)
    Pr = P.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    F.conv2d(Pr, wr, padding=2).backward(dpred)
    scratch = torch.empty_like(P)
    dD = torch.empty(D.shape, device=dev, dtype=torch.bfloat16)
    g_w, g_b = torch.zeros_like(w), torch.zeros(OC, device=dev)
    L.glue("scot_recovery_conv5_bwd", P, w, dpred, scratch, dD, g_w, g_b, B, OC, H, H, ps)
    dD_ref = Pr.grad.view(B, OC, H // ps, ps, H // ps, ps).permute(0, 2, 4, 1, 3, 5).reshape(D.shape)
    assert rel(dD.float(), dD_ref) < 4e-3  # bf16 storage
    assert rel(g_w, wr.grad) < 1e-4 and rel(g_b, dD.float().view(-1, OC, ps * ps).sum((0, 2))) < 1e-4


@pytest.mark.parametrize("B,res,C", [(2, 16, 32), (1, 12, 96), (1, 8, 192), (1, 4, 384)])
def test_convnext_dwconv7_and_layer_scale(L, B, res, C):
    """(1, 12, 96): image width not a multiple of the 8-pixel strip; (1, 8, 192): 24-quad channel chunks leave idle threads;
    (1, 4, 384): three channel chunks per pixel group, image smaller than the filter"""
    torch.manual_seed(C)
    x = torch.randn(B, res, res, C, device=dev)
    w = torch.randn(C, 1, 7, 7, device=dev) * 0.1
    b = torch.randn(C, device=dev)
    out = torch.empty_like(x)
    L.glue("scot_convnext_dwconv7_fwd", x, w, b, out, B, res, C)
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, b, padding=3, groups=C)
    assert rel(out, ref.permute(0, 2, 3, 1)) < 1e-5
    dout = torch.randn_like(x)
    ref.backward(dout.permute(0, 3, 1, 2))
    g_in = torch.randn_like(x)
    g_out = torch.empty_like(x)
    g_w = torch.zeros_like(w)
    L.glue("scot_convnext_dwconv7_bwd", x, w, dout, g_in, g_out, g_w, B, res, C)
    assert rel(g_out - g_in, xr.grad.permute(0, 2, 3, 1)) < 1e-4
    assert rel(g_w, wr.grad) < 1e-4
    rows = B * res * res
    z = torch.randn(rows, C, device=dev)
    gamma = torch.randn(C, device=dev)
    o = torch.empty(rows, C, device=dev)
    zb = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
    L.glue("scot_convnext_scale_add_fwd", x.view(rows, C), z, gamma, o, zb, rows, C)
    assert rel(o, x.view(rows, C) + gamma * z) < 1e-6 and torch.equal(zb, z.bfloat16())
    g = torch.randn(rows, C, device=dev)
    dz = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
    gg, gb = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    L.glue("scot_convnext_scale_add_bwd", g, zb, gamma, dz, gg, gb, rows, C)
    assert torch.equal(dz, (gamma * g).bfloat16())
    assert rel(gg, (g * zb.float()).sum(0)) < 1e-4 and rel(gb, dz.float().sum(0)) < 1e-4


@pytest.mark.parametrize("p,slices", [(1, [0, 1, 3, 4, 5]), (2, None), (1, None)])
def test_loss_forward_backward(L, p, slices):
    """relative L1 over channel groups / plain l1 / plain mse (scOT/model.py:1425-1484), incl. p=1 without slices"""
    B, OC, H = 3, 5, 32
    pred = torch.randn(B, OC, H, H, device=dev)
    labels = torch.randn(B, OC, H, H, device=dev)
    sums = torch.zeros(64, device=dev)
    loss = torch.zeros(1, device=dev)
    sl = (ctypes.c_int * len(slices))(*slices) if slices else None
    n = len(slices) if slices else 0
    lib = L.load()
    L.check(lib.scot_loss_fwd(L.ptr(pred), L.ptr(labels), L.ptr(sums), L.ptr(loss), sl, n, p, B, OC, H * H, L.cur_stream()))
    pr = pred.double().requires_grad_(True)
    import types

    ref = O.scot_loss(pr, labels.double(), types.SimpleNamespace(p=p, channel_slice_list_normalized_loss=slices))
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    ref.backward()
    dpred = torch.empty_like(pred)
    gs = torch.ones(1, device=dev)
    L.check(lib.scot_loss_bwd(L.ptr(pred), L.ptr(labels), L.ptr(sums), L.ptr(gs), None, None, 0, L.ptr(dpred), sl, n, p, B, OC,
                              H * H, L.cur_stream()))
    assert rel(dpred, pr.grad) < 1e-5


def test_attention_token_map_is_exact(L):
    """Window partition, cyclic shift and window reverse are pure index maps inside the attention kernels: with
    k = 0 (uniform scores), a zero bias table and v = one-hot token ids every output row is the exact average of the
    one-hot rows of its window / mask region, i.e. it identifies exactly which tokens were grouped with which."""
    for (Bn, res, ws, shift) in [(2, 32, 16, 8), (2, 32, 16, 0), (1, 16, 8, 4), (2, 8, 8, 0)]:
        heads, hd = 1, 32
        C = heads * hd
        M = Bn * res * res
        code = torch.arange(M, device=dev)
        v = torch.stack([((code >> k) & 1).float() for k in range(hd)], 1)  # bit pattern of the token id (ids < 2^32)
        qkv = torch.zeros(M, 3 * C, device=dev)
        qkv[:, :C] = 1.0
        qkv[:, 2 * C:] = v
        tab2 = torch.zeros((2 * ws - 1) ** 2, heads, device=dev)
        alpha = torch.ones(heads, device=dev)
        out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(Bn * (res // ws) ** 2 * heads, ws * ws, device=dev)
        L.attn_fwd(qkv.bfloat16(), out, lse, tab2, alpha, Bn, res, ws, shift, heads, hd)
        # reference grouping with the reference's own tensor ops
        x = v.view(Bn, res, res, hd)
        if shift:
            x = torch.roll(x, (-shift, -shift), (1, 2))
        xw = O.window_partition(x, ws).view(-1, ws * ws, hd)
        mask = O.shift_attn_mask(res, ws, shift, torch.float32)
        if mask is None:
            att = torch.full((xw.shape[0], ws * ws, ws * ws), 1.0 / (ws * ws), device=dev)
        else:
            m = (mask.to(dev) == 0).float()
            m = m.repeat(Bn, 1, 1)
            att = m / m.sum(-1, keepdim=True)
        o = O.window_reverse((att @ xw).view(-1, ws, ws, hd), ws, res, res)
        if shift:
            o = torch.roll(o, (shift, shift), (1, 2))
        # averages of {0,1} bits over 16 .. 256 tokens: multiples of 1/256, exactly representable; the kernel's P is
        # exp2(0) / count in fp32 -> bf16 products are exact up to the final bf16 rounding of the mean
        assert torch.equal(out.float(), o.reshape(M, hd).bfloat16().float()), (Bn, res, ws, shift)


def test_standalone_conditional_layer_norm_module():
    """the drop-in ConditionalLayerNorm class (scOT/model.py:143-160) evaluated stand-alone runs the CUDA kernels"""
    from poseidon_b200.scOT.model import ConditionalLayerNorm

    torch.manual_seed(0)
    m = ConditionalLayerNorm(96).cuda()
    with torch.no_grad():
        m.weight.bias.fill_(1.0)
        m.weight.weight.normal_(0, 0.2)
        m.bias.weight.normal_(0, 0.2)
    x = torch.randn(3, 8, 8, 96, device=dev, requires_grad=True)
    t = torch.rand(3, device=dev)
    y = m(x, t)
    mean = x.mean(-1, keepdim=True)
    var = (x ** 2).mean(-1, keepdim=True) - mean ** 2
    xn = (x - mean) / (var + 1e-5).sqrt()
    tt = t.reshape(-1, 1)
    ref = m.weight(tt).view(3, 1, 1, 96) * xn + m.bias(tt).view(3, 1, 1, 96)
    assert rel(y, ref) < 1e-5
    g = torch.randn_like(y)
    gx, gw = torch.autograd.grad(ref, [x, m.weight.weight], g, retain_graph=True)
    y.backward(g)
    assert rel(x.grad, gx) < 6e-3 and rel(m.weight.weight.grad, gw) < 6e-3  # zhat is saved in bf16
    with pytest.raises(RuntimeError):
        m.cpu()(x.detach().cpu(), t.cpu())


LAYER_KEYS = ["attention.self.logit_scale", "attention.self.continuous_position_bias_mlp.0.weight",
              "attention.self.continuous_position_bias_mlp.0.bias", "attention.self.continuous_position_bias_mlp.2.weight",
              "attention.self.query.weight", "attention.self.query.bias", "attention.self.key.weight",
              "attention.self.value.weight", "attention.self.value.bias", "attention.output.dense.weight",
              "attention.output.dense.bias", "layernorm_before.weight.weight", "layernorm_before.weight.bias",
              "layernorm_before.bias.weight", "layernorm_before.bias.bias", "intermediate.dense.weight",
              "intermediate.dense.bias", "output.dense.weight", "output.dense.bias", "layernorm_after.weight.weight",
              "layernorm_after.weight.bias", "layernorm_after.bias.weight", "layernorm_after.bias.bias"]


@pytest.mark.parametrize("geom", [(2, 32, 96, 3, 16, 8), (3, 16, 64, 4, 8, 0), (2, 8, 64, 2, 4, 2)])
def test_standalone_layer_forward_backward(L, geom):
    """scot_layer_fwd / scot_layer_bwd (one ScOTLayer, scOT/model.py:500-581) against the fp64 oracle layer, parity precision"""
    from oracle.weights import make_weight

    Bn, res, C, heads, ws, shift = geom
    hidden = 4 * C
    shapes = {"attention.self.logit_scale": (heads, 1, 1), "attention.self.continuous_position_bias_mlp.0.weight": (512, 2),
              "attention.self.continuous_position_bias_mlp.0.bias": (512,),
              "attention.self.continuous_position_bias_mlp.2.weight": (heads, 512), "attention.self.query.weight": (C, C),
              "attention.self.query.bias": (C,), "attention.self.key.weight": (C, C), "attention.self.value.weight": (C, C),
              "attention.self.value.bias": (C,), "attention.output.dense.weight": (C, C), "attention.output.dense.bias": (C,),
              "intermediate.dense.weight": (hidden, C), "intermediate.dense.bias": (hidden,), "output.dense.weight": (C, hidden),
              "output.dense.bias": (C,)}
    for n in ("layernorm_before", "layernorm_after"):
        shapes.update({f"{n}.weight.weight": (C, 1), f"{n}.weight.bias": (C,), f"{n}.bias.weight": (C, 1), f"{n}.bias.bias": (C,)})
    assert len(LAYER_KEYS) == 23
    keys = LAYER_KEYS[:11] + LAYER_KEYS[11:15] + LAYER_KEYS[15:19] + LAYER_KEYS[19:]
    w = {"blk." + k: make_weight("encoder.layers.0.blocks.0." + k, shapes[k], seed=5) for k in keys}
    params = [w["blk." + k].to(dev).contiguous() for k in keys]
    lay = L.Layer(Bn, res, C, heads, ws, shift, 4.0, True, 1e-5, precision=1)
    assert lay.n == 23  # 9 attention.self + 2 attention.output + 4 + 2 + 2 + 4 tensors
    torch.manual_seed(1)
    x = torch.randn(Bn * res * res, C, device=dev)
    t = torch.rand(Bn, device=dev)
    y = lay.forward(params, x, t)
    wd = {k: v.double().to(dev).requires_grad_(True) for k, v in w.items()}
    xd = x.double().view(Bn, res * res, C).requires_grad_(True)
    ref = O.scot_layer(xd, t.double(), wd, "blk", dict(res=res, ws=ws, heads=heads), shift, 1e-5, True)
    assert rel(y, ref.reshape(-1, C)) < 1e-4
    dy = torch.randn_like(y)
    ref.backward(dy.double().view_as(ref))
    grads = [torch.zeros_like(p) for p in params]
    dx = lay.backward(grads, t, dy)
    assert rel(dx, xd.grad.reshape(-1, C)) < 1e-3
    for k, g in zip(keys, grads):
        r = rel(g, wd["blk." + k].grad)
        assert r < 5e-3, (k, r)
