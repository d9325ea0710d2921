"""CPU oracle for the scOT hot path — TEST INFRASTRUCTURE ONLY.

A plain-torch, functional restatement of `ScOT.forward` of the reference (camlab-ethz/poseidon,
`scOT/model.py:1318-1509`) and of the HuggingFace swinv2 arithmetic it imports
(`transformers/models/swinv2/modeling_swinv2.py`, "HF" below; transformers 5.5.0 is the version installed
in this image and therefore the behaviour the reference exhibits here). It works in any floating dtype
(fp64 for goldens, fp32 for the timed CPU baseline) and gets gradients from torch autograd.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module, and only as the checker / reported baseline. The product path
(`poseidon_b200.scOT.model`) never imports it and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors, so this restatement is
pinned against the *reference itself* executed in the authoring container: `oracle/make_golden.py`
imports `/root/reference/scOT/model.py` (unmodified, three compatibility shims for transformers 5.5.0)
and stores its outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this file against
those fixtures.

Parameters are addressed by the reference's state_dict names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------------
# geometry helpers (integer / index maps — must be bit exact)
# ------------------------------------------------------------------------------------------------------
def stage_geometry(cfg) -> List[dict]:
    """Per stage: resolution, channels, heads, effective window and shift (scOT/model.py:412-440)."""
    grid = cfg.image_size // cfg.patch_size
    out = []
    for s, depth in enumerate(cfg.depths):
        res = grid // (2 ** s)
        ws = res if res <= cfg.window_size else cfg.window_size  # model.py:428-430
        shift = 0 if res <= ws else cfg.window_size // 2  # model.py:431-440 (tuple compare)
        out.append(dict(res=res, dim=cfg.embed_dim * 2 ** s, heads=cfg.num_heads[s], ws=ws, shift=shift, depth=depth))
    return out


def window_partition(x: Tensor, ws: int) -> Tensor:
    """HF:146-155. [B,H,W,C] -> [B*nW, ws, ws, C]."""
    b, h, w, c = x.shape
    x = x.view(b, h // ws, ws, w // ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, c)


def window_reverse(win: Tensor, ws: int, h: int, w: int) -> Tensor:
    """HF:158-166."""
    c = win.shape[-1]
    x = win.view(-1, h // ws, w // ws, ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, h, w, c)


def shift_attn_mask(res: int, ws: int, shift: int, dtype) -> Optional[Tensor]:
    """scOT/model.py:442-478: [nW, N, N] of {0,-100}."""
    if shift == 0:
        return None
    img = torch.zeros((1, res, res, 1), dtype=dtype)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def relative_coords_table(ws: int) -> Tensor:
    """HF:489-510, computed in fp32 exactly like the non-persistent buffer of the reference."""
    rc = torch.arange(-(ws - 1), ws, dtype=torch.int64).float()
    t = torch.stack(torch.meshgrid([rc, rc], indexing="ij")).permute(1, 2, 0).contiguous().unsqueeze(0)
    if ws > 1:
        t[:, :, :, 0] /= ws - 1
        t[:, :, :, 1] /= ws - 1
    t *= 8
    t = torch.sign(t) * torch.log2(torch.abs(t) + 1.0) / math.log2(8)
    return t.view(-1, 2)  # [(2ws-1)^2, 2]


def relative_position_index(ws: int) -> Tensor:
    """HF:512-523. [N, N] int64."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


# ------------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------------
def cond_layer_norm(x: Tensor, t: Tensor, p: Dict[str, Tensor], pre: str, eps: float, conditioned: bool) -> Tensor:
    """ConditionalLayerNorm (scOT/model.py:143-160) or LayerNorm (:135-140) on the last dim."""
    if not conditioned:
        return F.layer_norm(x, (x.shape[-1],), p[pre + ".weight"], p[pre + ".bias"], eps)
    mean = x.mean(dim=-1, keepdim=True)
    var = (x ** 2).mean(dim=-1, keepdim=True) - mean ** 2
    xn = (x - mean) / (var + eps).sqrt()
    tt = t.reshape(-1, 1).to(x.dtype)
    w = F.linear(tt, p[pre + ".weight.weight"], p[pre + ".weight.bias"])  # [B, C]
    b = F.linear(tt, p[pre + ".bias.weight"], p[pre + ".bias.bias"])
    shape = [x.shape[0]] + [1] * (x.dim() - 2) + [x.shape[-1]]
    return w.view(shape) * xn + b.view(shape)


def window_attention(xw: Tensor, p: Dict[str, Tensor], pre: str, heads: int, ws: int, mask: Optional[Tensor]) -> Tensor:
    """Swinv2Attention (HF:421-487 + :528-538) on windows xw [B*nW, N, C]."""
    bw, n, c = xw.shape
    d = c // heads
    sp = pre + ".self"
    q = F.linear(xw, p[sp + ".query.weight"], p.get(sp + ".query.bias")).view(bw, n, heads, d).transpose(1, 2)
    k = F.linear(xw, p[sp + ".key.weight"]).view(bw, n, heads, d).transpose(1, 2)
    v = F.linear(xw, p[sp + ".value.weight"], p.get(sp + ".value.bias")).view(bw, n, heads, d).transpose(1, 2)
    attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)  # eps 1e-12
    scale = torch.clamp(p[sp + ".logit_scale"], max=math.log(1.0 / 0.01)).exp()
    attn = attn * scale
    coords = relative_coords_table(ws).to(xw)  # same dtype AND device as the activations (the GPU eager baseline runs it on cuda)
    hid = F.relu(F.linear(coords, p[sp + ".continuous_position_bias_mlp.0.weight"],
                          p[sp + ".continuous_position_bias_mlp.0.bias"]))
    table = F.linear(hid, p[sp + ".continuous_position_bias_mlp.2.weight"])  # [(2ws-1)^2, heads]
    idx = relative_position_index(ws).view(-1).to(xw.device)
    bias = table[idx].view(n, n, heads).permute(2, 0, 1).contiguous()
    attn = attn + (16 * torch.sigmoid(bias)).unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        m = mask.to(xw).unsqueeze(1).unsqueeze(0)
        attn = attn.view(bw // nw, nw, heads, n, n) + m
        attn = attn + m  # HF 5.5.0 adds the mask twice (HF:465-468)
        attn = attn.view(-1, heads, n, n)
    attn = F.softmax(attn, dim=-1)
    ctx = (attn @ v).permute(0, 2, 1, 3).contiguous().view(bw, n, c)
    return F.linear(ctx, p[pre + ".output.dense.weight"], p[pre + ".output.dense.bias"])


def scot_layer(x: Tensor, t: Tensor, p, pre: str, g: dict, shift: int, eps: float, cond: bool) -> Tensor:
    """ScOTLayer.forward (scOT/model.py:500-581) on x [B, res*res, C]."""
    b, _, c = x.shape
    res, ws = g["res"], g["ws"]
    h = x.view(b, res, res, c)
    if shift > 0:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
    xw = window_partition(h, ws).view(-1, ws * ws, c)
    mask = shift_attn_mask(res, ws, shift, x.dtype)
    a = window_attention(xw, p, pre + ".attention", g["heads"], ws, mask)
    a = window_reverse(a.view(-1, ws, ws, c), ws, res, res)
    if shift > 0:
        a = torch.roll(a, shifts=(shift, shift), dims=(1, 2))
    a = a.view(b, res * res, c)
    x = x + cond_layer_norm(a, t, p, pre + ".layernorm_before", eps, cond)
    m = F.linear(x, p[pre + ".intermediate.dense.weight"], p[pre + ".intermediate.dense.bias"])
    m = F.gelu(m)
    m = F.linear(m, p[pre + ".output.dense.weight"], p[pre + ".output.dense.bias"])
    return x + cond_layer_norm(m, t, p, pre + ".layernorm_after", eps, cond)


def patch_merging(x: Tensor, t: Tensor, p, pre: str, res: int, cond: bool) -> Tensor:
    """ScOTPatchMerging.forward (scOT/model.py:680-712)."""
    b, _, c = x.shape
    z = x.view(b, res, res, c)
    z = torch.cat([z[:, 0::2, 0::2], z[:, 1::2, 0::2], z[:, 0::2, 1::2], z[:, 1::2, 1::2]], -1).view(b, -1, 4 * c)
    z = F.linear(z, p[pre + ".reduction.weight"])
    return cond_layer_norm(z, t, p, pre + ".norm", 1e-5, cond)


def patch_unmerging(x: Tensor, t: Tensor, p, pre: str, res: int, cond: bool) -> Tensor:
    """ScOTPatchUnmerging.forward (scOT/model.py:737-760)."""
    b, _, c = x.shape
    z = F.linear(x, p[pre + ".upsample.weight"])
    z = z.reshape(b, res, res, 2, 2, c // 2).permute(0, 1, 3, 2, 4, 5).reshape(b, 4 * res * res, c // 2)
    z = cond_layer_norm(z, t, p, pre + ".norm", 1e-5, cond)
    return F.linear(z, p[pre + ".mixup.weight"])


def convnext_block(x: Tensor, t: Tensor, p, pre: str, res: int, eps: float, cond: bool) -> Tensor:
    """ConvNeXtBlock.forward (scOT/model.py:198-217)."""
    b, _, c = x.shape
    z = x.reshape(b, res, res, c).permute(0, 3, 1, 2)
    z = F.conv2d(z, p[pre + ".dwconv.weight"], p[pre + ".dwconv.bias"], padding=3, groups=c).permute(0, 2, 3, 1)
    z = cond_layer_norm(z, t, p, pre + ".norm", eps, cond)
    z = F.linear(z, p[pre + ".pwconv1.weight"], p[pre + ".pwconv1.bias"])
    z = F.gelu(z)
    z = F.linear(z, p[pre + ".pwconv2.weight"], p[pre + ".pwconv2.bias"])
    z = p[pre + ".weight"] * z
    return x + z.reshape(b, res * res, c)


def scot_loss(pred: Tensor, labels: Tensor, cfg) -> Tensor:
    """scOT/model.py:1425-1484."""
    fn = F.l1_loss if cfg.p == 1 else F.mse_loss
    sl = cfg.channel_slice_list_normalized_loss
    if sl is None:
        return fn(pred, labels)
    terms = []
    for i in range(len(sl) - 1):
        a, b = sl[i], sl[i + 1]
        terms.append(fn(pred[:, a:b], labels[:, a:b]) / (fn(labels[:, a:b], torch.zeros_like(labels[:, a:b])) + 1e-10))
    return torch.mean(torch.stack(terms))


def scot_forward(cfg, p: Dict[str, Tensor], pixel_values: Tensor, time: Optional[Tensor] = None,
                 labels: Optional[Tensor] = None, pixel_mask: Optional[Tensor] = None):
    """ScOT.forward (scOT/model.py:1318-1509) for inputs whose resolution equals cfg.image_size.

    Returns (loss or None, prediction [B, out, H, W]). `cfg` is any object with the ScOTConfig fields.
    """
    cond = bool(cfg.use_conditioning)
    eps = cfg.layer_norm_eps
    geo = stage_geometry(cfg)
    ps = cfg.patch_size
    b = pixel_values.shape[0]
    # embeddings (model.py:295-310, 345-366); embeddings.norm uses the default eps (model.py:342)
    x = F.conv2d(pixel_values, p["embeddings.patch_embeddings.projection.weight"],
                 p["embeddings.patch_embeddings.projection.bias"], stride=ps).flatten(2).transpose(1, 2)
    x = cond_layer_norm(x, time, p, "embeddings.norm", 1e-5, cond)
    # encoder (model.py:816-861, 1008-1099)
    skips = []
    ns = len(geo)
    for s, g in enumerate(geo):
        inp = x
        for i in range(g["depth"]):
            shift = 0 if i % 2 == 0 else g["shift"]
            x = scot_layer(x, time, p, f"encoder.layers.{s}.blocks.{i}", g, shift, eps, cond)
        skips.append(x)
        if s < ns - 1:
            x = patch_merging(x + inp, time, p, f"encoder.layers.{s}.downsample", g["res"], cond)
    # residual (ConvNeXt) blocks on the skips (model.py:1388-1393)
    for s, g in enumerate(geo):
        nblk = int(cfg.skip_connections[s]) if s < len(cfg.skip_connections) else 0
        for k in range(nblk):
            skips[s] = convnext_block(skips[s], time, p, f"residual_blocks.{s}.{k}", g["res"], eps, cond)
    # decoder (model.py:916-961, 1145-1240); decoder.layers[j] is stage ns-1-j; block list is built reversed
    x = skips[-1]
    for j in range(ns):
        s = ns - 1 - j
        g = geo[s]
        if j > 0:
            x = x + skips[s]
        for bi in range(g["depth"]):
            i = g["depth"] - 1 - bi  # original index of stored block bi (model.py:900)
            shift = 0 if i % 2 == 0 else g["shift"]
            x = scot_layer(x, time, p, f"decoder.layers.{j}.blocks.{bi}", g, shift, eps, cond)
        if s > 0:
            x = patch_unmerging(x, time, p, f"decoder.layers.{j}.upsample", g["res"], cond)
    # patch recovery (model.py:639-647)
    res0 = geo[0]["res"]
    z = x.transpose(1, 2).reshape(b, -1, res0, res0)
    z = F.conv_transpose2d(z, p["patch_recovery.projection.weight"], p["patch_recovery.projection.bias"], stride=ps)
    pred = F.conv2d(z, p["patch_recovery.mixup.weight"], None, padding=2)
    if getattr(cfg, "learn_residual", False):
        pv = pixel_values[:, : cfg.num_out_channels] if cfg.num_channels > cfg.num_out_channels else pixel_values
        pred = pred + pv
    if pixel_mask is not None:
        pred = pred.clone()
        pred[pixel_mask] = labels[pixel_mask].to(pred.dtype)
    loss = scot_loss(pred, labels, cfg) if labels is not None else None
    return loss, pred
