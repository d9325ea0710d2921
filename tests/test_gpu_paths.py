"""Product paths of ScOT.forward that the fixtures of the shipped training configs do not reach (GPU), each against the
fp64 oracle on a small model in BOTH precisions (parity: tight, bf16: the bf16 noise floor):
  * per-pixel boolean mask with the full prediction shape (SE-AF / Airfoil: scOT/problems/fluids/compressible.py:46-53,
    scOT/model.py:1422-1423) — forward overwrite, loss, and gradients (zero through masked pixels);
  * learn_residual=True (scOT/model.py:1411);
  * p=1 without channel_slice_list (plain L1) and p=2 with slices;
  * inputs whose resolution differs from config.image_size (spectral resize, scOT/model.py:1293-1316,1360-1366,1416-1420).
"""
import math
import types

import numpy as np
import pytest
import torch

from oracle import scot_oracle as O
from oracle.weights import make_inputs, make_weights

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


BASE = dict(image_size=32, patch_size=4, num_channels=3, num_out_channels=3, embed_dim=32, depths=[2, 2], num_heads=[2, 4],
            skip_connections=[1, 0], window_size=4, mlp_ratio=4.0, drop_path_rate=0.0, use_conditioning=True, p=1,
            channel_slice_list_normalized_loss=[0, 1, 3], residual_model="convnext", learn_residual=False)
# (output / loss, global gradient). The objective contains the L1 loss: in bf16 a prediction that lands on the other side of
# its label flips sign(pred - label), so the bf16 gradient bound is looser than the smooth-objective bound of test_gpu_model
TOL = {"parity": (2e-4, 2e-3), "bf16": (3e-2, 0.15)}


def build(precision, **over):
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    cfgd = dict(BASE, **over)
    cfg = ScOTConfig(**cfgd)
    model = ScOT(cfg)
    w = make_weights({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=3)
    model.load_state_dict(w, strict=True)
    model = model.cuda()
    model.precision = precision
    ocfg = types.SimpleNamespace(**cfgd)
    ocfg.layer_norm_eps = 1e-5
    return cfg, ocfg, w, model


def oracle(ocfg, w, x, t, y, pm):
    wr = {k: v.double().requires_grad_(True) for k, v in w.items()}
    loss, pred = O.scot_forward(ocfg, wr, x.double(), t.double(), y.double() if y is not None else None, pm)
    return loss, pred, wr


def grad_err(model, wr):
    g = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
    num = sum((g[k].double() - wr[k].grad).pow(2).sum() for k in g)
    den = sum(wr[k].grad.pow(2).sum() for k in g)
    return float((num / den).sqrt())


@pytest.mark.parametrize("precision", ["parity", "bf16"])
def test_per_pixel_mask(precision):
    cfg, ocfg, w, model = build(precision)
    x, t, y, _ = make_inputs(4, 3, 3, 32, seed=1)
    pm = torch.rand(4, 3, 32, 32, generator=torch.Generator().manual_seed(5)) < 0.25  # full-shape boolean mask
    out = model(pixel_values=x.cuda(), time=t.cuda(), labels=y.cuda(), pixel_mask=pm.cuda())
    out.loss.backward()
    loss, pred, wr = oracle(ocfg, w, x, t, y, pm)
    loss.backward()
    to, tg = TOL[precision]
    assert torch.equal(out.output.detach().cpu()[pm], y[pm])  # exact copy of the labels under the mask
    assert rel(out.output.detach().cpu(), pred.detach()) < to
    assert abs(float(out.loss) - float(loss)) < to * abs(float(loss))
    if precision == "parity":  # the L1 sign() makes bf16 loss gradients incomparable element-wise; parity is smooth enough
        assert grad_err(model, wr) < 2e-2
    # an Airfoil-style single-channel model takes the [B,1,H,W] mask as is
    cfg1, ocfg1, w1, model1 = build(precision, num_channels=1, num_out_channels=1, channel_slice_list_normalized_loss=[0, 1])
    x1, t1, y1, _ = make_inputs(2, 1, 1, 32, seed=2)
    pm1 = x1 > 0.8
    o1 = model1(pixel_values=x1.cuda(), time=t1.cuda(), labels=y1.cuda(), pixel_mask=pm1.cuda())
    l1, p1, _ = oracle(ocfg1, w1, x1, t1, y1, pm1)
    assert torch.equal(o1.output.detach().cpu()[pm1], y1[pm1]) and rel(o1.output.detach().cpu(), p1.detach()) < to
    assert abs(float(o1.loss) - float(l1)) < to * abs(float(l1))


@pytest.mark.parametrize("precision", ["parity", "bf16"])
@pytest.mark.parametrize("over", [dict(learn_residual=True), dict(p=1, channel_slice_list_normalized_loss=None),
                                  dict(p=2, channel_slice_list_normalized_loss=[0, 2, 3]),
                                  dict(learn_residual=True, num_channels=4, channel_slice_list_normalized_loss=None, p=2)])
def test_residual_and_loss_variants(precision, over):
    cfg, ocfg, w, model = build(precision, **over)
    x, t, y, _ = make_inputs(3, cfg.num_channels, 3, 32, seed=4)
    out = model(pixel_values=x.cuda(), time=t.cuda(), labels=y.cuda())
    G = torch.randn(out.output.shape, generator=torch.Generator().manual_seed(9))
    (out.loss + (out.output * G.cuda()).sum()).backward()
    loss, pred, wr = oracle(ocfg, w, x, t, y, None)
    (loss + (pred * G.double()).sum()).backward()
    to, tg = TOL[precision]
    assert rel(out.output.detach().cpu(), pred.detach()) < to
    assert abs(float(out.loss) - float(loss)) < to * abs(float(loss))
    assert grad_err(model, wr) < tg


def _fft_resize(img, target):
    """independent numpy restatement of ScOT._upsample/_downsample (scOT/model.py:1293-1316): keep / zero-pad the
    centred spectrum, norm='forward'"""
    a = np.fft.fftshift(np.fft.fft2(img.double().numpy(), norm="forward"), axes=(-2, -1))
    n = a.shape[-1]
    if target > n:
        p = (target - n) // 2
        a = np.pad(a, [(0, 0), (0, 0), (p, p), (p, p)])
    else:
        lo = n // 2 - target // 2
        a = a[..., lo:lo + target, lo:lo + target]
    return torch.from_numpy(np.fft.ifft2(np.fft.ifftshift(a, axes=(-2, -1)), norm="forward").real)


@pytest.mark.parametrize("size", [16, 64])
def test_resized_inputs(size):
    """resolution != config.image_size: input resampled spectrally to 32, prediction resampled back, loss on the result"""
    cfg, ocfg, w, model = build("parity")
    x, t, y, _ = make_inputs(2, 3, 3, size, seed=6)
    out = model(pixel_values=x.cuda(), time=t.cuda(), labels=y.cuda())
    assert tuple(out.output.shape) == (2, 3, size, size)
    xr = _fft_resize(x, 32)
    _, pred, _ = oracle(ocfg, w, xr, t, None, None)
    pr = _fft_resize(pred.detach(), size)
    assert rel(out.output.detach().cpu(), pr) < 2e-4
    loss = O.scot_loss(pr, y.double(), ocfg)
    assert abs(float(out.loss) - float(loss)) < 2e-4 * abs(float(loss))
    # inference without labels takes the same path
    with torch.no_grad():
        o2 = model(pixel_values=x.cuda(), time=t.cuda())
    assert rel(o2.output.cpu(), pr) < 2e-4
