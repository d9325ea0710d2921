"""GPU parity for the remaining BASELINE.json configurations, at reduced depth so that the fp64 CPU oracle finishes in
seconds: Poseidon-L geometry (embed 192 -> head_dim 64, C up to 1536; configs[3]) and a 256x256 grid (64x64 tokens, 16
windows per sample at stage 0, shifted windows in stages 0 AND 1; configs[4]). Engine vs oracle (fp64) on the same seeded
weights / inputs in both precisions: bf16 at the bf16 operand noise floor (as test_gpu_model.py), parity (split-bf16 GEMMs,
fp32 attention) at the north-star tolerance."""
import types

import pytest
import torch

from oracle import scot_oracle as O
from oracle.weights import make_inputs, make_weights

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


CASES = {
    "L128_lite": dict(image_size=128, patch_size=4, num_channels=5, num_out_channels=5, embed_dim=192, depths=[2, 2, 1, 1],
                      num_heads=[3, 6, 12, 24], skip_connections=[1, 1, 1, 0], window_size=16, mlp_ratio=4.0,
                      drop_path_rate=0.0, use_conditioning=True, p=1, channel_slice_list_normalized_loss=[0, 1, 3, 4, 5],
                      residual_model="convnext"),
    "T256_lite": dict(image_size=256, patch_size=4, num_channels=4, num_out_channels=4, embed_dim=48, depths=[2, 2, 2, 2],
                      num_heads=[3, 6, 12, 24], skip_connections=[2, 1, 1, 0], window_size=16, mlp_ratio=4.0,
                      drop_path_rate=0.0, use_conditioning=True, p=1, channel_slice_list_normalized_loss=[0, 1, 3, 4],
                      residual_model="convnext"),
}


TOL = {"bf16": (3e-2, 1e-2, 0.1), "parity": (2e-4, 1e-4, 2e-3)}  # output rel-L2, loss rel, gradient rel-L2 (global and median)


@pytest.mark.parametrize("precision", ["bf16", "parity"])
@pytest.mark.parametrize("name", list(CASES))
def test_engine_matches_oracle_on_other_baseline_geometries(name, precision):
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    cfgd = CASES[name]
    cfg = ScOTConfig(**cfgd)
    model = ScOT(cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    w = make_weights(shapes, seed=0)
    model.load_state_dict(w, strict=True)
    model = model.cuda()
    model.precision = precision
    to, tl, tg = TOL[precision]
    x, t, y, pm = make_inputs(1, cfg.num_channels, cfg.num_out_channels, cfg.image_size, seed=0)
    out = model(pixel_values=x.cuda(), time=t.cuda(), labels=y.cuda())
    G = torch.randn(out.output.shape, generator=torch.Generator().manual_seed(7))
    out.output.backward(G.cuda())
    ocfg = types.SimpleNamespace(**cfgd)
    ocfg.layer_norm_eps, ocfg.learn_residual = 1e-5, False
    wr = {k: v.double().requires_grad_(True) for k, v in w.items()}
    loss, pred = O.scot_forward(ocfg, wr, x.double(), t.double(), y.double(), None)
    (pred * G.double()).sum().backward()
    assert rel(out.output.cpu(), pred.detach()) < to, rel(out.output.cpu(), pred.detach())
    assert abs(float(out.loss.detach()) - float(loss.detach())) < tl * float(loss.detach())
    grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
    assert all(g is not None and torch.isfinite(g).all() for g in grads.values())
    num = sum((grads[k].double() - wr[k].grad).pow(2).sum() for k in grads)
    den = sum(wr[k].grad.pow(2).sum() for k in grads)
    assert float((num / den).sqrt()) < tg, float((num / den).sqrt())
    errs = torch.tensor([rel(grads[k], wr[k].grad) for k in grads])
    assert float(errs.median()) < tg, float(errs.median())
