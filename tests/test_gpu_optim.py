"""GPU: the fused flat-buffer optimizer step (scot_grad_sq_norm + scot_adamw_step through poseidon_b200.optim.FlatAdamW)
against torch.optim.AdamW + torch.nn.utils.clip_grad_norm_ — the reference's optimizer path (scOT/train.py:286,
HF Trainer max_grad_norm) — on the same parameters, groups and gradients. fp32 element-wise arithmetic: tolerance 2e-6
relative (fused multiply-add contraction and sqrt/div rounding differ in the last bits)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _model():
    from oracle.weights import make_inputs, make_weights
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    rec = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    cfg = ScOTConfig(**rec["config"])
    model = ScOT(cfg)
    model.load_state_dict(make_weights(rec["shapes"], seed=0), strict=True)
    model = model.cuda()
    x, t, y, pm = make_inputs(rec["batch"], cfg.num_channels, cfg.num_out_channels, cfg.image_size, seed=0,
                              mask_channels=rec["mask_channels"])
    return model, (x.cuda(), t.cuda(), y.cuda(), pm.cuda())


@pytest.mark.parametrize("clip", [None, 0.05])
def test_flat_adamw_matches_torch_adamw(clip):
    from poseidon_b200.optim import FlatAdamW, build_param_groups

    model, (x, t, y, pm) = _model()
    model.grad_mode = "assign"
    model(pixel_values=x, time=t, labels=y, pixel_mask=pm).loss.backward()  # creates the flat buffers, binds .grad
    groups = build_param_groups(model, 0.01, 5e-4, 1e-4)
    opt = FlatAdamW(groups, model, lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01, max_grad_norm=clip)
    # reference: independent copies of the parameters, same grouping by name
    names = {id(p): n for n, p in model.named_parameters()}
    ref = {n: p.detach().clone().requires_grad_(True) for n, p in model.named_parameters()}
    ref_groups = [{**{k: v for k, v in g.items() if k != "params"}, "params": [ref[names[id(p)]] for p in g["params"]]}
                  for g in groups]
    ropt = torch.optim.AdamW(ref_groups, lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    rsched = torch.optim.lr_scheduler.LambdaLR(ropt, lambda s: 1.0 / (1 + s))
    gen = torch.Generator(device="cuda").manual_seed(3)
    flat = model.flat_parameters
    covered = torch.zeros_like(flat, dtype=torch.bool)
    for p in model.parameters():
        off = (p.data_ptr() - flat.data_ptr()) // 4
        covered[off:off + p.numel()] = True
    for step in range(4):
        gflat = model.flat_gradients
        # the engine only ever writes gradients of real parameters: padding / the key-bias slot stay zero
        gflat.copy_(torch.randn(gflat.shape, generator=gen, device="cuda") * (0.1 + step) * covered)
        for n, p in model.named_parameters():
            ref[n].grad = p.grad.detach().clone()
        if clip is not None:
            total = torch.nn.utils.clip_grad_norm_(list(ref.values()), clip)
        opt.step()
        ropt.step()
        sched.step()
        rsched.step()
        if clip is not None:
            assert abs(float(opt.grad_norm()) - float(total)) < 1e-5 * float(total)
        for n, p in model.named_parameters():
            err = float((p.detach() - ref[n].detach()).abs().max())
            assert err <= 2e-6 * float(ref[n].detach().abs().max()) + 1e-9, (step, n, err)
    # padding and the (bias-less) key slot of the fused [bq | 0 | bv] vector stay untouched
    assert float(flat[~covered].abs().max()) == 0.0
    # the next forward sees the updated weights (bf16 mirror is refreshed by the engine)
    out = model(pixel_values=x, time=t, labels=y, pixel_mask=pm)
    assert torch.isfinite(out.loss)


def test_flat_adamw_state_dict_roundtrip():
    from poseidon_b200.optim import FlatAdamW, build_param_groups

    model, (x, t, y, pm) = _model()
    model.grad_mode = "assign"
    model(pixel_values=x, time=t, labels=y, pixel_mask=pm).loss.backward()
    opt = FlatAdamW(build_param_groups(model, 0.01), model, lr=1e-3)
    opt.step()
    opt.step()
    sd = opt.state_dict()
    p0 = model.flat_parameters.clone()
    opt2 = FlatAdamW(build_param_groups(model, 0.01), model, lr=1e-3)
    opt2.load_state_dict(sd)
    opt.step()
    p_a = model.flat_parameters.clone()
    model.flat_parameters.copy_(p0)
    opt2.step()
    assert torch.equal(p_a, model.flat_parameters)
