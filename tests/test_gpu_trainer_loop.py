"""The training-step contract of the reference's Trainer on the drop-in model (GPU): what HF `Trainer.training_step`
+ the inner loop do (scOT/trainer.py:605-635 -> transformers Trainer): `loss = model(**batch).loss; loss.backward();
clip_grad_norm_(params, max_grad_norm); optimizer.step(); model.zero_grad()` with the default collator's kwargs
(pixel_values / labels / time / pixel_mask) and torch.optim.AdamW over the reference's parameter groups. Two steps are
run with (a) stock torch.optim.AdamW in the default grad_mode="autograd" and (b) FlatAdamW — in autograd mode (gathers
p.grad) and in assign mode — and compared with the fp64 oracle trained the same way; parity precision => tight bounds.
"""
import types

import pytest
import torch

from oracle import scot_oracle as O
from oracle.weights import make_inputs, make_weights

pytestmark = pytest.mark.gpu

CFG = dict(image_size=32, patch_size=4, num_channels=3, num_out_channels=3, embed_dim=32, depths=[2, 2], num_heads=[2, 4],
           skip_connections=[1, 0], window_size=4, mlp_ratio=4.0, drop_path_rate=0.0, use_conditioning=True, p=2,
           channel_slice_list_normalized_loss=[0, 1, 3], residual_model="convnext")
LR, WD, CLIP, STEPS = 1e-2, 0.01, 5.0, 2


def batches():
    out = []
    for s in range(STEPS):
        x, t, y, pm = make_inputs(4, 3, 3, 32, seed=10 + s, mask_channels=(2,))
        out.append({"pixel_values": x, "labels": y, "time": t, "pixel_mask": pm})
    return out


def oracle_training(w):
    from poseidon_b200.optim import build_param_groups
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    ocfg = types.SimpleNamespace(**CFG)
    ocfg.layer_norm_eps, ocfg.learn_residual = 1e-5, False
    with torch.device("meta"):
        meta = ScOT(ScOTConfig(**CFG))
    names = {id(p): n for n, p in meta.named_parameters()}
    wr = {k: v.double().clone().requires_grad_(True) for k, v in w.items()}
    groups = [{"params": [wr[names[id(p)]] for p in g["params"]], "weight_decay": g["weight_decay"]}
              for g in build_param_groups(meta, WD)]
    opt = torch.optim.AdamW(groups, lr=LR)
    losses = []
    for b in batches():
        loss, _ = O.scot_forward(ocfg, wr, b["pixel_values"].double(), b["time"].double(), b["labels"].double(), b["pixel_mask"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(wr.values()), CLIP)
        opt.step()
        opt.zero_grad()
        losses.append(float(loss))
    return losses, {k: v.detach() for k, v in wr.items()}


@pytest.fixture(scope="module")
def reference_run():
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    with torch.device("meta"):
        meta = ScOT(ScOTConfig(**CFG))
    w = make_weights({k: tuple(v.shape) for k, v in meta.state_dict().items()}, seed=7)
    return w, oracle_training(w)


@pytest.mark.parametrize("optimizer,grad_mode", [("torch", "autograd"), ("flat", "autograd"), ("flat", "assign")])
def test_two_training_steps_match_the_oracle(reference_run, optimizer, grad_mode):
    from poseidon_b200.optim import FlatAdamW, build_param_groups
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    w, (ref_losses, ref_w) = reference_run
    model = ScOT(ScOTConfig(**CFG))
    model.load_state_dict(w, strict=True)
    model = model.cuda()
    model.precision = "parity"
    model.grad_mode = grad_mode
    groups = build_param_groups(model, WD)
    if optimizer == "torch":
        opt = torch.optim.AdamW(groups, lr=LR)
    else:
        opt = FlatAdamW(groups, model, lr=LR, max_grad_norm=CLIP)
    losses = []
    for b in batches():
        out = model(**{k: v.cuda() for k, v in b.items()})
        out.loss.backward()
        if optimizer == "torch":
            torch.nn.utils.clip_grad_norm_(model.parameters(), CLIP)
        opt.step()
        model.zero_grad()  # HF Trainer: model.zero_grad() (set_to_none=True)
        losses.append(float(out.loss))
    for a, b_ in zip(losses, ref_losses):
        assert abs(a - b_) < 2e-4 * abs(b_), (losses, ref_losses)
    num = sum((p.detach().cpu().double() - ref_w[k]).pow(2).sum() for k, p in model.named_parameters())
    den = sum((ref_w[k] - w[k].double()).pow(2).sum() for k in ref_w)  # relative to the size of the UPDATE
    assert float((num / den).sqrt()) < 2e-2


def test_gradient_accumulation_in_assign_mode():
    """two micro-batches accumulate in the flat buffer; zero_grad(set_to_none=True) starts a new window"""
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    w = make_weights({k: tuple(v.shape) for k, v in ScOT(ScOTConfig(**CFG)).state_dict().items()}, seed=7)
    model = ScOT(ScOTConfig(**CFG))
    model.load_state_dict(w, strict=True)
    model = model.cuda()
    model.precision = "parity"
    model.grad_mode = "assign"
    bs = batches()
    gs = []
    for b in bs:
        model.zero_grad()
        model(**{k: v.cuda() for k, v in b.items()}).loss.backward()
        gs.append(model.flat_gradients.clone())
    model.zero_grad()
    for b in bs:
        model(**{k: v.cuda() for k, v in b.items()}).loss.backward()
    acc = model.flat_gradients.clone()
    assert float((acc - (gs[0] + gs[1])).norm() / acc.norm()) < 1e-5
    # the torch default optimizer.zero_grad(set_to_none=True) only drops .grad: the next backward must not pile up
    for p in model.parameters():
        p.grad = None
    model(**{k: v.cuda() for k, v in bs[0].items()}).loss.backward()
    assert float((model.flat_gradients - gs[0]).norm() / gs[0].norm()) < 1e-5
