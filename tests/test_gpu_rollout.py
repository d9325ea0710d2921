"""GPU: device-side autoregressive rollout (poseidon_b200.runtime.ARRollout) against (1) the reference's rollout loop
(scOT/trainer.py:452-603 semantics) written around the public `model(**inputs)` call and (2) the CPU oracle (fp64)
stepped the same way. bf16 operand noise compounds over the steps: 3e-2 per step (see test_gpu_model.py)."""
import os
import types

import pytest
import torch

from oracle import scot_oracle as O
from oracle.weights import make_inputs, make_weights

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


def _setup(cin, cout):
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    rec = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    cfgd = dict(rec["config"])
    cfgd.update(num_channels=cin, num_out_channels=cout, channel_slice_list_normalized_loss=None)
    cfg = ScOTConfig(**cfgd)
    model = ScOT(cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    w = make_weights(shapes, seed=0)
    model.load_state_dict(w, strict=True)
    model = model.cuda()
    x, t, y, _ = make_inputs(2, cin, cout, cfg.image_size, seed=0)
    return cfgd, cfg, w, model, x, t, y


@pytest.mark.parametrize("cin,cout,steps", [(3, 3, 3), (4, 3, 2), (3, 3, [1, 2])])
def test_rollout_matches_reference_loop_and_oracle(cin, cout, steps):
    from poseidon_b200.runtime import ARRollout

    cfgd, cfg, w, model, x, t, y = _setup(cin, cout)
    factors = [1.0 / steps] * steps if isinstance(steps, int) else [float(i) for i in steps]
    # (1) the reference Trainer's loop around the public forward
    inp = x.cuda()
    ref_outs = []
    with torch.no_grad():
        for f in factors:
            out = model(pixel_values=inp, time=t.cuda() * f).output
            ref_outs.append(out)
            inp = out if cin == cout else torch.cat([out, inp[:, cout:]], dim=1)
    ro = ARRollout(model, batch=2)
    last, loss = ro.run(x.cuda(), t.cuda(), steps)
    assert loss is None and torch.equal(last, ref_outs[-1])
    allsteps, _ = ro.run(x.cuda(), t.cuda(), steps, output_all_steps=True)
    assert allsteps.shape[1] == len(factors)
    for i, r in enumerate(ref_outs):
        assert torch.equal(allsteps[:, i], r), i
    # (2) the CPU oracle stepped the same way
    ocfg = types.SimpleNamespace(**cfgd)
    ocfg.layer_norm_eps, ocfg.learn_residual = 1e-5, False
    wd = {k: v.double() for k, v in w.items()}
    xin = x.double()
    with torch.no_grad():
        for i, f in enumerate(factors):
            _, pred = O.scot_forward(ocfg, wd, xin, t.double() * f, None, None)
            assert rel(allsteps[:, i].cpu(), pred) < 3e-2 * (i + 1), (i, rel(allsteps[:, i].cpu(), pred))
            xin = pred if cin == cout else torch.cat([pred, xin[:, cout:]], dim=1)


def test_rollout_with_labels_averages_the_step_losses():
    from poseidon_b200.runtime import ARRollout

    cfgd, cfg, w, model, x, t, y = _setup(3, 3)
    ro = ARRollout(model, batch=2, with_labels=True)
    last, loss = ro.run(x.cuda(), t.cuda(), 2, labels=y.cuda())
    inp, tot = x.cuda(), 0.0
    with torch.no_grad():
        for _ in range(2):
            out = model(pixel_values=inp, time=t.cuda() / 2, labels=y.cuda())
            tot += float(out.loss)
            inp = out.output
    assert abs(float(loss) - tot / 2) < 1e-5 * abs(tot)
    assert torch.equal(last, inp)
