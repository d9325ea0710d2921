"""The side-branch schedules of the engine must not change results (GPU).

By default the ConvNeXt blocks on the skip connections run on a second stream beside the deeper stages
(SCOT_CNX_OVERLAP, engine.cu) and the dk/dv kernel of the window-attention backward runs on a third stream beside the dq
kernel (SCOT_ATTN_BWD_SPLIT). Both issue exactly the same kernels on the same data as the in-line order, so the
prediction must be BIT-IDENTICAL to the in-line schedule and the gradients equal up to the run-to-run noise of the fp32
atomics (split reductions; an fp32 last-bit difference occasionally flips the bf16 rounding of a downstream operand).
The noise floor is measured on the spot from independent engines with identical settings. Checked on a small model
that has every block type (shifted windows, ConvNeXt skips at two stages, merging / unmerging, conditioned norms),
eagerly and through the CUDA-graph step.
"""
import os

import pytest
import torch

from oracle.weights import make_inputs, make_weights

pytestmark = pytest.mark.gpu

INLINE = {"SCOT_CNX_OVERLAP": "0", "SCOT_ATTN_BWD_SPLIT": "0"}
CFG = dict(image_size=64, patch_size=4, num_channels=3, num_out_channels=3, embed_dim=32, depths=[2, 2, 2],
           num_heads=[2, 4, 8], skip_connections=[2, 1, 0], window_size=8, mlp_ratio=4.0, drop_path_rate=0.0,
           use_conditioning=True, p=1, channel_slice_list_normalized_loss=[0, 1, 3], residual_model="convnext")


def run(env, use_graph, batch=4, nrep=3):
    """One engine built under `env` (the knobs are read when an engine first runs); returns nrep (pred, loss, grads)."""
    from poseidon_b200.runtime import GraphedTrainStep
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    old = {k: os.environ.get(k) for k in INLINE}
    try:
        os.environ.update(INLINE)
        os.environ.update(env)
        cfg = ScOTConfig(**CFG)
        model = ScOT(cfg)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict(make_weights(shapes, seed=0), strict=True)
        model = model.cuda()
        x, t, y, _ = make_inputs(batch, 3, 3, 64, seed=0)
        step = GraphedTrainStep(model, batch, torch.device("cuda", 0), use_graph=use_graph)
        step.load_batch(x, t, y)
        outs = []
        for _ in range(nrep):  # replays must be reproducible as well
            step.run()
            torch.cuda.synchronize()
            outs.append((step.pred.clone(), step.loss.clone(), step.st["gflat"].clone()))
        return outs
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def rel(a, b):
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def inline():
    """In-line schedule: reference outputs + the run-to-run noise floor of its own gradients."""
    a = run({}, use_graph=False)
    b = run({}, use_graph=True)
    pred0, loss0, g0 = a[0]
    for pred, _, _ in a + b:
        assert torch.equal(pred, pred0)  # the forward pass is deterministic
    noise = max(rel(g, g0) for _, _, g in a[1:] + b)
    return pred0, loss0, g0, noise


@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "graph"])
@pytest.mark.parametrize("env", [
    {"SCOT_CNX_OVERLAP": "1"},
    {"SCOT_ATTN_BWD_SPLIT": "8"},
    {"SCOT_ATTN_BWD_SPLIT": "16"},
    {"SCOT_CNX_OVERLAP": "1", "SCOT_ATTN_BWD_SPLIT": "16"},
], ids=["cnx", "attn8", "attn16", "defaults"])
def test_side_branch_schedule_is_result_neutral(inline, env, use_graph):
    pred0, loss0, g0, noise = inline
    tol = max(10.0 * noise, 1e-5)
    for pred, loss, g in run(env, use_graph):
        assert torch.equal(pred, pred0)
        assert abs(float(loss) - float(loss0)) <= 1e-6 * abs(float(loss0))  # the loss sums are fp32 atomics
        assert torch.isfinite(g).all()
        assert rel(g, g0) < tol, (rel(g, g0), noise)
