"""Input pipeline and rollout at scale (GPU): DevicePrefetcher ordering / slot reuse under back-pressure, and the
autoregressive rollout at the C5 geometry (Poseidon-L widths, 256 x 256, shifted windows in stages 0 and 1) against the
fp64 oracle stepped the same way (reference scOT/trainer.py:452-603)."""
import types

import pytest
import torch

from oracle import scot_oracle as O
from oracle.weights import make_weights

pytestmark = pytest.mark.gpu


def test_device_prefetcher_order_and_slot_reuse():
    from poseidon_b200.runtime import DevicePrefetcher

    n = 9
    host = [{"pixel_values": torch.full((2, 3, 64, 64), float(i)).pin_memory(), "time": torch.full((2,), float(i)).pin_memory()}
            for i in range(n)]
    feed = DevicePrefetcher(iter(host), "cuda", depth=2)
    sink = torch.zeros(n, device="cuda")
    big = torch.randn(4096, 4096, device="cuda")
    seen_ptrs = set()
    for i, batch in enumerate(feed):
        # a slow consumer: the copy of batch i+2 must not overwrite slot (i % 2) before this kernel has read it
        for _ in range(3):
            big = big @ big * 1e-4
        sink[i] = batch["pixel_values"].mean() + batch["time"].mean()
        seen_ptrs.add(batch["pixel_values"].data_ptr())
    torch.cuda.synchronize()
    assert i == n - 1
    assert torch.equal(sink.cpu(), torch.arange(n, dtype=torch.float32) * 2)
    assert len(seen_ptrs) == 2  # two device slots, reused
    with pytest.raises(StopIteration):
        next(feed)


def test_rollout_at_c5_geometry_matches_oracle():
    """Poseidon-L widths (embed 192, head_dim 64) at 256 x 256 with reduced depth; 3 autoregressive steps, parity mode"""
    from poseidon_b200.runtime import ARRollout
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    cfgd = dict(image_size=256, patch_size=4, num_channels=5, num_out_channels=5, embed_dim=192, depths=[2, 2, 2, 2],
                num_heads=[3, 6, 12, 24], skip_connections=[1, 1, 1, 0], window_size=16, mlp_ratio=4.0, drop_path_rate=0.0,
                use_conditioning=True, p=1, channel_slice_list_normalized_loss=[0, 1, 3, 4, 5], residual_model="convnext")
    model = ScOT(ScOTConfig(**cfgd))
    w = make_weights({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=2)
    model.load_state_dict(w, strict=True)
    model = model.cuda()
    model.precision = "parity"
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 5, 256, 256, generator=g)
    t = torch.rand(1, generator=g)
    steps = 3
    ro = ARRollout(model, 1)
    out, _ = ro.run(x.cuda(), t.cuda(), steps, output_all_steps=True)
    ocfg = types.SimpleNamespace(**cfgd)
    ocfg.layer_norm_eps, ocfg.learn_residual = 1e-5, False
    wd = {k: v.double() for k, v in w.items()}
    cur = x.double()
    with torch.no_grad():
        for s in range(steps):
            _, pred = O.scot_forward(ocfg, wd, cur, t.double() / steps, None, None)
            r = float((out[:, s].cpu().double() - pred).norm() / pred.norm())
            assert r < 1e-3 * (s + 1), (s, r)
            cur = pred
    # the bf16 engine on the same rollout stays at the bf16 noise floor per step
    model.precision = "bf16"
    ro2 = ARRollout(model, 1)
    out2, _ = ro2.run(x.cuda(), t.cuda(), steps, output_all_steps=True)
    assert float((out2[:, 0].cpu().double() - out[:, 0].cpu().double()).norm() / out[:, 0].cpu().double().norm()) < 3e-2
