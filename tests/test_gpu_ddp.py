"""Data-parallel correctness on real GPUs (needs >= 2 devices: `gpurun --gpus 2 -- pytest tests/test_gpu_ddp.py`):
two ranks, each `GraphedTrainStep` on its half of the batch + THE single NCCL all-reduce of the flat gradient buffer
(runtime.GraphedTrainStep.allreduce), must reproduce the gradient of the full batch computed by one rank; then one
FlatAdamW step must leave both replicas bit-identical."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(image_size=32, patch_size=4, num_channels=2, num_out_channels=2, embed_dim=32, depths=[2, 2], num_heads=[2, 4],
           skip_connections=[1, 0], window_size=4, mlp_ratio=4.0, drop_path_rate=0.0, use_conditioning=True, p=2,
           channel_slice_list_normalized_loss=None, residual_model="convnext")


def _worker(rank, world, port, out_path):
    import torch.distributed as dist

    from oracle.weights import make_inputs, make_weights
    from poseidon_b200.optim import FlatAdamW, build_param_groups
    from poseidon_b200.runtime import GraphedTrainStep
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    model = ScOT(ScOTConfig(**CFG))
    w = make_weights({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=1)
    model.load_state_dict(w, strict=True)
    model = model.to(dev)
    model.precision = "parity"
    x, t, y, _ = make_inputs(8, 2, 2, 32, seed=3)
    sl = slice(rank * 4, rank * 4 + 4)  # shard by PDE sample
    step = GraphedTrainStep(model, 4, dev, world_size=world)
    step.load_batch(x[sl].to(dev), t[sl].to(dev), y[sl].to(dev))
    step.run()
    step.allreduce()
    torch.cuda.synchronize(dev)
    g = model.flat_gradients.clone()
    opt = FlatAdamW(build_param_groups(model, 0.01), model, lr=1e-2, max_grad_norm=5.0)
    step.optimizer = opt
    step.optimizer_step()
    torch.cuda.synchronize(dev)
    flat_after = model.flat_parameters.clone()
    gathered = [torch.empty_like(flat_after) for _ in range(world)]
    dist.all_gather(gathered, flat_after)
    if rank == 0:
        # full batch on one rank (MSE: plain mean over samples -> mean of the per-rank gradients)
        m2 = ScOT(ScOTConfig(**CFG))
        m2.load_state_dict(w, strict=True)
        m2 = m2.to(dev)
        m2.precision = "parity"
        s2 = GraphedTrainStep(m2, 8, dev, world_size=1)
        s2.load_batch(x.to(dev), t.to(dev), y.to(dev))
        s2.run()
        torch.cuda.synchronize(dev)
        ref = m2.flat_gradients
        torch.save({"rel": float((g - ref).norm() / ref.norm()), "replicas_equal": bool(torch.equal(gathered[0], gathered[1])),
                    "moved": float((flat_after - m2.flat_parameters).abs().max())}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_step_equals_full_batch(tmp_path):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "ddp.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    rec = torch.load(out)
    assert rec["rel"] < 1e-4, rec          # sharded + all-reduced gradient == full-batch gradient
    assert rec["replicas_equal"], rec      # one collective, identical optimizer step on both ranks
    assert rec["moved"] > 0


def test_split_backward_equals_whole_backward():
    """single GPU: the two-part backward (scot_engine_backward_part 1 + 2, two CUDA graphs) gives the gradient of the
    one-part backward; the range [grad_split, end) is final after part 1 (checked by running part 1 alone)"""
    from oracle.weights import make_inputs, make_weights
    from poseidon_b200.runtime import GraphedTrainStep
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    dev = torch.device("cuda", 0)
    cfg = dict(CFG, image_size=64, depths=[2, 2, 2], num_heads=[2, 4, 8], skip_connections=[1, 1, 0])
    x, t, y, _ = make_inputs(4, 2, 2, 64, seed=3)
    grads = {}
    for mode in ("whole", "split"):
        model = ScOT(ScOTConfig(**cfg))
        w = make_weights({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=1)
        model.load_state_dict(w, strict=True)
        model = model.to(dev)
        model.precision = "parity"
        step = GraphedTrainStep(model, 4, dev, world_size=1, overlap_allreduce=(mode == "split"))
        step.load_batch(x.to(dev), t.to(dev), y.to(dev))
        step.run()
        torch.cuda.synchronize()
        grads[mode] = model.flat_gradients.clone()
        if mode == "split":
            split = step.split
            assert 0 < split < grads[mode].numel()
            step.graph.replay()  # part 1 only (zeroes, forward, first half of the backward)
            torch.cuda.synchronize()
            part1 = model.flat_gradients.clone()
            assert float((part1[split:] - grads[mode][split:]).norm() / grads[mode][split:].norm()) < 1e-5
            assert float(part1[:split].abs().max()) == 0.0  # nothing of the second range is touched by part 1
    assert float((grads["split"] - grads["whole"]).norm() / grads["whole"].norm()) < 1e-5
