"""Host logic of the data-parallel path on CPU (gloo, world_size 2): per-rank sample sharding and the single
all-reduce over one flat gradient buffer reproduce the single-process mean gradient (what torch DDP does with ~25
bucketed all-reduces in the reference's accelerate launch)."""
import os
import socket
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import scot_oracle as O
from oracle.weights import make_inputs, make_weights

CFG = dict(image_size=32, patch_size=4, num_channels=2, num_out_channels=2, embed_dim=16, depths=[1, 1], num_heads=[1, 2],
           skip_connections=[1, 0], window_size=4, mlp_ratio=2.0, drop_path_rate=0.0, use_conditioning=True, p=2,
           channel_slice_list_normalized_loss=None, residual_model="convnext")


def _shapes():
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    with torch.device("meta"):
        m = ScOT(ScOTConfig(**CFG))
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def _grads(w, x, t, y):
    cfg = types.SimpleNamespace(**CFG)
    cfg.layer_norm_eps, cfg.learn_residual = 1e-5, False
    wr = {k: v.double().requires_grad_(True) for k, v in w.items()}
    loss, _ = O.scot_forward(cfg, wr, x.double(), t.double(), y.double(), None)
    loss.backward()
    return float(loss), wr


def _flatten(table, elems, grads):
    flat = torch.zeros(elems, dtype=torch.float64)
    for k, (off, numel, shape) in table.items():
        flat[off:off + numel] = grads[k].grad.reshape(-1)
    return flat


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from poseidon_b200 import _lib
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    with torch.device("meta"):
        m = ScOT(ScOTConfig(**CFG))
    eng = _lib.Engine(m._desc(), batch=2)  # host-side plan only: parameter table of the flat buffer
    w = make_weights(_shapes(), seed=0)
    x, t, y, _ = make_inputs(4, 2, 2, 32, seed=0)
    sl = slice(rank * 2, rank * 2 + 2)  # shard by PDE sample
    _, wr = _grads(w, x[sl], t[sl], y[sl])
    flat = _flatten(eng.table, eng.param_elems, wr) / world  # pre-scaled loss gradient (runtime.GraphedTrainStep.gscale)
    dist.all_reduce(flat)  # THE single collective of the step
    if rank == 0:
        torch.save(flat, out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_equals_full_batch_gradient(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    flat = torch.load(out)
    from poseidon_b200 import _lib
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    with torch.device("meta"):
        m = ScOT(ScOTConfig(**CFG))
    eng = _lib.Engine(m._desc(), batch=4)
    w = make_weights(_shapes(), seed=0)
    x, t, y, _ = make_inputs(4, 2, 2, 32, seed=0)
    _, wr = _grads(w, x, t, y)  # MSE loss is a plain mean over samples -> mean of the per-rank gradients
    ref = _flatten(eng.table, eng.param_elems, wr)
    assert float((flat - ref).norm() / ref.norm()) < 1e-10
