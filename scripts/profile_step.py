"""Runs N eager (non-graph) train steps of Poseidon-B so that ncu can list / profile every kernel of one step."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from poseidon_b200 import _lib
from poseidon_b200.runtime import GraphedTrainStep
from poseidon_b200.scOT.model import ScOT, ScOTConfig

model_name = sys.argv[1] if len(sys.argv) > 1 else "B"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = bench.model_config(model_name, 5)
torch.manual_seed(0)
model = ScOT(ScOTConfig(**cfg))
bench.realistic_init_(model)
model = model.cuda()
step = GraphedTrainStep(model, batch, use_graph=False)
g = torch.Generator().manual_seed(1)
step.load_batch(torch.randn(batch, 5, 128, 128, generator=g), torch.rand(batch, generator=g), torch.randn(batch, 5, 128, 128, generator=g))
lib = _lib.load()
for i in range(nsteps):
    c0 = lib.scot_launch_count()
    step.run()
    torch.cuda.synchronize()
    print("step", i, "launches", lib.scot_launch_count() - c0, "loss", float(step.loss))

if len(sys.argv) > 4 and sys.argv[4] == "opt":  # one fused clip + AdamW step as well (for the ncu rows of optim.cu)
    from poseidon_b200.optim import FlatAdamW, build_param_groups
    opt = FlatAdamW(build_param_groups(model, 0.01), model, lr=1e-6, max_grad_norm=5.0)
    opt.step()
    torch.cuda.synchronize()
    print("optimizer step done")
