import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from poseidon_b200.runtime import GraphedTrainStep
from poseidon_b200.scOT.model import ScOT, ScOTConfig
cfg = bench.model_config("B", 5)
model = ScOT(ScOTConfig(**cfg)); bench.realistic_init_(model); model = model.cuda()
B = 64
g = torch.Generator().manual_seed(1)
hx = torch.randn(B, 5, 128, 128, generator=g).pin_memory(); hy = torch.randn(B, 5, 128, 128, generator=g).pin_memory(); ht = torch.rand(B, generator=g).pin_memory()
def sync(): torch.cuda.synchronize()
def timeit(fn, n=5):
    fn(); sync(); t0 = time.perf_counter()
    for _ in range(n): fn()
    sync(); return (time.perf_counter() - t0) / n * 1e3
step = GraphedTrainStep(model, B, use_graph=False)
step.load_batch(hx, ht, hy)
print("eager engine step (no graph) ms:", timeit(step.run))
t0 = time.perf_counter(); 
for _ in range(5): step.run()
cpu_ms = (time.perf_counter() - t0) / 5 * 1e3; sync()
print("  host-side enqueue time ms:", cpu_ms)
gstep = GraphedTrainStep(model, B, use_graph=True); gstep.load_batch(hx, ht, hy)
print("graphed step ms:", timeit(gstep.run))
model.grad_mode = "assign"
for p in model.parameters(): p.grad = None
x = hx.cuda(); y = hy.cuda(); t = ht.cuda()
def api():
    model.flat_gradients.zero_()
    out = model(pixel_values=x, time=t, labels=y); out.loss.backward()
print("public API fwd+bwd (assign) ms:", timeit(api))
def api_fwd():
    with torch.no_grad(): model(pixel_values=x, time=t, labels=y)
print("public API fwd only (no_grad) ms:", timeit(api_fwd))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); api(); sync(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
