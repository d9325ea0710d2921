"""Do two half-batch training steps on two streams overlap usefully? (aggregate samples/s vs one full batch)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from poseidon_b200.runtime import GraphedTrainStep
from poseidon_b200.scOT.model import ScOT, ScOTConfig
cfg = bench.model_config("B", 5)
def mk(batch):
    m = ScOT(ScOTConfig(**cfg)); bench.realistic_init_(m); m = m.cuda()
    st = GraphedTrainStep(m, batch)
    g = torch.Generator().manual_seed(1)
    st.load_batch(torch.randn(batch, 5, 128, 128, generator=g), torch.rand(batch, generator=g), torch.randn(batch, 5, 128, 128, generator=g))
    return m, st
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
m64, s64 = mk(64)
print("batch 64 one stream ms:", timeit(s64.run))
del m64, s64; torch.cuda.empty_cache()
ma, sa = mk(32); mb, sb = mk(32)
print("batch 32 one stream ms:", timeit(sa.run))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): sa.run()
    with torch.cuda.stream(s2): sb.run()
print("2 x batch 32 on two streams ms (per pair):", timeit(both))
mc, sc = mk(16); md, sd = mk(16)
s3, s4 = torch.cuda.Stream(), torch.cuda.Stream()
