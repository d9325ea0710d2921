"""Launches the LayerNorm backward kernel once per Poseidon-B stage shape (for ncu captures / timing)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseidon_b200 import _lib as L
dev = "cuda"
B = 64
def run(M, C, rps, reps=3):
    dy = torch.randn(M, C, device=dev); zh = torch.randn(M, C, device=dev).bfloat16(); rstd = torch.rand(M, device=dev) + 0.5
    t = torch.rand(B, device=dev); aw = torch.randn(C, device=dev); ab = torch.randn(C, device=dev)
    dz = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    g = [torch.zeros(C, device=dev) for _ in range(5)]
    for _ in range(reps):
        L.cln_bwd(dy, zh, rstd, t, aw, ab, dz, False, g[0], g[1], g[2], g[3], g[4], M, C, rps)
    torch.cuda.synchronize()
    s = torch.cuda.Event(True); e = torch.cuda.Event(True); s.record()
    for _ in range(20):
        L.cln_bwd(dy, zh, rstd, t, aw, ab, dz, False, g[0], g[1], g[2], g[3], g[4], M, C, rps)
    e.record(); torch.cuda.synchronize()
    print(M, C, "us/launch (incl. host launch gaps)", round(s.elapsed_time(e) / 20 * 1e3, 1), flush=True)
for M, C, rps in [(65536, 96, 1024), (16384, 192, 256), (4096, 384, 64), (1024, 768, 16)]:
    run(M, C, rps)
