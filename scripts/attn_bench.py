"""Times the window-attention kernels at the Poseidon-B batch-64 stage shapes (CUDA events, L2 larger than the tensors is
not flushed: the qkv tensor alone is 37 MB at stage 0, results are for shares / A-B only).
    python scripts/attn_bench.py [fwd|bwd|both]
"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseidon_b200 import _lib as L  # noqa: E402

dev = "cuda"
CASES = [(64, 32, 16, 0, 3, 32), (64, 32, 16, 8, 3, 32), (64, 16, 16, 0, 6, 32), (64, 8, 8, 0, 12, 32), (64, 4, 4, 0, 24, 32),
         (32, 32, 16, 8, 3, 64), (32, 32, 16, 0, 3, 16)]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "both"
    torch.manual_seed(0)
    for case in CASES:
        Bn, res, ws, shift, heads, hd = case
        C = heads * hd
        M = Bn * res * res
        qkv = (torch.randn(M, 3 * C, device=dev) * 1.5).bfloat16()
        w1 = torch.randn(512, 2, device=dev)
        b1 = torch.randn(512, device=dev) * 0.1
        w2 = torch.randn(heads, 512, device=dev) / 512 ** 0.5
        ls = math.log(10.0) + 0.3 * torch.randn(heads, 1, 1, device=dev)
        cpb = L.CpbLayerBuffers(w1, b1, w2, ls, ws, heads)
        cpb.forward()
        nwin = Bn * (res // ws) ** 2
        out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(nwin * heads, ws * ws, device=dev)
        rec = {"case": case}
        fwd = lambda: L.attn_fwd(qkv, out, lse, cpb.tab2, cpb.alpha, Bn, res, ws, shift, heads, hd)
        fwd()
        if what in ("fwd", "both"):
            rec["fwd_us"] = round(timeit(fwd), 2)
        if what in ("bwd", "both"):
            d_o = torch.randn(M, C, device=dev).bfloat16()
            dqkv = torch.zeros(M, 3 * C, device=dev, dtype=torch.bfloat16)
            partial = torch.zeros(64, device=dev)
            gq, gv = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
            bwd = lambda: L.attn_bwd(qkv, out, d_o, lse, cpb.tab2, cpb.alpha, dqkv, partial, cpb.dtab, cpb.dalpha, gq, gv, Bn, res,
                                     ws, shift, heads, hd)
            rec["bwd_us"] = round(timeit(bwd), 2)
        units = nwin * heads
        N = ws * ws
        rec["fwd_gflop"] = round(4 * N * N * hd * units / 1e9, 3)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
