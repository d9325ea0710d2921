import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseidon_b200 import _lib as L
dev = "cuda"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s = torch.cuda.Event(True); e = torch.cuda.Event(True); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
M, N, K = 65536, 384, 96
A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); Bt = torch.randn(K, N, device=dev).bfloat16()
bias = torch.randn(N, device=dev)
big = torch.empty(3 * M * N * 2 + (64 << 20), dtype=torch.uint8, device=dev)
base = (-big.data_ptr()) % (2 << 20)
def view(off):
    return big[base + off: base + off + M * N * 2].view(torch.bfloat16).view(M, N)
cs = torch.zeros(N, device=dev)
for pad in [0, 256, 1024, 4096, 65536, (1 << 20) + 4096 + 256, (2 << 20), (3 << 20) + 8192 + 512]:
    o0 = view(0); o1 = view(M * N * 2 + pad)
    g = timeit(lambda: L.gemm(A, B, M, N, K, mode=L.EPI_GELU, bias=bias, out0=o0, out1=o1))
    gb = timeit(lambda: L.gemm(A, Bt, M, N, K, b_mn=True, mode=L.EPI_GELU_BWD, out0=o0, aux=o1, colsum=cs))
    print(json.dumps({"pad": pad, "gelu_us": round(g, 1), "gelu_bwd_us": round(gb, 1), "delta_bytes": M * N * 2 + pad}), flush=True)
