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
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def run(M, N, K):
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); Bt = torch.randn(K, N, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16); ob2 = torch.empty_like(ob); of = torch.empty(M, N, device=dev)
    cs = torch.zeros(N, device=dev)
    res = {}
    res["bf16"] = timeit(lambda: L.gemm(A, B, M, N, K, mode=L.EPI_BF16, bias=bias, out0=ob))
    res["f32"] = timeit(lambda: L.gemm(A, B, M, N, K, mode=L.EPI_F32, bias=bias, out0=of))
    res["gelu"] = timeit(lambda: L.gemm(A, B, M, N, K, mode=L.EPI_GELU, bias=bias, out0=ob, out1=ob2))
    res["dgrad_bf16"] = timeit(lambda: L.gemm(A, Bt, M, N, K, b_mn=True, mode=L.EPI_BF16, out0=ob))
    res["dgrad_gelubwd"] = timeit(lambda: L.gemm(A, Bt, M, N, K, b_mn=True, mode=L.EPI_GELU_BWD, out0=ob, aux=ob2, colsum=cs))
    res["dgrad_rmw"] = timeit(lambda: L.gemm(A, Bt, M, N, K, b_mn=True, mode=L.EPI_RMW_F32, out0=of))
    res["torch_mm"] = timeit(lambda: torch.matmul(A, B.t()))
    mb = lambda *t: sum(x.numel() * x.element_size() for x in t) / 1e6
    res["MB_bf16"] = mb(A, B, ob)
    print(json.dumps({"shape": [M, N, K], **{k: round(v, 1) for k, v in res.items()}}), flush=True)
for shp in [(65536, 384, 96), (65536, 96, 384), (65536, 288, 96), (65536, 96, 96), (16384, 768, 192), (4096, 1536, 384), (1024, 3072, 768), (8192, 8192, 1024)]:
    run(*shp)
# wgrad
for (tok, Nw, Kw) in [(65536, 288, 96), (65536, 384, 96), (16384, 768, 192), (1024, 3072, 768)]:
    dY = torch.randn(tok, Nw, device=dev).bfloat16(); X = torch.randn(tok, Kw, device=dev).bfloat16(); out = torch.zeros(Nw, Kw, device=dev)
    us = timeit(lambda: L.gemm(dY, X, Nw, Kw, tok, a_mn=True, b_mn=True, mode=L.EPI_ATOMIC_F32, out0=out))
    print(json.dumps({"wgrad": [tok, Nw, Kw], "us": round(us, 1), "MB": (dY.numel() + X.numel()) * 2 / 1e6}), flush=True)
