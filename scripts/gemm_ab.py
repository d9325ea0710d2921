"""A/B microbenchmark of the GEMM epilogue variants on the per-stage Poseidon-B shapes. Three buffer sets are rotated
so that no launch finds its streams in L2. Usage: SCOT_GEMM_ASYNC_EPI=0|1 python scripts/gemm_ab.py [tag]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseidon_b200 import _lib as L
dev = "cuda"
tag = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("SCOT_GEMM_ASYNC_EPI", "1")

def timeit(fn, n=30):
    for i in range(6): fn(i)
    torch.cuda.synchronize(); s = torch.cuda.Event(True); e = torch.cuda.Event(True); s.record()
    for i in range(n): fn(i)
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3

def run(M, C):
    H = 4 * C
    sets = []
    nset = 3 if M * H * 2 * 3 > 40e6 else 8
    for _ in range(nset):
        sets.append(dict(x=torch.randn(M, C, device=dev).bfloat16(), w1=torch.randn(H, C, device=dev).bfloat16() * 0.1,
                         w2=torch.randn(C, H, device=dev).bfloat16() * 0.1, wq=torch.randn(3 * C, C, device=dev).bfloat16() * 0.1,
                         b1=torch.randn(H, device=dev), bq=torch.randn(3 * C, device=dev),
                         h=torch.empty(M, H, device=dev, dtype=torch.bfloat16), g=torch.empty(M, H, device=dev, dtype=torch.bfloat16),
                         dh=torch.empty(M, H, device=dev, dtype=torch.bfloat16), qkv=torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16),
                         dz=torch.randn(M, C, device=dev).bfloat16(), do=torch.empty(M, C, device=dev, dtype=torch.bfloat16),
                         cs=torch.zeros(H, device=dev)))
    for s in sets:
        s["h"].normal_()
    r = {"M": M, "C": C, "tag": tag}
    r["qkv_bf16"] = timeit(lambda i: L.gemm(sets[i % nset]["x"], sets[i % nset]["wq"], M, 3 * C, C, mode=L.EPI_BF16, bias=sets[i % nset]["bq"], out0=sets[i % nset]["qkv"]))
    r["mlp1_gelu"] = timeit(lambda i: L.gemm(sets[i % nset]["x"], sets[i % nset]["w1"], M, H, C, mode=L.EPI_GELU, bias=sets[i % nset]["b1"], out0=sets[i % nset]["h"], out1=sets[i % nset]["g"]))
    r["mlp2_dgrad_gelubwd"] = timeit(lambda i: L.gemm(sets[i % nset]["dz"], sets[i % nset]["w2"], M, H, C, b_mn=True, mode=L.EPI_GELU_BWD, out0=sets[i % nset]["dh"], aux=sets[i % nset]["h"], colsum=sets[i % nset]["cs"]))
    r["proj_dgrad_bf16"] = timeit(lambda i: L.gemm(sets[i % nset]["dz"], sets[i % nset]["w2"][:, :C], M, C, C, b_mn=True, mode=L.EPI_BF16, out0=sets[i % nset]["do"]))
    mb = {"qkv_bf16": M * C * 2 + M * 3 * C * 2, "mlp1_gelu": M * C * 2 + 2 * M * H * 2, "mlp2_dgrad_gelubwd": M * C * 2 + 2 * M * H * 2,
          "proj_dgrad_bf16": 2 * M * C * 2}
    r["GBs"] = {k: round(mb[k] / r[k] / 1e3, 0) for k in mb}
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)

for M, C in [(65536, 96), (16384, 192), (4096, 384), (1024, 768)]:
    run(M, C)
