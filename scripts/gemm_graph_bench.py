"""Device-side timing of the epilogue-fused GEMMs at the Poseidon-B stage shapes: 24 launches cycling over 3 buffer sets
(> L2) captured in one CUDA graph, so that the host launch rate (~15 us per ctypes call) does not bound the measurement."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseidon_b200 import _lib as L
dev = "cuda"
NSET, NL = 3, 24
def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(NSET): fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(NL): fn(i % NSET)
        for _ in range(3): g.replay()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * NL) * 1e3
res = {}
for (M, N, K) in [(65536, 288, 96), (65536, 384, 96), (16384, 576, 192), (16384, 768, 192), (4096, 1536, 384), (1024, 3072, 768)]:
    S = [dict(A=torch.randn(M, K, device=dev).bfloat16(), ob=torch.empty(M, N, device=dev, dtype=torch.bfloat16),
              ob2=torch.empty(M, N, device=dev, dtype=torch.bfloat16), of=torch.empty(M, K, device=dev),
              dz=torch.randn(M, N, device=dev).bfloat16()) for _ in range(NSET)]
    B = torch.randn(N, K, device=dev).bfloat16(); Bt = torch.randn(K, N, device=dev).bfloat16(); bias = torch.randn(N, device=dev)
    cs = torch.zeros(N, device=dev)
    r = {}
    r["bf16"] = graph_time(lambda i: L.gemm(S[i]["A"], B, M, N, K, mode=L.EPI_BF16, bias=bias, out0=S[i]["ob"]))
    r["gelu"] = graph_time(lambda i: L.gemm(S[i]["A"], B, M, N, K, mode=L.EPI_GELU, bias=bias, out0=S[i]["ob"], out1=S[i]["ob2"]))
    r["gelu_bwd"] = graph_time(lambda i: L.gemm(S[i]["A"], Bt, M, N, K, b_mn=True, mode=L.EPI_GELU_BWD, out0=S[i]["ob"], aux=S[i]["ob2"], colsum=cs))
    # the reverse direction (N -> K): fc2 / dgrad with fp32 outputs
    Bk = torch.randn(K, N, device=dev).bfloat16()
    r["f32_rev"] = graph_time(lambda i: L.gemm(S[i]["dz"], Bk, M, K, N, mode=L.EPI_F32, bias=None, out0=S[i]["of"]))
    r["rmw_rev"] = graph_time(lambda i: L.gemm(S[i]["dz"], B, M, K, N, b_mn=True, mode=L.EPI_RMW_F32, out0=S[i]["of"]))
    res[f"{M}x{N}x{K}"] = {k: round(v, 1) for k, v in r.items()}
    print(f"{M}x{N}x{K}", json.dumps(res[f"{M}x{N}x{K}"]), flush=True)
