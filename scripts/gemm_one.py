import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseidon_b200 import _lib as L
dev = "cuda"
M, N, K = 65536, 384, 96
A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); Bt = torch.randn(K, N, device=dev).bfloat16()
bias = torch.randn(N, device=dev)
ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16); ob2 = torch.empty_like(ob); of = torch.empty(M, N, device=dev); cs = torch.zeros(N, device=dev)
for _ in range(2):
    L.gemm(A, B, M, N, K, mode=L.EPI_BF16, bias=bias, out0=ob)
    L.gemm(A, B, M, N, K, mode=L.EPI_GELU, bias=bias, out0=ob, out1=ob2)
    L.gemm(A, Bt, M, N, K, b_mn=True, mode=L.EPI_GELU_BWD, out0=ob, aux=ob2, colsum=cs)
    L.gemm(A, Bt, M, N, K, b_mn=True, mode=L.EPI_RMW_F32, out0=of)
torch.cuda.synchronize()
