"""Round-2 bring-up of gemm_async_epi2_kernel (SCOT_GEMM_ASYNC_V2=1): bit-compare against the validated v1 kernel on
the bf16-output epilogue modes (v1 = gemm_async_epi_kernel) and the fp32-output modes (v1 = gemm_tc_kernel) and time both. Run under a timeout — v2 has never executed on hardware:

    timeout 180 python scripts/gemm_v2_check.py

The knob is read per launch, so one process can alternate between the two kernels. Expected: outputs bit-identical
(same arithmetic, only the order of waits / the staging buffers differ), bias-gradient column sums equal up to the
order of the fp32 atomics.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseidon_b200 import _lib as L  # noqa: E402

dev = "cuda"
# SCOT_GEMM_ASYNC_V2 bit mask: 0 = validated kernels, 1 = v2, 3 = v2 + two-group GELU kernel, 5 = v2 + small-K kernel
# (resident weight tile, 32-wide K tail with 64-byte swizzle; only the K = 96 shapes below take that path)
LEVELS = (0, 1, 3, 5)
SHAPES = [(65536, 384, 96), (16384, 768, 192), (4096, 1536, 384), (1024, 3072, 768), (1000, 192, 96), (65536, 288, 96)]


def run(v2, fn):
    os.environ["SCOT_GEMM_ASYNC_V2"] = str(int(v2))
    fn()
    torch.cuda.synchronize()


def timeit(v2, fn, n=20):
    os.environ["SCOT_GEMM_ASYNC_V2"] = str(int(v2))
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
    torch.manual_seed(0)
    report = []
    for M, N, K in SHAPES:
        A = torch.randn(M, K, device=dev).bfloat16()
        B = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        Bt = (torch.randn(K, N, device=dev) / K ** 0.5).bfloat16()  # MN-major B operand of a dgrad
        bias = torch.randn(N, device=dev)
        aux = torch.randn(M, N, device=dev).bfloat16()
        acc0 = torch.randn(M, N, device=dev)  # initial content of the `+=` target
        outs = {}
        for v2 in LEVELS:
            o_bf = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            o_bfT = torch.zeros_like(o_bf)
            o_g0, o_g1 = torch.zeros_like(o_bf), torch.zeros_like(o_bf)
            o_bw = torch.zeros_like(o_bf)
            cs = torch.zeros(N, device=dev)
            run(v2, lambda: L.gemm(A, B, M, N, K, mode=L.EPI_BF16, bias=bias, out0=o_bf))
            run(v2, lambda: L.gemm(A, Bt, M, N, K, b_mn=True, mode=L.EPI_BF16, out0=o_bfT))
            run(v2, lambda: L.gemm(A, B, M, N, K, mode=L.EPI_GELU, bias=bias, out0=o_g0, out1=o_g1))
            run(v2, lambda: L.gemm(A, Bt, M, N, K, b_mn=True, mode=L.EPI_GELU_BWD, out0=o_bw, aux=aux, colsum=cs))
            o_f = torch.zeros(M, N, device=dev)
            o_rmw = acc0.clone()
            run(v2, lambda: L.gemm(A, B, M, N, K, mode=L.EPI_F32, bias=bias, out0=o_f))
            run(v2, lambda: L.gemm(A, Bt, M, N, K, b_mn=True, mode=L.EPI_RMW_F32, out0=o_rmw))
            t = {
                "bf16": timeit(v2, lambda: L.gemm(A, B, M, N, K, mode=L.EPI_BF16, bias=bias, out0=o_bf)),
                "gelu": timeit(v2, lambda: L.gemm(A, B, M, N, K, mode=L.EPI_GELU, bias=bias, out0=o_g0, out1=o_g1)),
                "gelu_bwd": timeit(v2, lambda: L.gemm(A, Bt, M, N, K, b_mn=True, mode=L.EPI_GELU_BWD, out0=o_bw, aux=aux)),
                "f32": timeit(v2, lambda: L.gemm(A, B, M, N, K, mode=L.EPI_F32, bias=bias, out0=o_f)),
                "rmw": timeit(v2, lambda: L.gemm(A, Bt, M, N, K, b_mn=True, mode=L.EPI_RMW_F32, out0=torch.empty_like(o_rmw).zero_())),
            }
            outs[v2] = (o_bf, o_bfT, o_g0, o_g1, o_bw, o_f, o_rmw, cs, t)
        a = outs[0]
        names = ("bf16", "bf16_T", "gelu_d", "gelu", "gelu_bwd", "f32", "rmw")
        rec = {"shape": [M, N, K], "us_v1": {k: round(v, 2) for k, v in a[8].items()}}
        for lvl in LEVELS[1:]:
            b = outs[lvl]
            rec[f"equal_l{lvl}"] = {n: bool(torch.equal(x, y)) for n, x, y in zip(names, a[:7], b[:7])}
            rec[f"colsum_rel_l{lvl}"] = float((a[7] - b[7]).norm() / (a[7].norm() + 1e-30))
            rec[f"us_l{lvl}"] = {k: round(v, 2) for k, v in b[8].items()}
        print(json.dumps(rec), flush=True)
        report.append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(report, open("gpurun_out/gemm_v2_check.json", "w"), indent=1)


if __name__ == "__main__":
    main()
