"""GPU bring-up check for the tcgen05 GEMM: every (major, epilogue) combination over scOT's shapes
against torch.matmul on the same bf16 inputs. Writes gpurun_out/gemm_bringup.json."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseidon_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"
res = []

def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

def run(name, fn):
    t0 = time.time()
    try:
        r = fn()
        torch.cuda.synchronize()
        r = dict(r); r["name"] = name; r["ok"] = bool(r.get("err", 1) < r.get("tol", 1e-2))
    except Exception as ex:  # noqa
        r = {"name": name, "ok": False, "exc": repr(ex)[:500]}
    r["sec"] = round(time.time() - t0, 3)
    res.append(r); print(json.dumps(r), flush=True)

def fwd_case(M, N, K, impl, mode=L.EPI_BF16):
    def f():
        A = torch.randn(M, K, device=dev).bfloat16(); B = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        bias = torch.randn(N, device=dev)
        ref = A.float() @ B.float().t() + bias
        if mode == L.EPI_BF16:
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            L.gemm(A, B, M, N, K, mode=mode, bias=bias, out0=out, impl=impl)
            return {"err": rel(out.float(), ref), "tol": 6e-3}
        if mode == L.EPI_F32:
            out = torch.empty(M, N, device=dev)
            L.gemm(A, B, M, N, K, mode=mode, bias=bias, out0=out, impl=impl)
            return {"err": rel(out, ref), "tol": 1e-4}
        if mode == L.EPI_GELU:
            h = torch.empty(M, N, device=dev, dtype=torch.bfloat16); g = torch.empty_like(h)
            L.gemm(A, B, M, N, K, mode=mode, bias=bias, out0=h, out1=g, impl=impl)
            gref = torch.nn.functional.gelu(h.float())
            return {"err": max(rel(h.float(), ref), rel(g.float(), gref)), "tol": 6e-3}
        if mode == L.EPI_ADD_F32_BF16:
            res_ = torch.randn(M, N, device=dev); o = torch.empty(M, N, device=dev); ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            L.gemm(A, B, M, N, K, mode=mode, bias=bias, out0=o, out1=ob, aux=res_, impl=impl)
            return {"err": max(rel(o, ref + res_), rel(ob.float(), ref + res_) / 40), "tol": 1e-4}
    return f

def dgrad_case(M, N, K, impl, mode):
    # dX[M,N] = dY[M,K] @ W[K,N]   (W stored [K(out feats of fwd), N(in feats)]; B operand MN-major)
    def f():
        dY = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(K, N, device=dev) / K ** 0.5).bfloat16()
        ref = dY.float() @ W.float()
        if mode == L.EPI_BF16:
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            L.gemm(dY, W, M, N, K, b_mn=True, mode=mode, out0=out, impl=impl)
            return {"err": rel(out.float(), ref), "tol": 6e-3}
        if mode == L.EPI_RMW_F32:
            g = torch.randn(M, N, device=dev); g0 = g.clone()
            L.gemm(dY, W, M, N, K, b_mn=True, mode=mode, out0=g, impl=impl)
            return {"err": rel(g, g0 + ref), "tol": 1e-4}
        if mode == L.EPI_GELU_BWD:
            h = torch.randn(M, N, device=dev).bfloat16()
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); cs = torch.zeros(N, device=dev)
            L.gemm(dY, W, M, N, K, b_mn=True, mode=mode, out0=out, aux=h, colsum=cs, impl=impl)
            hh = h.float().requires_grad_(True); torch.nn.functional.gelu(hh).backward(ref)
            return {"err": max(rel(out.float(), hh.grad), rel(cs, out.float().sum(0))), "tol": 6e-3}
    return f

def wgrad_case(Mtok, Nw, Kw, impl, mode):
    # dW[Nw,Kw] = dY[Mtok,Nw]^T @ X[Mtok,Kw]
    def f():
        dY = torch.randn(Mtok, Nw, device=dev).bfloat16(); X = torch.randn(Mtok, Kw, device=dev).bfloat16()
        ref = dY.float().t() @ X.float()
        out = torch.zeros(Nw, Kw, device=dev) if mode == L.EPI_ATOMIC_F32 else torch.empty(Nw, Kw, device=dev)
        L.gemm(dY, X, Nw, Kw, Mtok, a_mn=True, b_mn=True, mode=mode, out0=out, impl=impl)
        return {"err": rel(out, ref), "tol": 1e-4}
    return f

only_simt = "--simt-only" in sys.argv
impls = [("simt", L.GEMM_SIMT)] + ([] if only_simt else [("tc", L.GEMM_TCGEN05)])
for iname, impl in impls:
    # smallest first: a single tile, single k-block
    run(f"{iname}.fwd.128x128x64", fwd_case(128, 128, 64, impl))
    run(f"{iname}.fwd.128x128x64.f32", fwd_case(128, 128, 64, impl, L.EPI_F32))
    run(f"{iname}.fwd.256x96x96.f32", fwd_case(256, 96, 96, impl, L.EPI_F32))
    run(f"{iname}.fwd.1024x288x96", fwd_case(1024, 288, 96, impl))
    run(f"{iname}.fwd.1000x64x48", fwd_case(1000, 64, 48, impl))
    run(f"{iname}.fwd.4096x384x96.gelu", fwd_case(4096, 384, 96, impl, L.EPI_GELU))
    run(f"{iname}.fwd.4096x96x384.addres", fwd_case(4096, 96, 384, impl, L.EPI_ADD_F32_BF16))
    run(f"{iname}.fwd.1024x2304x768", fwd_case(1024, 2304, 768, impl))
    run(f"{iname}.fwd.1024x768x3072.f32", fwd_case(1024, 768, 3072, impl, L.EPI_F32))
    run(f"{iname}.dgrad.128x128x64", dgrad_case(128, 128, 64, impl, L.EPI_BF16))
    run(f"{iname}.dgrad.4096x96x288", dgrad_case(4096, 96, 288, impl, L.EPI_BF16))
    run(f"{iname}.dgrad.4096x96x288.rmw", dgrad_case(4096, 96, 288, impl, L.EPI_RMW_F32))
    run(f"{iname}.dgrad.4096x384x96.gelubwd", dgrad_case(4096, 384, 96, impl, L.EPI_GELU_BWD))
    run(f"{iname}.dgrad.1024x768x2304", dgrad_case(1024, 768, 2304, impl, L.EPI_BF16))
    run(f"{iname}.wgrad.128x128x64.f32", wgrad_case(64, 128, 128, impl, L.EPI_F32))
    run(f"{iname}.wgrad.4096tok.288x96.f32", wgrad_case(4096, 288, 96, impl, L.EPI_F32))
    run(f"{iname}.wgrad.65536tok.288x96.atomic", wgrad_case(65536, 288, 96, impl, L.EPI_ATOMIC_F32))
    run(f"{iname}.wgrad.1024tok.2304x768.atomic", wgrad_case(1024, 2304, 768, impl, L.EPI_ATOMIC_F32))

# timing of the big shapes (tcgen05 only)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s = torch.cuda.Event(True); e = torch.cuda.Event(True); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
if not only_simt and all(r["ok"] for r in res if r["name"].startswith("tc.")):
    for (M, N, K) in [(65536, 288, 96), (65536, 384, 96), (65536, 96, 384), (16384, 576, 192), (4096, 1152, 384), (1024, 2304, 768), (1024, 3072, 768), (8192, 8192, 8192)]:
        A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: L.gemm(A, B, M, N, K, out0=out))
        ms_t = timeit(lambda: torch.matmul(A, B.t()))
        r = {"name": f"time.fwd.{M}x{N}x{K}", "ms": ms, "torch_ms": ms_t, "tflops": 2 * M * N * K / ms / 1e9, "gbs": (M * K + N * K + M * N) * 2 / ms / 1e6, "ok": True}
        res.append(r); print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_bringup.json", "w"), indent=1)
bad = [r["name"] for r in res if not r["ok"]]
print("FAILED:", bad if bad else "none")
