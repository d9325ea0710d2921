"""ctypes binding of libscot_b200.so (the C ABI declared in include/scot_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised. The product path never routes through the oracle or any CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscot_b200.so")

EPI_BF16, EPI_F32, EPI_GELU, EPI_GELU_BWD, EPI_RMW_F32, EPI_ATOMIC_F32, EPI_ADD_F32_BF16 = range(7)
GEMM_TCGEN05, GEMM_SIMT = 0, 1


class ScotEpilogue(C.Structure):
    _fields_ = [
        ("mode", C.c_int),
        ("bias", C.c_void_p),
        ("out0", C.c_void_p),
        ("ld0", C.c_long),
        ("out1", C.c_void_p),
        ("ld1", C.c_long),
        ("aux", C.c_void_p),
        ("ldaux", C.c_long),
        ("colsum", C.c_void_p),
    ]


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Loads the engine library, building it with nvcc first if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m poseidon_b200.build`")
        from . import build as _build

        _build.build()
    lib = C.CDLL(LIB_PATH)
    lib.scot_abi_version.restype = C.c_int
    lib.scot_last_error.restype = C.c_char_p
    lib.scot_launch_count.restype = C.c_ulonglong
    _declare(lib)
    _lib = lib
    return lib


def _declare(lib):
    vp, i, l = C.c_void_p, C.c_int, C.c_long
    lib.scot_gemm_bf16.argtypes = [vp, l, i, vp, l, i, i, i, i, C.POINTER(ScotEpilogue), i, vp]
    lib.scot_gemm_bf16.restype = i


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().scot_last_error().decode(errors="replace")
        raise RuntimeError(f"libscot_b200 {what} failed (rc={rc}): {msg}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(A, B, M, N, K, *, a_mn=False, b_mn=False, mode=EPI_BF16, bias=None, out0=None, out1=None, aux=None,
         colsum=None, impl=GEMM_TCGEN05):
    """D[m,n] = sum_k A(m,k) B(n,k); see scot_b200.h. Tensors are 2-D row-major torch CUDA tensors."""
    lib = load()
    e = ScotEpilogue(mode, ptr(bias), ptr(out0), out0.stride(0) if out0 is not None else 0, ptr(out1),
                     out1.stride(0) if out1 is not None else 0, ptr(aux), aux.stride(0) if aux is not None else 0,
                     ptr(colsum))
    rc = lib.scot_gemm_bf16(ptr(A), A.stride(0), int(a_mn), ptr(B), B.stride(0), int(b_mn), M, N, K, C.byref(e),
                            impl, cur_stream())
    check(rc, "scot_gemm_bf16")
