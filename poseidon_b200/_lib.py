"""ctypes binding of libscot_b200.so (the C ABI declared in include/scot_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised. The product path never routes through any CPU implementation.
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
    lib.scot_set_split_offset.argtypes = [C.c_size_t]
    lib.scot_set_split_offset.restype = None
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


class split_offset:
    """Context manager for the per-op entry points in "parity" (split-bf16) precision: inside it every bf16 tensor T
    passed to the library is the pair (T, tensor `nbytes` bytes after T); see scot_set_split_offset in scot_b200.h."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)

    def __enter__(self):
        load().scot_set_split_offset(self.nbytes)
        return self

    def __exit__(self, *exc):
        load().scot_set_split_offset(0)
        return False


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


# ---------------------------------------------------------------------------------------------------
# per-op wrappers (tests) and the whole-model engine
# ---------------------------------------------------------------------------------------------------
class ScotModelDesc(C.Structure):
    _fields_ = [
        ("image_size", C.c_int), ("patch_size", C.c_int), ("num_channels", C.c_int), ("num_out_channels", C.c_int),
        ("embed_dim", C.c_int), ("num_stages", C.c_int),
        ("depths", C.c_int * 4), ("num_heads", C.c_int * 4), ("skip_blocks", C.c_int * 4),
        ("window_size", C.c_int), ("mlp_ratio", C.c_float),
        ("use_conditioning", C.c_int), ("learn_residual", C.c_int), ("loss_p", C.c_int),
        ("n_slices", C.c_int), ("slices", C.c_int * 10), ("layer_norm_eps", C.c_float),
        ("precision", C.c_int),
    ]


class ScotLayerDesc(C.Structure):
    _fields_ = [("batch", C.c_int), ("res", C.c_int), ("C", C.c_int), ("heads", C.c_int), ("window", C.c_int), ("shift", C.c_int),
                ("mlp_ratio", C.c_float), ("use_conditioning", C.c_int), ("layer_norm_eps", C.c_float), ("precision", C.c_int)]


class ScotWgradProblem(C.Structure):
    _fields_ = [("dY", C.c_void_p), ("ld_dy", C.c_long), ("X", C.c_void_p), ("ld_x", C.c_long), ("dW", C.c_void_p),
                ("ld_dw", C.c_long), ("tokens", C.c_long), ("n_out", C.c_int), ("n_in", C.c_int)]


class ScotCpbLayer(C.Structure):
    _fields_ = [("w1", C.c_int), ("b1", C.c_int), ("w2", C.c_int), ("ls", C.c_int), ("tab2", C.c_int), ("alpha", C.c_int),
                ("dtab", C.c_int), ("dalpha", C.c_int), ("dpre", C.c_int), ("ws", C.c_short), ("heads", C.c_short)]


class ScotCpbTable(C.Structure):
    _fields_ = [("n", C.c_int), ("layer", ScotCpbLayer * 64)]


def _declare_engine(lib):
    vp, i, l, f = C.c_void_p, C.c_int, C.c_long, C.c_float
    lib.scot_gemm_wgrad_group.argtypes = [C.POINTER(ScotWgradProblem), i, i, vp]
    lib.scot_gemm_wgrad_group.restype = i
    lib.scot_cln_fwd.argtypes = [vp] * 11 + [l, i, i, i, f, vp]
    lib.scot_cln_fwd.restype = i
    lib.scot_cln_bwd.argtypes = [vp] * 7 + [i] + [vp] * 5 + [l, i, i, i, vp]
    lib.scot_cln_bwd.restype = i
    lib.scot_cpb_fwd.argtypes = [C.POINTER(ScotCpbTable), vp, vp, vp]
    lib.scot_cpb_fwd.restype = i
    lib.scot_cpb_bwd.argtypes = [C.POINTER(ScotCpbTable), vp, vp, vp, vp]
    lib.scot_cpb_bwd.restype = i
    lib.scot_attn_fwd.argtypes = [vp] * 5 + [i] * 6 + [vp]
    lib.scot_attn_fwd.restype = i
    lib.scot_attn_bwd_partial_bytes.argtypes = [i, i, i]
    lib.scot_attn_bwd_partial_bytes.restype = C.c_size_t
    lib.scot_attn_bwd.argtypes = [vp] * 8 + [C.c_size_t] + [vp] * 4 + [i] * 6 + [vp]
    lib.scot_attn_bwd.restype = i
    lib.scot_engine_create.argtypes = [C.POINTER(ScotModelDesc), i, C.POINTER(vp)]
    lib.scot_engine_create.restype = i
    lib.scot_engine_destroy.argtypes = [vp]
    lib.scot_engine_destroy.restype = None
    lib.scot_engine_num_params.argtypes = [vp]
    lib.scot_engine_num_params.restype = l
    lib.scot_engine_param_elems.argtypes = [vp]
    lib.scot_engine_param_elems.restype = l
    lib.scot_engine_param_info.argtypes = [vp, l, C.c_char_p, i, C.POINTER(l), C.POINTER(l), C.POINTER(i), C.POINTER(l)]
    lib.scot_engine_param_info.restype = i
    lib.scot_engine_workspace_bytes.argtypes = [vp]
    lib.scot_engine_workspace_bytes.restype = C.c_size_t
    lib.scot_engine_forward.argtypes = [vp] * 7 + [i, vp, vp, i, vp]
    lib.scot_engine_forward.restype = i
    lib.scot_engine_backward.argtypes = [vp] * 6 + [i, vp]
    lib.scot_engine_backward.restype = i
    lib.scot_engine_bind_io.argtypes = [vp] * 5 + [i, vp]
    lib.scot_engine_bind_io.restype = i
    lib.scot_engine_backward_part.argtypes = [vp] * 6 + [i, i, vp]
    lib.scot_engine_backward_part.restype = i
    lib.scot_engine_grad_split.argtypes = [vp]
    lib.scot_engine_grad_split.restype = l
    lib.scot_grad_sq_norm.argtypes = [vp, l, vp, vp]
    lib.scot_grad_sq_norm.restype = i
    lib.scot_adamw_step.argtypes = [vp] * 6 + [l, vp, i, vp, f, f, vp]
    lib.scot_adamw_step.restype = i
    lib.scot_lp_plane_sums.argtypes = [vp, vp, vp, i, l, l, vp]
    lib.scot_lp_plane_sums.restype = i
    lib.scot_layer_num_params.argtypes = [C.POINTER(ScotLayerDesc)]
    lib.scot_layer_num_params.restype = i
    lib.scot_layer_workspace_bytes.argtypes = [C.POINTER(ScotLayerDesc)]
    lib.scot_layer_workspace_bytes.restype = C.c_size_t
    lib.scot_layer_fwd.argtypes = [C.POINTER(ScotLayerDesc), C.POINTER(vp), vp, vp, vp, vp, C.c_size_t, vp]
    lib.scot_layer_fwd.restype = i
    lib.scot_layer_bwd.argtypes = [C.POINTER(ScotLayerDesc), C.POINTER(vp), vp, vp, vp, vp, C.c_size_t, vp]
    lib.scot_layer_bwd.restype = i
    # glue ops
    for name, args in (
        ("scot_cast_f32_bf16", [vp, vp, l, vp]),
        ("scot_embed_im2col", [vp, vp, i, i, i, i, i, vp]),
        ("scot_merge_gather", [vp, vp, vp, i, i, i, vp]),
        ("scot_merge_scatter", [vp, vp, vp, i, i, i, vp]),
        ("scot_convnext_dwconv7_fwd", [vp, vp, vp, vp, i, i, i, vp]),
        ("scot_convnext_dwconv7_bwd", [vp, vp, vp, vp, vp, vp, i, i, i, vp]),
        ("scot_convnext_scale_add_fwd", [vp, vp, vp, vp, vp, l, i, vp]),
        ("scot_convnext_scale_add_bwd", [vp, vp, vp, vp, vp, vp, l, i, vp]),
        ("scot_recovery_unshuffle", [vp, vp, i, i, i, i, i, vp]),
        ("scot_recovery_conv5_fwd", [vp, vp, vp, i, vp, vp, i, vp, i, i, i, i, vp]),
        ("scot_recovery_conv5_bwd", [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp]),
        ("scot_loss_fwd", [vp, vp, vp, vp, vp, i, i, i, i, l, vp]),
        ("scot_loss_bwd", [vp, vp, vp, vp, vp, vp, i, vp, vp, i, i, i, i, l, vp]),
    ):
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i


_declare_base = _declare


def _declare(lib):  # noqa: F811  (extends the GEMM-only declaration above)
    _declare_base(lib)
    _declare_engine(lib)


def cln_fwd(z, residual, time, aw, ab, cw, cb, x_out, xb_out, zhat, rstd, rows, Cdim, rows_per_sample, perm_res=0,
            eps=1e-5):
    check(load().scot_cln_fwd(ptr(z), ptr(residual), ptr(time), ptr(aw), ptr(ab), ptr(cw), ptr(cb), ptr(x_out),
                              ptr(xb_out), ptr(zhat), ptr(rstd), rows, Cdim, rows_per_sample, perm_res, eps, cur_stream()),
          "scot_cln_fwd")


def cln_bwd(dy, zhat, rstd, time, aw, ab, dz, dz_is_f32, g_aw, g_ab, g_cw, g_cb, g_bias_prev, rows, Cdim,
            rows_per_sample, perm_res=0):
    check(load().scot_cln_bwd(ptr(dy), ptr(zhat), ptr(rstd), ptr(time), ptr(aw), ptr(ab), ptr(dz), int(dz_is_f32),
                              ptr(g_aw), ptr(g_ab), ptr(g_cw), ptr(g_cb), ptr(g_bias_prev), rows, Cdim, rows_per_sample,
                              perm_res, cur_stream()), "scot_cln_bwd")


class CpbLayerBuffers:
    """Test helper: packs one attention layer's bias-MLP parameters / gradients / outputs into the flat
    parameter + arena form the batched C entry points expect."""

    def __init__(self, w1, b1, w2, ls, ws, heads):
        import torch

        dev = w1.device
        self.ws, self.heads = ws, heads
        R = (2 * ws - 1) ** 2
        sizes = [w1.numel(), b1.numel(), w2.numel(), ls.numel()]
        offs = [0]
        for n in sizes[:-1]:
            offs.append(offs[-1] + (n + 63) // 64 * 64)
        self.params = torch.zeros(offs[-1] + sizes[-1] + 64, device=dev)
        for o, t in zip(offs, (w1, b1, w2, ls)):
            self.params[o:o + t.numel()] = t.reshape(-1).float()
        self.grads = torch.zeros_like(self.params)
        self.offs = offs
        unit = lambda n: (n * 4 + 255) // 256  # noqa: E731
        aoff, cur = [], 0
        for n in (R * heads, heads, R * heads, heads, R * heads):
            aoff.append(cur)
            cur += unit(n)
        self.arena = torch.zeros(cur * 64 + 64, device=dev)  # floats; 256 B units = 64 floats
        shift = (-self.arena.data_ptr()) % 256 // 4
        self.arena = self.arena[shift:]
        self.aoff = aoff
        self.R = R
        t = ScotCpbTable()
        t.n = 1
        L_ = t.layer[0]
        L_.w1, L_.b1, L_.w2, L_.ls = offs
        L_.tab2, L_.alpha, L_.dtab, L_.dalpha, L_.dpre = aoff
        L_.ws, L_.heads = ws, heads
        self.table = t

    def view(self, k, n):
        return self.arena[self.aoff[k] * 64: self.aoff[k] * 64 + n]

    @property
    def tab2(self):
        return self.view(0, self.R * self.heads).view(self.R, self.heads)

    @property
    def alpha(self):
        return self.view(1, self.heads)

    @property
    def dtab(self):
        return self.view(2, self.R * self.heads).view(self.R, self.heads)

    @property
    def dalpha(self):
        return self.view(3, self.heads)

    def grad(self, k, shape):
        n = 1
        for s_ in shape:
            n *= s_
        return self.grads[self.offs[k]: self.offs[k] + n].view(shape)

    def forward(self):
        check(load().scot_cpb_fwd(C.byref(self.table), ptr(self.params), ptr(self.arena), cur_stream()), "scot_cpb_fwd")

    def backward(self):
        check(load().scot_cpb_bwd(C.byref(self.table), ptr(self.params), ptr(self.grads), ptr(self.arena), cur_stream()),
              "scot_cpb_bwd")


def wgrad_group(problems, impl=GEMM_TCGEN05):
    """problems: list of (dY [tokens, n_out] bf16, X [tokens, n_in] bf16, dW [n_out, n_in] f32); dW += dY^T X, one launch"""
    arr = (ScotWgradProblem * len(problems))()
    for k, (dY, X, dW) in enumerate(problems):
        arr[k] = ScotWgradProblem(dY.data_ptr(), dY.stride(0), X.data_ptr(), X.stride(0), dW.data_ptr(), dW.stride(0),
                                  dY.shape[0], dY.shape[1], X.shape[1])
    check(load().scot_gemm_wgrad_group(arr, len(problems), impl, cur_stream()), "scot_gemm_wgrad_group")


def attn_fwd(qkv, out, lse, tab2, alpha, batch, res, ws, shift, heads, hd):
    check(load().scot_attn_fwd(ptr(qkv), ptr(out), ptr(lse), ptr(tab2), ptr(alpha), batch, res, ws, shift, heads, hd,
                               cur_stream()), "scot_attn_fwd")


def attn_bwd(qkv, o, d_o, lse, tab2, alpha, dqkv, partial, dtab, dalpha, g_qbias, g_vbias, batch, res, ws, shift, heads,
             hd):
    check(load().scot_attn_bwd(ptr(qkv), ptr(o), ptr(d_o), ptr(lse), ptr(tab2), ptr(alpha), ptr(dqkv), ptr(partial),
                               partial.numel() * partial.element_size(), ptr(dtab), ptr(dalpha), ptr(g_qbias),
                               ptr(g_vbias), batch, res, ws, shift, heads, hd, cur_stream()), "scot_attn_bwd")


class Layer:
    """Stand-alone ScOTLayer through scot_layer_fwd / scot_layer_bwd. `params`: list of fp32 CUDA tensors in the reference's
    state_dict order of a layer (see scot_b200.h)."""

    def __init__(self, batch, res, Cdim, heads, window, shift, mlp_ratio=4.0, use_conditioning=True, eps=1e-5, precision=0):
        import torch

        self.desc = ScotLayerDesc(batch, res, Cdim, heads, window, shift, mlp_ratio, int(use_conditioning), eps, precision)
        lib = load()
        self.n = lib.scot_layer_num_params(C.byref(self.desc))
        nbytes = lib.scot_layer_workspace_bytes(C.byref(self.desc))
        if nbytes == 0:
            check(1, "scot_layer_workspace_bytes")
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
        shift_ = (-ws.data_ptr()) % 256
        self.ws = ws[shift_:shift_ + nbytes]
        self.rows, self.C = batch * res * res, Cdim

    def forward(self, params, x, time):
        import torch

        assert len(params) == self.n
        arr = (C.c_void_p * self.n)(*[p.data_ptr() for p in params])
        y = torch.empty(self.rows, self.C, device=x.device)
        check(load().scot_layer_fwd(C.byref(self.desc), arr, ptr(x), ptr(time), ptr(y), ptr(self.ws), self.ws.numel(),
                                    cur_stream()), "scot_layer_fwd")
        return y

    def backward(self, grads, time, dy):
        import torch

        arr = (C.c_void_p * self.n)(*[g.data_ptr() if g is not None else None for g in grads])
        dx = torch.empty(self.rows, self.C, device=dy.device)
        check(load().scot_layer_bwd(C.byref(self.desc), arr, ptr(time), ptr(dy), ptr(dx), ptr(self.ws), self.ws.numel(),
                                    cur_stream()), "scot_layer_bwd")
        return dx


def glue(name, *args):
    """Calls one of the glue entry points (scot_embed_im2col, scot_merge_gather, ... see scot_b200.h): tensors are passed
    as torch CUDA tensors (or None), scalars as ints; the current stream is appended."""
    import torch

    conv = [ptr(a) if (a is None or isinstance(a, torch.Tensor)) else a for a in args]
    check(getattr(load(), name)(*conv, cur_stream()), name)


class Engine:
    """Host handle of the native engine for one (config, batch) pair. Owns no device memory."""

    def __init__(self, desc: ScotModelDesc, batch: int):
        lib = load()
        self._lib = lib
        h = C.c_void_p()
        check(lib.scot_engine_create(C.byref(desc), batch, C.byref(h)), "scot_engine_create")
        self.handle = h
        self.batch = batch
        self.param_elems = lib.scot_engine_param_elems(h)
        self.workspace_bytes = lib.scot_engine_workspace_bytes(h)
        self.table = {}
        buf = C.create_string_buffer(512)
        off, num, nd = C.c_long(), C.c_long(), C.c_int()
        shp = (C.c_long * 4)()
        for i in range(lib.scot_engine_num_params(h)):
            check(lib.scot_engine_param_info(h, i, buf, 512, C.byref(off), C.byref(num), C.byref(nd), shp))
            self.table[buf.value.decode()] = (off.value, num.value, tuple(shp[k] for k in range(nd.value)))

    def __del__(self):
        try:
            if self.handle:
                self._lib.scot_engine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def forward(self, params, arena, pixel_values, time, labels, mask, mask_mode, pred, loss, impl=GEMM_TCGEN05):
        check(self._lib.scot_engine_forward(self.handle, ptr(params), ptr(arena), ptr(pixel_values), ptr(time),
                                            ptr(labels), ptr(mask), mask_mode, ptr(pred), ptr(loss), impl, cur_stream()),
              "scot_engine_forward")

    def bind_io(self, pixel_values, time, labels, mask, mask_mode, pred):
        check(self._lib.scot_engine_bind_io(self.handle, ptr(pixel_values), ptr(time), ptr(labels), ptr(mask), mask_mode,
                                            ptr(pred)), "scot_engine_bind_io")

    def backward(self, params, grads, arena, grad_loss, grad_pred, impl=GEMM_TCGEN05, part=0):
        """part 0: whole backward; 1 / 2: the two halves of scot_engine_backward_part (gradient elements
        [grad_split, end) are final after part 1)"""
        check(self._lib.scot_engine_backward_part(self.handle, ptr(params), ptr(grads), ptr(arena), ptr(grad_loss),
                                                  ptr(grad_pred), impl, part, cur_stream()), "scot_engine_backward")

    @property
    def grad_split(self) -> int:
        return int(self._lib.scot_engine_grad_split(self.handle))
