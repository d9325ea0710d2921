"""Drop-in replacement for the reference's `scOT/model.py` (camlab-ethz/poseidon) on B200.

Same public surface — `ScOTConfig`, `ScOT`, `ScOTOutput`, `LayerNorm`, `ConditionalLayerNorm`, identical
module tree / parameter names / HF checkpoint layout (reference scOT/model.py:57-160, 1243-1282), so
`ScOT.from_pretrained()`, `save_pretrained()` and the reference's `scOT/trainer.py` work on it unchanged —
but `ScOT.forward` (reference :1318-1509) and its backward run entirely inside the hand-written sm_100a
engine `libscot_b200.so` (include/scot_b200.h) through one `torch.autograd.Function`.

PyTorch is used here for what it is good at: owning device memory (parameters live as views of ONE
flat fp32 buffer, gradients as views of one flat buffer -> a single NCCL all-reduce per step), streams
and the HF plumbing. There is no CPU / eager fallback: without a CUDA device `forward` raises.
"""
from __future__ import annotations

import collections
import math
from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import torch
from torch import nn
from transformers import PretrainedConfig, PreTrainedModel
from transformers.utils import ModelOutput

from .. import _lib


@dataclass
class ScOTOutput(ModelOutput):
    """reference scOT/model.py:57-63"""

    loss: Optional[torch.FloatTensor] = None
    output: torch.FloatTensor = None
    hidden_states: Optional[Tuple[torch.FloatTensor]] = None
    attentions: Optional[Tuple[torch.FloatTensor]] = None
    reshaped_hidden_states: Optional[Tuple[torch.FloatTensor]] = None


class ScOTConfig(PretrainedConfig):
    """Field-for-field mirror of the reference's ScOTConfig (scOT/model.py:66-132)."""

    model_type = "swinv2"

    attribute_map = {
        "num_attention_heads": "num_heads",
        "num_hidden_layers": "num_layers",
    }

    def __init__(
        self,
        image_size=224,
        patch_size=4,
        num_channels=3,
        num_out_channels=1,
        embed_dim=96,
        depths=[2, 2, 6, 2],
        num_heads=[3, 6, 12, 24],
        skip_connections=[True, True, True],
        window_size=7,
        mlp_ratio=4.0,
        qkv_bias=True,
        hidden_dropout_prob=0.0,
        attention_probs_dropout_prob=0.0,
        drop_path_rate=0.1,
        hidden_act="gelu",
        use_absolute_embeddings=False,
        initializer_range=0.02,
        layer_norm_eps=1e-5,
        p=1,
        channel_slice_list_normalized_loss=None,
        residual_model="convnext",
        use_conditioning=False,
        learn_residual=False,
        **kwargs,
    ):
        super().__init__(**kwargs)
        self.image_size = image_size
        self.patch_size = patch_size
        self.num_channels = num_channels
        self.embed_dim = embed_dim
        self.depths = depths
        self.num_layers = len(depths)
        self.num_heads = num_heads
        self.skip_connections = skip_connections
        self.window_size = window_size
        self.mlp_ratio = mlp_ratio
        self.qkv_bias = qkv_bias
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.drop_path_rate = drop_path_rate
        self.hidden_act = hidden_act
        self.use_absolute_embeddings = use_absolute_embeddings
        self.use_conditioning = use_conditioning
        self.learn_residual = learn_residual if self.use_conditioning else False
        self.layer_norm_eps = layer_norm_eps
        self.initializer_range = initializer_range
        self.hidden_size = int(embed_dim * 2 ** (len(depths) - 1))
        self.pretrained_window_sizes = (0, 0, 0, 0)
        self.num_out_channels = num_out_channels
        self.p = p
        self.channel_slice_list_normalized_loss = channel_slice_list_normalized_loss
        self.residual_model = residual_model


# ------------------------------------------------------------------------------------------------------
# parameter-holding modules (names and shapes are the checkpoint contract; arithmetic lives in the engine)
# ------------------------------------------------------------------------------------------------------
class LayerNorm(nn.LayerNorm):
    """reference scOT/model.py:135-140 (kept importable: scOT/trainer.py:230 uses it for param grouping)."""

    def forward(self, x, time=None):
        return super().forward(x)


class _ClnFunction(torch.autograd.Function):
    """Stand-alone ConditionalLayerNorm through the same CUDA kernels the engine uses (scot_cln_fwd / scot_cln_bwd)."""

    @staticmethod
    def forward(ctx, x, time, aw, ab, cw, cb, eps):
        C = x.shape[-1]
        xs = x.detach().to(torch.float32).contiguous()
        rows = xs.numel() // C
        t = time.detach().reshape(-1).to(device=x.device, dtype=torch.float32).contiguous()
        if rows % t.numel():
            raise ValueError("ConditionalLayerNorm: time must have one entry per sample")
        rps = rows // t.numel()
        f = lambda v: v.detach().reshape(-1).to(torch.float32).contiguous()  # noqa: E731
        aw_, ab_, cw_, cb_ = f(aw), f(ab), f(cw), f(cb)
        y = torch.empty_like(xs)
        zhat = torch.empty(xs.shape, device=x.device, dtype=torch.bfloat16)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        _lib.cln_fwd(xs, None, t, aw_, ab_, cw_, cb_, y, None, zhat, rstd, rows, C, rps, 0, eps)
        ctx.save_for_backward(zhat, rstd, t, aw_, ab_)
        ctx.meta = (rows, C, rps, x.shape, aw.shape, ab.shape)
        return y.view(x.shape).to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        zhat, rstd, t, aw_, ab_ = ctx.saved_tensors
        rows, C, rps, xshape, wshape, bshape = ctx.meta
        dys = dy.detach().to(torch.float32).contiguous()
        dz = torch.empty(rows, C, device=dy.device, dtype=torch.float32)
        g = [torch.zeros(C, device=dy.device, dtype=torch.float32) for _ in range(4)]
        _lib.cln_bwd(dys, zhat, rstd, t, aw_, ab_, dz, True, g[0], g[1], g[2], g[3], None, rows, C, rps, 0)
        return dz.view(xshape).to(dy.dtype), None, g[0].view(wshape), g[1].view(bshape), g[2].view(wshape), g[3].view(bshape), None


class ConditionalLayerNorm(nn.Module):
    """reference scOT/model.py:143-160: same parameters (two Linear(1, dim)), same call signature. The arithmetic is the
    fused CUDA kernel of csrc/norm.cu — inside ScOT the engine calls it directly, stand-alone use goes through
    `_ClnFunction`. There is no torch / CPU evaluation path."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Linear(1, dim)
        self.bias = nn.Linear(1, dim)

    def forward(self, x, time):
        if not x.is_cuda:
            raise RuntimeError("poseidon_b200 ConditionalLayerNorm runs on a CUDA (sm_100a) device only")
        return _ClnFunction.apply(x, time, self.weight.weight, self.weight.bias, self.bias.weight, self.bias.bias, self.eps)


def _norm(config, dim, eps=None):
    cls = ConditionalLayerNorm if config.use_conditioning else LayerNorm
    return cls(dim) if eps is None else cls(dim, eps=eps)


class _Holder(nn.Module):
    """Container whose children only carry parameters."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("sub-modules of the B200 ScOT carry parameters only; call ScOT.forward")


class _SelfAttention(_Holder):
    def __init__(self, config, dim, heads):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.log(10 * torch.ones((heads, 1, 1))))
        self.continuous_position_bias_mlp = nn.Sequential(
            nn.Linear(2, 512, bias=True), nn.ReLU(inplace=True), nn.Linear(512, heads, bias=False)
        )
        self.query = nn.Linear(dim, dim, bias=config.qkv_bias)
        self.key = nn.Linear(dim, dim, bias=False)
        self.value = nn.Linear(dim, dim, bias=config.qkv_bias)


class _Dense(_Holder):
    def __init__(self, i, o):
        super().__init__()
        self.dense = nn.Linear(i, o)


class _Attention(_Holder):
    def __init__(self, config, dim, heads):
        super().__init__()
        self.self = _SelfAttention(config, dim, heads)
        self.output = _Dense(dim, dim)


class _Layer(_Holder):
    def __init__(self, config, dim, heads):
        super().__init__()
        self.attention = _Attention(config, dim, heads)
        self.layernorm_before = _norm(config, dim, config.layer_norm_eps)
        self.intermediate = _Dense(dim, int(config.mlp_ratio * dim))
        self.output = _Dense(int(config.mlp_ratio * dim), dim)
        self.layernorm_after = _norm(config, dim, config.layer_norm_eps)


class _Merging(_Holder):
    def __init__(self, config, dim):
        super().__init__()
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = _norm(config, 2 * dim)


class _Unmerging(_Holder):
    def __init__(self, config, dim):
        super().__init__()
        self.upsample = nn.Linear(dim, 2 * dim, bias=False)
        self.mixup = nn.Linear(dim // 2, dim // 2, bias=False)
        self.norm = _norm(config, dim // 2)


class _EncodeStage(_Holder):
    def __init__(self, config, dim, depth, heads, downsample):
        super().__init__()
        self.blocks = nn.ModuleList([_Layer(config, dim, heads) for _ in range(depth)])
        self.downsample = _Merging(config, dim) if downsample else None


class _DecodeStage(_Holder):
    def __init__(self, config, dim, depth, heads, upsample):
        super().__init__()
        self.blocks = nn.ModuleList([_Layer(config, dim, heads) for _ in range(depth)])
        self.upsample = _Unmerging(config, dim) if upsample else None


class _Stack(_Holder):
    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)


class _PatchEmbeddings(_Holder):
    def __init__(self, config):
        super().__init__()
        self.projection = nn.Conv2d(config.num_channels, config.embed_dim, kernel_size=config.patch_size,
                                    stride=config.patch_size)


class _Embeddings(_Holder):
    def __init__(self, config):
        super().__init__()
        self.patch_embeddings = _PatchEmbeddings(config)
        self.norm = _norm(config, config.embed_dim)


class _PatchRecovery(_Holder):
    def __init__(self, config):
        super().__init__()
        self.projection = nn.ConvTranspose2d(config.embed_dim, config.num_out_channels, kernel_size=config.patch_size,
                                             stride=config.patch_size)
        self.mixup = nn.Conv2d(config.num_out_channels, config.num_out_channels, kernel_size=5, stride=1, padding=2,
                               bias=False)


class ConvNeXtBlock(_Holder):
    """parameters of reference scOT/model.py:163-196"""

    def __init__(self, config, dim, layer_scale_init_value=1e-6):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = _norm(config, dim, config.layer_norm_eps)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.weight = nn.Parameter(layer_scale_init_value * torch.ones((dim)), requires_grad=True)


# ------------------------------------------------------------------------------------------------------
# autograd bridge
# ------------------------------------------------------------------------------------------------------
class _GraphSlot:
    """Static device buffers + captured CUDA graphs for one call signature of `ScOT.forward`
    (batch, labels?, mask mode, grad?). The first call of a signature runs eagerly (one-time kernel
    attribute setup), the second captures `engine.forward` (and `engine.backward`) into CUDA graphs,
    later calls copy the inputs into the static buffers and replay: ~1200 kernel launches become one
    `cudaGraphLaunch`, so the step time no longer depends on the speed of the host."""

    def __init__(self, model, st, batch, has_labels, mask_mode, mask_shape):
        cfg = model.config
        dev = st["device"]
        s = cfg.image_size
        self.calls = 0
        self.x = torch.zeros(batch, cfg.num_channels, s, s, device=dev)
        self.t = torch.zeros(batch, device=dev) if cfg.use_conditioning else None
        self.y = torch.zeros(batch, cfg.num_out_channels, s, s, device=dev) if has_labels else None
        self.mask = torch.zeros(mask_shape, dtype=torch.uint8, device=dev) if mask_mode else None
        self.mask_mode = mask_mode
        self.pred = torch.empty(batch, cfg.num_out_channels, s, s, device=dev)
        self.loss = torch.zeros(1, device=dev) if has_labels else None
        self.gl = torch.ones(1, device=dev)
        self.g_fwd = None
        self.g_bwd = None

    def capture(self, model, st, with_backward):
        eng = st["engine"]
        torch.cuda.synchronize(st["device"])
        self.g_fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fwd):
            eng.forward(st["flat"], st["arena"], self.x, self.t, self.y, self.mask, self.mask_mode, self.pred, self.loss,
                        model.gemm_impl)
        if with_backward and self.loss is not None:
            self.g_bwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_bwd, pool=self.g_fwd.pool()):
                eng.backward(st["flat"], st["gflat"], st["arena"], self.gl, None, model.gemm_impl)


class _ScOTFunction(torch.autograd.Function):
    """One node for the whole model: forward/backward are single calls into the native engine (or
    replays of their CUDA graphs, see `_GraphSlot`)."""

    @staticmethod
    def forward(ctx, model, pixel_values, time, labels, mask, mask_mode, *params):
        st = model._state
        eng = st["engine"]
        ctx.set_materialize_grads(False)  # an unused output (the prediction, usually) arrives as None in backward
        ctx.model = model
        ctx.slot = None
        need_grad = len(params) > 0
        slot = None
        if model.use_cuda_graphs and not torch.cuda.is_current_stream_capturing():
            key = (labels is not None, mask_mode, tuple(mask.shape) if mask is not None else None, need_grad)
            slot = st["slots"].get(key)
            if slot is None:
                slot = st["slots"][key] = _GraphSlot(model, st, pixel_values.shape[0], labels is not None, mask_mode,
                                                     tuple(mask.shape) if mask is not None else None)
            slot.calls += 1
            if slot.calls < 2:
                slot = None
        if slot is not None:
            if slot.g_fwd is None:
                slot.capture(model, st, need_grad)
            slot.x.copy_(pixel_values, non_blocking=True)
            if slot.t is not None:
                slot.t.copy_(time, non_blocking=True)
            if slot.y is not None:
                slot.y.copy_(labels, non_blocking=True)
            if slot.mask is not None:
                slot.mask.copy_(mask, non_blocking=True)
            slot.g_fwd.replay()
            ctx.slot = slot
            # the static outputs are overwritten by the next replay: hand out copies
            pred = slot.pred.clone()
            if slot.loss is None:
                return pred, None
            return pred, slot.loss.clone().reshape(())
        pred = torch.empty((pixel_values.shape[0], model.config.num_out_channels) + tuple(pixel_values.shape[2:]),
                           device=pixel_values.device, dtype=torch.float32)
        loss = torch.zeros(1, device=pixel_values.device, dtype=torch.float32) if labels is not None else None
        eng.forward(st["flat"], st["arena"], pixel_values, time, labels, mask, mask_mode, pred, loss, model.gemm_impl)
        ctx.keep = (pixel_values, time, labels, mask, pred)  # the engine reads these again in backward
        if loss is None:
            return pred, None
        return pred, loss.reshape(())

    @staticmethod
    def backward(ctx, grad_pred, grad_loss):
        model = ctx.model
        st = model._state
        eng = st["engine"]
        nret = len(ctx.needs_input_grad)
        gl = None
        if grad_loss is not None:
            gl = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        gp = None
        if grad_pred is not None:
            gp = grad_pred.detach().to(torch.float32).contiguous()
        if gl is None and gp is None:
            return (None,) * nret
        gflat = st["gflat"]
        assign = model.grad_mode == "assign"
        if not assign:
            gflat.zero_()  # autograd accumulates the returned tensors into .grad itself
        elif all(p.grad is None for p in st["plist"]):
            # assign mode accumulates in place across micro-batches; `zero_grad(set_to_none=True)` (the torch / HF default)
            # only drops the .grad views, so "every .grad is None" is the signal that a new accumulation window starts
            gflat.zero_()
        slot = ctx.slot
        if slot is not None and slot.g_bwd is not None and gp is None:
            slot.gl.copy_(gl, non_blocking=True)
            slot.g_bwd.replay()
        else:
            if slot is not None:
                # eager backward after a replayed forward: point the engine at the static buffers of that forward
                eng.bind_io(slot.x, slot.t, slot.y, slot.mask, slot.mask_mode, slot.pred)
            eng.backward(st["flat"], gflat, st["arena"], gl, gp, model.gemm_impl)
        if assign:
            # gradients are views of the flat buffer; accumulation across micro-batches happens in place
            plist = st["plist"]
            if plist[0].grad is None or plist[-1].grad is None or plist[len(plist) // 2].grad is None:
                for p, gv in zip(plist, st["gviews"]):
                    if p.grad is None:
                        p.grad = gv
            return (None,) * nret
        return (None,) * 6 + tuple(st["gviews"])


class ScOT(PreTrainedModel):
    """B200-native ScOT (reference scOT/model.py:1243-1509). See the module docstring."""

    config_class = ScOTConfig
    config: ScOTConfig
    base_model_prefix = "swinv2"
    main_input_name = "pixel_values"
    supports_gradient_checkpointing = False
    _no_split_modules = ["_EncodeStage", "_DecodeStage"]

    def __init__(self, config, use_mask_token=False):
        super().__init__(config)
        if use_mask_token or config.use_absolute_embeddings:
            raise NotImplementedError("mask tokens / absolute position embeddings are not part of the B200 hot path")
        if config.residual_model != "convnext":
            raise NotImplementedError("only residual_model='convnext' (the reference's shipped setting, train.py:269)")
        if config.hidden_act != "gelu":
            raise NotImplementedError("only hidden_act='gelu'")
        if config.hidden_dropout_prob or config.attention_probs_dropout_prob or config.drop_path_rate:
            # the reference trains with all three at 0 (train.py:259-262); stochastic paths are not implemented
            raise NotImplementedError("dropout / drop_path > 0 is not supported by the B200 engine")
        if not config.qkv_bias:
            raise NotImplementedError("qkv_bias=False is not supported")
        self.config = config
        ns = len(config.depths)
        self.num_layers_encoder = ns
        self.num_layers_decoder = ns
        self.num_features = int(config.embed_dim * 2 ** (ns - 1))
        dims = [int(config.embed_dim * 2 ** i) for i in range(ns)]
        self.embeddings = _Embeddings(config)
        self.encoder = _Stack([
            _EncodeStage(config, dims[i], config.depths[i], config.num_heads[i], downsample=(i < ns - 1))
            for i in range(ns)
        ])
        self.decoder = _Stack([
            _DecodeStage(config, dims[i], config.depths[i], config.num_heads[i], upsample=(i > 0))
            for i in reversed(range(ns))
        ])
        self.patch_recovery = _PatchRecovery(config)
        self.residual_blocks = nn.ModuleList([
            nn.ModuleList([ConvNeXtBlock(config, dims[i]) for _ in range(int(depth))]) if int(depth) > 0
            else nn.ModuleList([nn.Identity()])
            for i, depth in enumerate(config.skip_connections)
        ])
        # engine state (created lazily on the first forward on a CUDA device)
        self._state = None
        self.grad_mode = "autograd"  # "autograd": grads flow through autograd (DDP/hooks work); "assign": .grad = flat views
        self.gemm_impl = _lib.GEMM_TCGEN05
        # "bf16": bf16 GEMM / attention operands (speed mode, what BASELINE.json's "training bf16" configs ask for).
        # "parity": split-bf16 operands (three tcgen05 passes per GEMM) + fp32 attention -> fp32-class results, the mode
        # that meets the north-star 1e-3 tolerance against the fp32 reference (scOT/train.py:311 trains in fp32).
        self.precision = "bf16"
        # replay CUDA graphs of the engine's forward / backward from the second call of a signature on (_GraphSlot)
        self.use_cuda_graphs = True
        self.post_init()

    # ---- HF plumbing ------------------------------------------------------------------------------
    @torch.no_grad()
    def _init_weights(self, module):
        """Same distributions as HF Swinv2PreTrainedModel._init_weights (modeling_swinv2.py:884-902)."""
        std = self.config.initializer_range
        if isinstance(module, (nn.Linear, nn.Conv2d)):
            nn.init.normal_(module.weight, mean=0.0, std=std)
            if module.bias is not None:
                nn.init.zeros_(module.bias)
        elif isinstance(module, nn.LayerNorm):
            nn.init.zeros_(module.bias)
            nn.init.ones_(module.weight)
        elif isinstance(module, _SelfAttention):
            nn.init.constant_(module.logit_scale, math.log(10))

    def get_input_embeddings(self):
        return self.embeddings.patch_embeddings

    # ---- spectral resize, cold path kept on torch.fft (reference model.py:1293-1316) ----------------
    def _downsample(self, image, target_size):
        image_size = image.shape[-2]
        freqs = torch.fft.fftfreq(image_size, d=1 / image_size)
        sel = torch.logical_and(freqs >= -target_size / 2, freqs <= target_size / 2 - 1)
        image_hat = torch.fft.fft2(image, norm="forward")
        image_hat = image_hat[:, :, sel, :][:, :, :, sel]
        return torch.fft.ifft2(image_hat, norm="forward").real

    def _upsample(self, image, target_size):
        image_size = image.shape[-2]
        image_hat = torch.fft.fftshift(torch.fft.fft2(image, norm="forward"))
        pad = (target_size - image_size) // 2
        real = nn.functional.pad(image_hat.real, (pad, pad, pad, pad), value=0.0)
        imag = nn.functional.pad(image_hat.imag, (pad, pad, pad, pad), value=0.0)
        image_hat = torch.fft.ifftshift(torch.complex(real, imag))
        return torch.fft.ifft2(image_hat, norm="forward").real

    # ---- engine state --------------------------------------------------------------------------------
    def _desc(self) -> "_lib.ScotModelDesc":
        cfg = self.config
        ns = len(cfg.depths)
        if ns > 4:
            raise NotImplementedError("at most 4 stages")
        d = _lib.ScotModelDesc()
        d.image_size, d.patch_size = int(cfg.image_size), int(cfg.patch_size)
        d.num_channels, d.num_out_channels = int(cfg.num_channels), int(cfg.num_out_channels)
        d.embed_dim, d.num_stages = int(cfg.embed_dim), ns
        for i in range(ns):
            d.depths[i] = int(cfg.depths[i])
            d.num_heads[i] = int(cfg.num_heads[i])
            d.skip_blocks[i] = int(cfg.skip_connections[i]) if i < len(cfg.skip_connections) else 0
        d.window_size = int(cfg.window_size)
        d.mlp_ratio = float(cfg.mlp_ratio)
        d.use_conditioning = int(bool(cfg.use_conditioning))
        d.learn_residual = int(bool(cfg.learn_residual))
        d.loss_p = int(cfg.p)
        sl = cfg.channel_slice_list_normalized_loss
        d.n_slices = 0 if sl is None else len(sl)
        if sl is not None:
            if len(sl) > 10:
                raise NotImplementedError("at most 9 loss channel groups")
            for i, v in enumerate(sl):
                d.slices[i] = int(v)
        d.layer_norm_eps = float(cfg.layer_norm_eps)
        if self.precision not in ("bf16", "parity"):
            raise ValueError(f"precision must be 'bf16' or 'parity', got {self.precision!r}")
        d.precision = 1 if self.precision == "parity" else 0
        return d

    def _ensure_state(self, device: torch.device, batch: int):
        """(Re)builds the flat parameter / gradient buffers and the engine for this (device, batch)."""
        st = self._state
        plist = None
        if st is not None and st["device"] == device:
            # cheap staleness check: parameters still alias the flat buffer?
            p0, p1 = st["plist"][0], st["plist"][-1]
            if p0.data_ptr() == st["views"][0].data_ptr() and p1.data_ptr() == st["views"][-1].data_ptr():
                if st["batch"] == batch and st["precision"] == self.precision:
                    return st
                plist = st["plist"]
        named = dict(self.named_parameters())
        eng = _lib.Engine(self._desc(), batch)
        if set(eng.table) != set(named):
            missing = sorted(set(eng.table) ^ set(named))[:8]
            raise RuntimeError(f"parameter table mismatch between the module tree and the engine: {missing}")
        if st is not None and plist is not None:
            flat, gflat, views, gviews = st["flat"], st["gflat"], st["views"], st["gviews"]
        else:
            flat = torch.zeros(eng.param_elems, device=device, dtype=torch.float32)
            gflat = torch.zeros(eng.param_elems, device=device, dtype=torch.float32)
            plist, views, gviews = [], [], []
            with torch.no_grad():
                for name, (off, numel, shape) in eng.table.items():
                    p = named[name]
                    if tuple(p.shape) != tuple(shape):
                        raise RuntimeError(f"{name}: shape {tuple(p.shape)} != engine {shape}")
                    if p.device != device:
                        raise RuntimeError(f"{name} is on {p.device}, expected {device}: call model.to('cuda') first")
                    v = flat[off:off + numel].view(shape)
                    v.copy_(p.data.to(torch.float32))
                    p.data = v
                    plist.append(p)
                    views.append(v)
                    gviews.append(gflat[off:off + numel].view(shape))
        arena = torch.empty(eng.workspace_bytes + 256, device=device, dtype=torch.uint8)
        shift = (-arena.data_ptr()) % 256
        arena = arena[shift:shift + eng.workspace_bytes]
        self._state = dict(device=device, batch=batch, precision=self.precision, engine=eng, flat=flat, gflat=gflat,
                           plist=plist, views=views, gviews=gviews, arena=arena, slots={},
                           anchor=torch.zeros(1, device=device, requires_grad=True))
        return self._state

    def zero_grad(self, set_to_none: bool = True):
        """Also clears the flat gradient buffer the engine accumulates into (grad_mode="assign" binds .grad to its views)."""
        if self._state is not None:
            self._state["gflat"].zero_()
        if set_to_none or self._state is None or self.grad_mode != "assign":
            super().zero_grad(set_to_none=set_to_none)

    @property
    def flat_parameters(self) -> torch.Tensor:
        """The single fp32 buffer all parameters view (available after the first forward)."""
        return self._state["flat"]

    @property
    def flat_gradients(self) -> torch.Tensor:
        """The single fp32 buffer the engine accumulates gradients into (one NCCL all-reduce per step)."""
        return self._state["gflat"]

    # ---- forward ----------------------------------------------------------------------------------
    def forward(
        self,
        pixel_values: Optional[torch.FloatTensor] = None,
        time: Optional[torch.FloatTensor] = None,
        bool_masked_pos: Optional[torch.BoolTensor] = None,
        head_mask: Optional[torch.FloatTensor] = None,
        pixel_mask: Optional[torch.BoolTensor] = None,
        labels: Optional[torch.FloatTensor] = None,
        output_attentions: Optional[bool] = None,
        output_hidden_states: Optional[bool] = None,
        return_dict: Optional[bool] = None,
    ) -> Union[Tuple, ScOTOutput]:
        cfg = self.config
        return_dict = return_dict if return_dict is not None else bool(getattr(cfg, "return_dict", True))
        if pixel_values is None:
            raise ValueError("pixel_values cannot be None")
        if bool_masked_pos is not None or head_mask is not None:
            raise NotImplementedError("bool_masked_pos / head_mask are not supported by the B200 engine")
        if output_attentions or output_hidden_states or cfg.output_attentions or cfg.output_hidden_states:
            raise NotImplementedError("attention maps / hidden states are never materialised by the fused engine")
        if not pixel_values.is_cuda:
            raise RuntimeError("poseidon_b200.ScOT runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        if pixel_values.shape[1] != cfg.num_channels:
            raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the configuration.")
        if cfg.use_conditioning and time is None:
            raise ValueError("time is required when use_conditioning=True")
        if pixel_mask is not None and labels is None:
            raise ValueError("pixel_mask needs labels (prediction[pixel_mask] = labels[pixel_mask])")

        image_size = pixel_values.shape[2]
        x = pixel_values.to(torch.float32)
        if image_size != cfg.image_size:  # spectral resize (cold path, torch.fft)
            x = self._upsample(x, cfg.image_size) if image_size < cfg.image_size else self._downsample(x, cfg.image_size)
            if labels is not None or pixel_mask is not None:
                return self._forward_resized(x, image_size, time, pixel_mask, labels, return_dict)
        x = x.contiguous()
        batch = x.shape[0]
        st = self._ensure_state(x.device, batch)
        t = time.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous() if cfg.use_conditioning else None
        if t is not None and t.numel() != batch:
            raise ValueError("time must have one entry per sample")
        y = labels.to(device=x.device, dtype=torch.float32).contiguous() if labels is not None else None
        mask, mask_mode = None, 0
        if pixel_mask is not None:
            pm = pixel_mask.to(device=x.device)
            if pm.dim() == 2 and tuple(pm.shape) == (batch, cfg.num_out_channels):
                mask, mask_mode = pm.to(torch.uint8).contiguous(), 1
            else:
                mask = pm.expand(y.shape).to(torch.uint8).contiguous()
                mask_mode = 2
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in st["plist"])
        if not need_grad:
            params = ()
        elif self.grad_mode == "assign":
            # the engine writes the gradients straight into the flat buffer that .grad views: autograd only has to CALL
            # backward, so one anchor tensor stands in for the ~1600 parameters (their bookkeeping costs ~1 ms per call)
            params = (st["anchor"],)
        else:
            params = st["plist"]
        pred, loss = _ScOTFunction.apply(self, x, t, y, mask, mask_mode, *params)
        if image_size != cfg.image_size:
            pred = self._upsample(pred, image_size) if image_size > cfg.image_size else self._downsample(pred, image_size)
        if not return_dict:
            return ((loss, pred) if loss is not None else (pred,))
        return ScOTOutput(loss=loss, output=pred)

    def _forward_resized(self, x, image_size, time, pixel_mask, labels, return_dict):
        """Resolution != config.image_size with labels: the loss is defined on the re-sampled prediction
        (reference model.py:1416-1484), so it is evaluated with torch ops on top of the engine output."""
        out = self.forward(pixel_values=x, time=time, return_dict=True)
        cfg = self.config
        pred = self._upsample(out.output, image_size) if image_size > cfg.image_size else self._downsample(out.output, image_size)
        if pixel_mask is not None:
            pred = pred.clone()
            pred[pixel_mask] = labels[pixel_mask].type_as(pred)
        fn = nn.functional.l1_loss if cfg.p == 1 else nn.functional.mse_loss
        sl = cfg.channel_slice_list_normalized_loss
        if sl is None:
            loss = fn(pred, labels)
        else:
            loss = torch.mean(torch.stack([
                fn(pred[:, sl[i]:sl[i + 1]], labels[:, sl[i]:sl[i + 1]])
                / (fn(labels[:, sl[i]:sl[i + 1]], torch.zeros_like(labels[:, sl[i]:sl[i + 1]])) + 1e-10)
                for i in range(len(sl) - 1)
            ]))
        if not return_dict:
            return (loss, pred)
        return ScOTOutput(loss=loss, output=pred)
