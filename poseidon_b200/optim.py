"""Fused optimizer step on the flat parameter / gradient buffers of `poseidon_b200.scOT.model.ScOT`
(SURVEY.md §8(f) rank 1): gradient-norm clipping + AdamW in two HBM-bound kernel launches of libscot_b200.so.

Replaces, for this model, what the reference does after every backward pass:
  accelerate `clip_grad_norm_(max_grad_norm)` + `torch.optim.AdamW.step()` ("adamw_torch", scOT/train.py:286) over the
  2-4 parameter groups built in `Trainer.create_optimizer` (scOT/trainer.py:281-400).
`build_param_groups` reproduces that grouping (same names in the same groups — checked against a fixture recorded from
the unmodified reference, tests/golden/param_groups.json); `FlatAdamW` is a `torch.optim.Optimizer`, so learning-rate
schedulers, `state_dict()` and HF `Trainer(optimizers=(opt, sched))` keep working. There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import torch
from torch import nn

from . import _lib


def build_param_groups(model: nn.Module, weight_decay: float, learning_rate_embedding_recovery: Optional[float] = None,
                       learning_rate_time_embedding: Optional[float] = None) -> List[Dict]:
    """The reference's parameter groups (scOT/trainer.py:281-400):
    decay = every parameter outside (Conditional)LayerNorm modules whose name does not contain "bias" (:281-285);
    "embeddings"/"patch_recovery" parameters get `learning_rate_embedding_recovery` (:310-313, :352-356) and the
    ConditionalLayerNorm parameters `learning_rate_time_embedding` (:287-293, :318-319), each only if that rate is set."""
    from transformers.trainer_pt_utils import get_parameter_names

    from .scOT.model import ConditionalLayerNorm, LayerNorm

    decay = get_parameter_names(model, [nn.LayerNorm, LayerNorm, ConditionalLayerNorm])
    decay = {n for n in decay if "bias" not in n}
    time_params = set()
    for name, module in model.named_modules():
        if isinstance(module, ConditionalLayerNorm):
            for pn, _ in module.named_parameters():
                time_params.add(f"{name}.{pn}")
    standard, no_decay, embeddings, time_emb = [], [], [], []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if learning_rate_embedding_recovery is not None and ("embeddings" in n or "patch_recovery" in n):
            embeddings.append(p)
        elif n in decay:
            standard.append(p)
        elif learning_rate_time_embedding is not None and n in time_params:
            time_emb.append(p)
        else:
            no_decay.append(p)
    groups = [{"params": standard, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}]
    if learning_rate_embedding_recovery is not None:
        groups.append({"params": embeddings, "lr": learning_rate_embedding_recovery, "weight_decay": weight_decay})
    if learning_rate_time_embedding is not None:
        groups.append({"params": time_emb, "lr": learning_rate_time_embedding, "weight_decay": 0.0})
    return groups


class FlatAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction, amsgrad=False) executed by one fused
    kernel over the model's flat fp32 buffers; optional fused `clip_grad_norm_`.

    `params` are ordinary parameter groups (e.g. from `build_param_groups`); `model` is the ScOT whose parameters they
    are. The flat buffers exist after the model's first forward on its CUDA device (or `model._ensure_state`), so the
    binding happens lazily at the first `step()`.
    """

    def __init__(self, params, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None, grad_scale: float = 1.0, allreduce: Optional[bool] = None):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) > 254:
            raise ValueError("at most 254 parameter groups")
        self.model = model
        self.max_grad_norm = max_grad_norm
        self.grad_scale = grad_scale
        # Data parallelism. grad_mode="autograd" (default): the gradients reach `p.grad` through autograd, DDP / accelerate
        # reduce them there and step() gathers them into the flat buffer. grad_mode="assign": `.grad` ARE views of the flat
        # buffer and nobody has reduced them — with allreduce=True (or None = "auto": a process group with world_size > 1
        # exists) step() sums the flat buffer over the ranks in ONE collective and scales by 1/world_size.
        self.allreduce = allreduce
        self._bound = None
        self._step = 0
        self._pending_state = None  # a state dict loaded before the flat buffers exist (HF Trainer resume)
        self._hp_ring, self._hp_events, self._hp_slot = [], [], 0

    # ---- binding to the flat buffers -----------------------------------------------------------------
    def _bind(self):
        st = self.model._state
        if st is None:
            raise RuntimeError("FlatAdamW: the model has no flat buffers yet — run one forward on its CUDA device first")
        flat, gflat = st["flat"], st["gflat"]
        if self._bound is not None and self._bound["flat"] is flat:
            return self._bound
        n = flat.numel()
        if n % 64:
            raise RuntimeError("flat parameter buffer length is not a multiple of 64")
        base = flat.data_ptr()
        cg = torch.full((n // 64,), 255, dtype=torch.uint8)
        for gid, group in enumerate(self.param_groups):
            for p in group["params"]:
                off = (p.data_ptr() - base) // 4
                if p.device != flat.device or off < 0 or off + p.numel() > n or (p.data_ptr() - base) % 4:
                    raise RuntimeError("FlatAdamW: a parameter does not live in the model's flat buffer")
                c0, c1 = off // 64, (off + p.numel() + 63) // 64
                seg = cg[c0:c1]
                if bool(((seg != 255) & (seg != gid)).any()):
                    raise RuntimeError("FlatAdamW: two parameter groups share a 64-element chunk of the flat buffer")
                seg.fill_(gid)
        dev = flat.device
        old = self._bound
        b = dict(flat=flat, gflat=gflat, chunk_group=cg.to(dev),
                 exp_avg=torch.zeros_like(flat), exp_avg_sq=torch.zeros_like(flat),
                 hp=torch.zeros(len(self.param_groups), 8, device=dev), sq_norm=torch.zeros(1, device=dev))
        # hyper-parameters travel through a RING of pinned host buffers guarded by events: with graph replay the host runs
        # steps ahead of the device, a single buffer would be rewritten before its pending H2D copy has read it
        self._hp_ring = [torch.zeros(len(self.param_groups), 8).pin_memory() for _ in range(4)]
        self._hp_events = [None] * 4
        if old is not None:  # the model re-allocated its buffers (e.g. moved): carry the moments over
            b["exp_avg"].copy_(old["exp_avg"])
            b["exp_avg_sq"].copy_(old["exp_avg_sq"])
        # expose the moments per parameter (views) so that state_dict() / HF checkpointing see ordinary Adam state
        for group in self.param_groups:
            for p in group["params"]:
                off = (p.data_ptr() - base) // 4
                self.state[p] = {"step": torch.tensor(float(self._step)),
                                 "exp_avg": b["exp_avg"][off:off + p.numel()].view(p.shape),
                                 "exp_avg_sq": b["exp_avg_sq"][off:off + p.numel()].view(p.shape)}
        self._bound = b
        if self._pending_state is not None:
            pending, self._pending_state = self._pending_state, None
            self.load_state_dict(pending)
        return b

    def _gather_grads(self, b):
        """Makes the flat gradient buffer hold what the optimizer must apply (see `allreduce` in __init__)."""
        st = self.model._state
        if getattr(self.model, "grad_mode", "assign") != "assign":
            # autograd mode: the (possibly DDP-reduced, possibly accumulated) gradients live in p.grad
            dst, src, missing = [], [], []
            for p, gv in zip(st["plist"], st["gviews"]):
                if p.grad is None:
                    missing.append(gv)
                elif p.grad.data_ptr() != gv.data_ptr():
                    dst.append(gv)
                    src.append(p.grad)
            if dst:
                torch._foreach_copy_(dst, src)
            if missing:
                torch._foreach_zero_(missing)
            return self.grad_scale
        scale = self.grad_scale
        dist = torch.distributed
        want = self.allreduce
        if want is None:
            want = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if want:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("FlatAdamW(allreduce=True) needs an initialised torch.distributed process group")
            dist.all_reduce(b["gflat"])
            scale = scale / dist.get_world_size()
        return scale

    # ---- one optimizer step --------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        b = self._bind()
        grad_scale = self._gather_grads(b)
        self._step += 1
        t = self._step
        k = self._hp_slot
        self._hp_slot = (k + 1) % len(self._hp_ring)
        if self._hp_events[k] is not None:
            self._hp_events[k].synchronize()  # the copy issued four steps ago has read this buffer
        hp = self._hp_ring[k]
        for gid, g in enumerate(self.param_groups):
            b1, b2 = g["betas"]
            hp[gid, 0], hp[gid, 1], hp[gid, 2], hp[gid, 3], hp[gid, 4] = g["lr"], g["weight_decay"], b1, b2, g["eps"]
            hp[gid, 5] = 1.0 - b1 ** t
            hp[gid, 6] = math.sqrt(1.0 - b2 ** t)
        b["hp"].copy_(hp, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._hp_events[k] = ev
        lib = _lib.load()
        stream = _lib.cur_stream()
        n = b["flat"].numel()
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        if clip:
            _lib.check(lib.scot_grad_sq_norm(_lib.ptr(b["gflat"]), n, _lib.ptr(b["sq_norm"]), stream), "scot_grad_sq_norm")
        _lib.check(lib.scot_adamw_step(_lib.ptr(b["flat"]), _lib.ptr(b["gflat"]), _lib.ptr(b["exp_avg"]),
                                       _lib.ptr(b["exp_avg_sq"]), None, _lib.ptr(b["chunk_group"]), n, _lib.ptr(b["hp"]),
                                       len(self.param_groups), _lib.ptr(b["sq_norm"]) if clip else None,
                                       float(self.max_grad_norm or 0.0), float(grad_scale), stream), "scot_adamw_step")
        self._last_scale = float(grad_scale)
        for s in self.state.values():
            s["step"].fill_(float(t))
        return loss

    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm of the last clipped step (device scalar; what clip_grad_norm_ returns)."""
        return self._bound["sq_norm"].sqrt() * getattr(self, "_last_scale", self.grad_scale)

    def zero_grad(self, set_to_none: bool = False):
        """Gradients are views of the flat buffer: clear it in one memset and keep the views bound."""
        st = self.model._state
        if st is not None:
            st["gflat"].zero_()
        else:
            super().zero_grad(set_to_none=set_to_none)

    def load_state_dict(self, state_dict):
        """Loads a torch.optim.AdamW / FlatAdamW state dict: the moments are copied INTO the flat buffers. Before the
        model's first forward (HF Trainer restores the optimizer before the first step) the dict is kept and applied when
        the flat buffers appear."""
        if self.model._state is None:
            self._pending_state = state_dict
            return
        b = self._bind()
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        saved_ids = [i for g in groups for i in g["params"]]
        params = [p for g in self.param_groups for p in g["params"]]
        if len(saved_ids) != len(params):
            raise ValueError("loaded state dict contains a different number of parameters")
        step = 0
        for sid, p in zip(saved_ids, params):
            s = state_dict["state"].get(sid)
            if s is None:
                continue
            self.state[p]["exp_avg"].copy_(s["exp_avg"])
            self.state[p]["exp_avg_sq"].copy_(s["exp_avg_sq"])
            step = max(step, int(float(s["step"])))
        for g_new, g_old in zip(self.param_groups, groups):
            for k, v in g_old.items():
                if k != "params":
                    g_new[k] = v
        self._step = step
        for s in self.state.values():
            s["step"].fill_(float(step))
        del b
