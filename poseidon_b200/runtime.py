"""Training-step runtime on top of the native engine: CUDA-graph captured forward+backward and the
data-parallel gradient exchange (one NCCL all-reduce over the flat gradient buffer per step).

The reference gets its data parallelism from accelerate -> torch DDP (~25 bucketed all-reduces per step,
SURVEY.md §2 row 11); here the gradients already live in one flat buffer, so a step is
    graph replay (zero grads, forward, backward)  ->  all_reduce(flat_grads)  [-> optimizer].
With more than one rank the backward graph is cut where ~85 % of the gradient bytes are final (decoder, ConvNeXt skips,
deepest encoder stage: one contiguous range of the flat buffer): that range is all-reduced on a communication stream
while the rest of the backward pass runs, the small remainder afterwards — two NCCL calls, one of them hidden.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .scOT.model import ScOT


class GraphedTrainStep:
    """Captures `zero_grad -> engine.forward -> engine.backward` for a fixed batch shape into a CUDA graph.

    Inputs are copied into static device buffers (`load_batch`), `run()` replays the graph; afterwards
    `model.flat_gradients` holds d(loss)/d(params) (already divided by `world_size` if `average=True`),
    `self.loss` the scalar loss and `self.pred` the prediction.
    """

    def __init__(self, model: ScOT, batch: int, device: Optional[torch.device] = None, use_mask: bool = False,
                 use_graph: bool = True, world_size: int = 1, average: bool = True, optimizer=None,
                 overlap_allreduce: Optional[bool] = None):
        cfg = model.config
        self.model = model
        self.device = device or next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs a CUDA device (no CPU fallback)")
        self.batch = batch
        self.world_size = world_size
        st = model._ensure_state(self.device, batch)
        self.st = st
        s = cfg.image_size
        self.x = torch.zeros(batch, cfg.num_channels, s, s, device=self.device)
        self.y = torch.zeros(batch, cfg.num_out_channels, s, s, device=self.device)
        self.t = torch.zeros(batch, device=self.device) if cfg.use_conditioning else None
        self.mask = torch.zeros(batch, cfg.num_out_channels, dtype=torch.uint8, device=self.device) if use_mask else None
        self.pred = torch.empty(batch, cfg.num_out_channels, s, s, device=self.device)
        self.loss = torch.zeros(1, device=self.device)
        # d(loss)/d(loss): pre-scaled by 1/world so that the summed all-reduce yields the DDP mean
        self.gscale = torch.full((1,), (1.0 / world_size) if average else 1.0, device=self.device)
        self.graph = None
        self.graph2 = None
        # overlap of the gradient exchange with the tail of the backward pass (default: whenever there is an exchange)
        self.overlap = (world_size > 1) if overlap_allreduce is None else bool(overlap_allreduce)
        self.split = st["engine"].grad_split if self.overlap else 0
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.overlap else None
        self.mid_event = torch.cuda.Event() if self.overlap else None
        self.optimizer = optimizer  # e.g. poseidon_b200.optim.FlatAdamW: fused clip + AdamW on the flat buffers
        self._impl = model.gemm_impl
        # bind .grad to the flat views once; the graph zeroes and refills the same memory every step
        for p, gv in zip(st["plist"], st["gviews"]):
            p.grad = gv
        if use_graph:
            self._capture()

    def _body(self, part: int = 0):
        """part 0: everything; 1: zero grads + forward + first part of the backward; 2: rest of the backward"""
        st = self.st
        if part != 2:
            st["gflat"].zero_()
            st["engine"].forward(st["flat"], st["arena"], self.x, self.t, self.y, self.mask, 1 if self.mask is not None else 0,
                                 self.pred, self.loss, self._impl)
        st["engine"].backward(st["flat"], st["gflat"], st["arena"], self.gscale, None, self._impl, part=part)

    def _capture(self):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up outside capture: one-time attribute setup, lazy module loading
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        if not self.overlap:
            with torch.cuda.graph(self.graph):
                self._body()
            return
        with torch.cuda.graph(self.graph):
            self._body(1)
        self.graph2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph2, pool=self.graph.pool()):
            self._body(2)

    def load_batch(self, pixel_values, time, labels, pixel_mask=None, non_blocking=True):
        self.x.copy_(pixel_values, non_blocking=non_blocking)
        self.y.copy_(labels, non_blocking=non_blocking)
        if self.t is not None:
            self.t.copy_(time, non_blocking=non_blocking)
        if self.mask is not None and pixel_mask is not None:
            self.mask.copy_(pixel_mask.to(torch.uint8), non_blocking=non_blocking)

    def run(self):
        if not self.overlap:
            if self.graph is not None:
                self.graph.replay()
            else:
                self._body()
            return
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body(1)
        self.mid_event.record()  # gradients [split, end) are final from here on
        if self.world_size > 1:
            self.comm_stream.wait_event(self.mid_event)
            with torch.cuda.stream(self.comm_stream):
                torch.distributed.all_reduce(self.st["gflat"][self.split:])
        if self.graph2 is not None:
            self.graph2.replay()
        else:
            self._body(2)

    def allreduce(self):
        """The gradient exchange of a data-parallel step (NCCL over NVLink / NVSwitch): one all-reduce of the flat buffer,
        or — overlapped mode — the remainder [0, split) plus the join with the communication stream, on which the bulk
        [split, end) has been in flight since the middle of the backward pass."""
        if self.world_size <= 1:
            return
        if not self.overlap:
            torch.distributed.all_reduce(self.st["gflat"])
            return
        if self.split > 0:
            torch.distributed.all_reduce(self.st["gflat"][:self.split])
        torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)

    def optimizer_step(self):
        """clip_grad_norm_ + AdamW on the flat buffers (two launches), after the gradient all-reduce."""
        if self.optimizer is not None:
            self.optimizer.step()

    def train_step(self):
        """graph replay (zero grads, forward, backward) -> one all-reduce -> fused optimizer step"""
        self.run()
        self.allreduce()
        self.optimizer_step()

    def launches_per_step(self) -> int:
        lib = _lib.load()
        before = lib.scot_launch_count()
        self._body()  # part 0 == parts 1 + 2
        torch.cuda.synchronize(self.device)
        return int(lib.scot_launch_count() - before)


class ARRollout:
    """Autoregressive rollout on the device (reference scOT/trainer.py:452-603 `_model_forward`
    with `ar_steps`, scOT/inference.py:210-235): the prediction is fed back as the next input — extra input channels
    (num_channels > num_out_channels) are carried over (trainer.py:488-501) — with the lead time divided by the
    number of steps (int `ar_steps`, :459) or multiplied per step (list `ar_steps`, :528-534).

    One CUDA graph holds `engine.forward` + the feedback copy; a rollout is `n` replays on one stream: no host
    synchronisation, no allocation and no Python-side tensor work between the steps.
    """

    def __init__(self, model: ScOT, batch: int, device: Optional[torch.device] = None, with_labels: bool = False,
                 mask_shape=None, use_graph: bool = True):
        cfg = model.config
        if not cfg.use_conditioning:
            raise ValueError("autoregressive rollouts need a time-conditioned model (reference trainer.py:453)")
        self.model = model
        self.device = device or next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("ARRollout needs a CUDA device (no CPU fallback)")
        self.batch = batch
        self.st = model._ensure_state(self.device, batch)
        s = cfg.image_size
        self.cin, self.cout = cfg.num_channels, cfg.num_out_channels
        if self.cout > self.cin:
            raise ValueError("rollout needs num_out_channels <= num_channels")
        self.x = torch.zeros(batch, self.cin, s, s, device=self.device)
        self.t = torch.zeros(batch, device=self.device)
        self.y = torch.zeros(batch, self.cout, s, s, device=self.device) if with_labels else None
        self.mask, self.mask_mode = None, 0
        if mask_shape is not None:
            if not with_labels:
                raise ValueError("pixel_mask needs labels")
            self.mask = torch.zeros(tuple(mask_shape), dtype=torch.uint8, device=self.device)
            self.mask_mode = 1 if len(mask_shape) == 2 else 2
        self.pred = torch.empty(batch, self.cout, s, s, device=self.device)
        self.loss = torch.zeros(1, device=self.device) if with_labels else None
        self.graph = None
        self._impl = model.gemm_impl
        if use_graph:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._body()  # warm-up outside capture
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body()

    def _body(self):
        st = self.st
        st["engine"].forward(st["flat"], st["arena"], self.x, self.t, self.y, self.mask, self.mask_mode, self.pred, self.loss,
                             self._impl)
        # feedback: the prediction replaces the first num_out_channels input channels, the rest is carried over
        self.x[:, :self.cout].copy_(self.pred)

    @torch.no_grad()
    def run(self, pixel_values, time, ar_steps, labels=None, pixel_mask=None, output_all_steps: bool = False):
        """Returns (output, loss): output [B, Cout, H, W] of the last step, or [B, steps, Cout, H, W] with
        `output_all_steps`; loss = mean over the steps (or the per-step stack), None without labels."""
        if isinstance(ar_steps, int):
            factors = [1.0 / ar_steps] * ar_steps
        else:
            factors = [float(i) for i in ar_steps]
        self.x.copy_(pixel_values, non_blocking=True)
        lead = time.to(device=self.device, dtype=torch.float32).reshape(-1)
        if (labels is not None) != (self.y is not None):
            raise ValueError("ARRollout was built with_labels=%s" % (self.y is not None))
        if labels is not None:
            self.y.copy_(labels, non_blocking=True)
        if self.mask is not None:
            if pixel_mask is None:
                raise ValueError("ARRollout was built with a pixel mask")
            self.mask.copy_(pixel_mask.to(torch.uint8), non_blocking=True)
        outs, losses = [], []
        for f in factors:
            torch.mul(lead, f, out=self.t)
            if self.graph is not None:
                self.graph.replay()
            else:
                self._body()
            if output_all_steps:
                outs.append(self.pred.clone())
            if self.loss is not None:
                losses.append(self.loss.clone())
        if output_all_steps:
            output = torch.stack(outs, dim=1)
            loss = torch.stack(losses, dim=0).reshape(-1) if losses else None
        else:
            output = self.pred.clone()
            loss = (torch.stack(losses).sum() / len(factors)) if losses else None
        return output, loss


class DevicePrefetcher:
    """Double-buffered host -> device input pipeline (the part that touches the hot path):
    the pinned host batch of step i+1 is copied on a side stream while step i computes, so the H2D transfer
    (42 MB per Poseidon-B step of 64 samples) leaves the critical path. `loader` yields dicts of (pinned) host tensors —
    e.g. the reference's default collator output `pixel_values / labels / time / pixel_mask`; iteration yields dicts of
    device tensors that stay valid until the next-but-one `next()`.
    """

    def __init__(self, loader, device, depth: int = 2):
        self.it = iter(loader)
        self.device = torch.device(device)
        self.depth = depth
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None] * depth
        self.ready = [None] * depth
        self.freed = [None] * depth
        self.i = 0  # next batch to hand out
        self.j = 0  # next batch to fetch

    def _fill(self) -> bool:
        try:
            batch = next(self.it)
        except StopIteration:
            return False
        k = self.j % self.depth
        with torch.cuda.stream(self.stream):
            if self.freed[k] is not None:
                self.stream.wait_event(self.freed[k])  # the step that consumed this slot has finished on the GPU
            slot = self.slots[k]
            if slot is None or set(slot) != set(batch) or any(slot[n].shape != batch[n].shape for n in batch):
                slot = self.slots[k] = {n: torch.empty(v.shape, dtype=v.dtype, device=self.device) for n, v in batch.items()}
            for n, v in batch.items():
                slot[n].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self.ready[k] = ev
        self.j += 1
        return True

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self.i > 0:
            ev = torch.cuda.Event()
            ev.record(cur)  # everything enqueued for the previous batch precedes this point
            self.freed[(self.i - 1) % self.depth] = ev
        while self.j < self.i + self.depth and self._fill():
            pass
        if self.i >= self.j:
            raise StopIteration
        k = self.i % self.depth
        cur.wait_event(self.ready[k])
        self.i += 1
        return self.slots[k]
