"""Training-step runtime on top of the native engine: CUDA-graph captured forward+backward and the
data-parallel gradient exchange (one NCCL all-reduce over the flat gradient buffer per step).

The reference gets its data parallelism from accelerate -> torch DDP (~25 bucketed all-reduces per step,
SURVEY.md §2 row 11); here the gradients already live in one flat buffer, so a step is
    graph replay (zero grads, forward, backward)  ->  all_reduce(flat_grads)  [-> optimizer].
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
                 use_graph: bool = True, world_size: int = 1, average: bool = True):
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
        self._impl = model.gemm_impl
        # bind .grad to the flat views once; the graph zeroes and refills the same memory every step
        for p, gv in zip(st["plist"], st["gviews"]):
            p.grad = gv
        if use_graph:
            self._capture()

    def _body(self):
        st = self.st
        st["gflat"].zero_()
        st["engine"].forward(st["flat"], st["arena"], self.x, self.t, self.y, self.mask, 1 if self.mask is not None else 0,
                             self.pred, self.loss, self._impl)
        st["engine"].backward(st["flat"], st["gflat"], st["arena"], self.gscale, None, self._impl)

    def _capture(self):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up outside capture: one-time attribute setup, lazy module loading
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()

    def load_batch(self, pixel_values, time, labels, pixel_mask=None, non_blocking=True):
        self.x.copy_(pixel_values, non_blocking=non_blocking)
        self.y.copy_(labels, non_blocking=non_blocking)
        if self.t is not None:
            self.t.copy_(time, non_blocking=non_blocking)
        if self.mask is not None and pixel_mask is not None:
            self.mask.copy_(pixel_mask.to(torch.uint8), non_blocking=non_blocking)

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()

    def allreduce(self):
        """The single gradient collective of a data-parallel step (NCCL over NVLink/NVSwitch)."""
        if self.world_size > 1:
            torch.distributed.all_reduce(self.st["gflat"])

    def launches_per_step(self) -> int:
        lib = _lib.load()
        before = lib.scot_launch_count()
        self._body()
        torch.cuda.synchronize(self.device)
        return int(lib.scot_launch_count() - before)
