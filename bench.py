#!/usr/bin/env python
"""Benchmark of the scOT hot path (BASELINE.json: samples/sec fwd+bwd, Poseidon-B 128x128; % tensor roofline).

  python bench.py --gpus N --steps K --warmup W            # our engine (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port) on host cores

A step = zero grads + forward + backward of Poseidon-B on one synthetic batch (64 x 5 x 128 x 128 per GPU,
bf16 GEMM operands / fp32 accumulation & residual stream) [+ one NCCL all-reduce of the flat gradient
buffer when N > 1]. Prints ONE JSON line (rank 0). See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODELS = {  # reference scOT/train.py:35-72
    "T": dict(embed_dim=48, depths=[4, 4, 4, 4], num_heads=[3, 6, 12, 24]),
    "B": dict(embed_dim=96, depths=[8, 8, 8, 8], num_heads=[3, 6, 12, 24]),
    "L": dict(embed_dim=192, depths=[8, 8, 8, 8], num_heads=[3, 6, 12, 24]),
}


def model_config(name: str, channels: int, size: int = 128):
    slices = {4: [0, 1, 3, 4], 5: [0, 1, 3, 4, 5]}.get(channels, [0, channels])
    return dict(image_size=size, patch_size=4, num_channels=channels, num_out_channels=channels,
                skip_connections=[2, 2, 2, 0], window_size=16, mlp_ratio=4.0, drop_path_rate=0.0, use_conditioning=True,
                p=1, channel_slice_list_normalized_loss=slices, residual_model="convnext", **MODELS[name])


def flops_forward_per_sample(cfg: dict):
    """Algorithmic forward FLOPs per sample (multiply-add = 2); excludes the CPB MLP."""
    grid = cfg["image_size"] // cfg["patch_size"]
    ns = len(cfg["depths"])
    total = 0.0
    for s in range(ns):
        res = grid >> s
        T = res * res
        C = cfg["embed_dim"] << s
        h = cfg["num_heads"][s]
        ws = min(cfg["window_size"], res)
        N, nW, hd = ws * ws, (res // ws) ** 2, C // h
        d = cfg["depths"][s]
        total += 2 * (2 * T * C * C * 12 * d)                # q,k,v,proj,mlp1,mlp2 (encoder + decoder)
        total += 2 * (4 * nW * h * N * N * hd * d)           # QK^T and PV
        if s < ns - 1:
            total += 2 * (T // 4) * (4 * C) * (2 * C)        # patch merging
        if s > 0:
            total += 2 * T * C * 2 * C + 2 * 4 * T * (C // 2) ** 2  # patch unmerging
        total += cfg["skip_connections"][s] * (2 * T * C * 49 + 16 * T * C * C)  # ConvNeXt
    C0 = cfg["embed_dim"]
    cin, cout, S = cfg["num_channels"], cfg["num_out_channels"], cfg["image_size"]
    total += 2 * grid * grid * C0 * 16 * (cin + cout) + 50 * S * S * cout * cout
    return total


def flops_cpb_per_step(cfg: dict):
    grid = cfg["image_size"] // cfg["patch_size"]
    tot = 0.0
    for s, d in enumerate(cfg["depths"]):
        res = grid >> s
        ws = min(cfg["window_size"], res)
        R = (2 * ws - 1) ** 2
        tot += 2 * d * (2 * R * 2 * 512 + 2 * R * 512 * cfg["num_heads"][s])
    return tot


def realistic_init_(model, seed=0):
    """Random-init weights of the architecture with O(1) activations (HF's default init makes every
    ConditionalLayerNorm scale ~0.02*t). Values do not affect the timing."""
    from poseidon_b200.scOT.model import ConditionalLayerNorm, ConvNeXtBlock

    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2 and "continuous_position_bias_mlp" not in name and "norm" not in name:
                fan_in = p[0].numel() if "patch_recovery.projection" not in name else p.shape[0]
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(fan_in))
        for mod in model.modules():
            if isinstance(mod, ConditionalLayerNorm):
                mod.weight.bias.fill_(1.0)
                mod.bias.bias.normal_(0, 0.02, generator=g)
            if isinstance(mod, ConvNeXtBlock):
                mod.weight.fill_(0.5)


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_leg(cfg: dict, batch: int, steps: int, warmup: int):
    """The reference's own (CPU, eager, fp32) implementation of the path: the oracle port, all host threads."""
    import types

    from oracle import scot_oracle as O
    from oracle.weights import make_inputs, make_weights

    # torch's intra-op pool stops scaling (and collapses from oversubscription) far below the 100+ cores of
    # the GPU hosts: use at most 32 threads and report the number actually used
    n = min(len(os.sched_getaffinity(0)), 32)
    torch.set_num_threads(n)
    from poseidon_b200 import _lib  # only for the parameter table (host side, no GPU work)

    ocfg = types.SimpleNamespace(**cfg)
    ocfg.layer_norm_eps, ocfg.learn_residual = 1e-5, False
    shapes = param_shapes(cfg)
    w = {k: v.requires_grad_(True) for k, v in make_weights(shapes, seed=0).items()}
    x, t, y, pm = make_inputs(batch, cfg["num_channels"], cfg["num_out_channels"], cfg["image_size"], seed=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for v in w.values():
            v.grad = None
        loss, _ = O.scot_forward(ocfg, w, x, t, y, None)
        loss.backward()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, n, sec


def param_shapes(cfg: dict):
    from poseidon_b200 import _lib
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    with torch.device("meta"):
        m = ScOT(ScOTConfig(**cfg))
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="B", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--channels", type=int, default=5)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=4)
    args = ap.parse_args()
    cfg = model_config(args.model, args.channels)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"Poseidon-{args.model} fwd+bwd, {args.channels}ch 128x128, batch {args.batch}/GPU"

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
        sps, cores, sec = cpu_reference_leg(cfg, args.cpu_batch, steps, warm)
        print(json.dumps({
            "impl": "reference", "metric": "samples/sec (fwd+bwd)", "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "note": "reference CPU path = oracle port (torch eager fp32) on host cores"},
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} fwd+bwd steps of batch {args.cpu_batch} after {warm} warm-up"},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the scOT engine has no CPU fallback (use --impl reference for the CPU leg)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    from poseidon_b200 import _lib
    from poseidon_b200.runtime import GraphedTrainStep
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    torch.manual_seed(0)
    model = ScOT(ScOTConfig(**cfg))
    realistic_init_(model)
    model = model.to(dev)
    B = args.batch
    step = GraphedTrainStep(model, B, dev, use_graph=not args.no_graph, world_size=world)
    gen = torch.Generator().manual_seed(1234 + rank)
    S = cfg["image_size"]
    hx = torch.randn(B, args.channels, S, S, generator=gen).pin_memory()
    hy = torch.randn(B, args.channels, S, S, generator=gen).pin_memory()
    ht = torch.rand(B, generator=gen).pin_memory()
    step.load_batch(hx, ht, hy)
    launches = step.launches_per_step() + 1  # + the gradient memset kernel issued by torch

    def one_step():
        step.run()
        step.allreduce()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (inputs already in HBM). Working set per step (~8 GB of activations,
    # 0.6 GB weights) is far larger than the 126 MB L2, so no explicit L2 flush is needed between steps.
    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    # ---- end to end through the public API: pinned host inputs -> H2D -> model(**batch) -> loss.backward() -> D2H
    model.grad_mode = "assign"
    for p in model.parameters():
        p.grad = None

    from poseidon_b200.runtime import DevicePrefetcher

    def host_batches():
        while True:  # the data loader: every step's batch sits in pinned host memory
            yield {"pixel_values": hx, "labels": hy, "time": ht}

    # double-buffered H2D: the copy of step i+1 is issued (on a side stream) inside step i's timed region
    feed = DevicePrefetcher(host_batches(), dev)

    def e2e_step():
        batch = next(feed)  # issues the H2D copies of the next batch, waits (on the GPU) for this one
        model.flat_gradients.zero_()
        out = model(pixel_values=batch["pixel_values"], time=batch["time"], labels=batch["labels"])
        out.loss.backward()
        if world > 1:
            torch.distributed.all_reduce(model.flat_gradients)
        return float(out.loss.detach())  # D2H read of the step result

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, args.steps // 2)
    for _ in range(n_e2e):
        last_loss = e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) / n_e2e * 1e3
    # ---- informational: the same step followed by the fused optimizer (grad-norm clip + AdamW on the flat buffers)
    from poseidon_b200.optim import FlatAdamW, build_param_groups
    opt = FlatAdamW(build_param_groups(model, 0.01), model, lr=1e-6, max_grad_norm=5.0)
    step.optimizer = opt
    for _ in range(3):
        step.train_step()
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for _ in range(max(3, args.steps // 2)):
        step.train_step()
    o1.record()
    barrier()
    opt_ms = o0.elapsed_time(o1) / max(3, args.steps // 2)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms, e2e_ms, opt_ms], device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms, e2e_ms, opt_ms = float(tt[0]), float(tt[1]), float(tt[2])
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # ---- dominant kernel (roofline): the tcgen05 GEMM family is ~45 % of the step (profiles/); its most expensive
    # instance is the stage-0 MLP up-projection with the fused GELU epilogue: [B*1024, C] x [4C, C]^T -> gelu'(h), gelu(h).
    # HBM bound (K = C = 96: 65 FLOP/B): algorithmic bytes = A + W + bias + two bf16 outputs. Timed alone with CUDA events on
    # the launching stream; three buffer sets (3 x 113 MB > 126 MB L2) are cycled so that no launch finds its data in L2.
    from poseidon_b200 import _lib as L
    C0 = cfg["embed_dim"]
    Mk, Nk, Kk = B * (cfg["image_size"] // cfg["patch_size"]) ** 2, 4 * C0, C0
    sets = []
    for _ in range(3):
        sets.append((torch.randn(Mk, Kk, device=dev).bfloat16(), torch.randn(Nk, Kk, device=dev).bfloat16(),
                     torch.randn(Nk, device=dev), torch.empty(Mk, Nk, device=dev, dtype=torch.bfloat16),
                     torch.empty(Mk, Nk, device=dev, dtype=torch.bfloat16)))

    def kern(i):
        a_, b_, bias_, o0, o1 = sets[i % 3]
        L.gemm(a_, b_, Mk, Nk, Kk, mode=L.EPI_GELU, bias=bias_, out0=o0, out1=o1)

    for i in range(6):
        kern(i)
    torch.cuda.synchronize(dev)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nk = 30
    k0.record()
    for i in range(nk):
        kern(i)
    k1.record()
    torch.cuda.synchronize(dev)
    kern_us = k0.elapsed_time(k1) / nk * 1e3
    kern_bytes = (Mk * Kk + Nk * Kk) * 2 + Nk * 4 + 2 * Mk * Nk * 2
    del sets

    total_samples = B * world
    sps = total_samples / (ms * 1e-3)
    f_step = 3.0 * (B * flops_forward_per_sample(cfg) + flops_cpb_per_step(cfg))  # per GPU
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        peak_tf, peak_src = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1590.0))), "MEASURED_PEAKS.json bf16_tflops_sustained"
        peak_hbm, hbm_src = float(pk.get("hbm_gbs", 6650.0)), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak_tf, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
        peak_hbm, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved_tf = f_step / (ms * 1e-3) / 1e12
    rec = {
        "metric": "samples/sec (fwd+bwd)", "value": sps, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "global_batch": total_samples, "parallelism": f"dp{world}",
                   "cuda_graph": not args.no_graph, "l2": "working set >> L2 (no flush needed)",
                   "gflop_per_sample_fwd_bwd": 3 * flops_forward_per_sample(cfg) / 1e9, "last_loss": last_loss},
        "e2e": {"value": total_samples / (e2e_ms * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": int((hx.numel() + hy.numel() + ht.numel()) * 4), "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms},
        "gpu_launches": launches * args.steps,
        # fwd+bwd + one all-reduce + fused clip/AdamW (FlatAdamW), device-resident inputs: the full training step
        "train_step_with_optimizer": {"ms_per_step": opt_ms, "samples_per_s": total_samples / (opt_ms * 1e-3)},
        "clocks": clocks,
        # dominant kernel: the tcgen05 GEMM with the fused GELU epilogue (async-epilogue variant) at the stage-0 MLP shape, HBM bound
        "roofline": {"bound": "hbm", "achieved": kern_bytes / (kern_us * 1e-6) / 1e9, "peak": peak_hbm, "unit": "GB/s",
                     "frac": kern_bytes / (kern_us * 1e-6) / 1e9 / peak_hbm,
                     # dram__bytes_read+write of this launch from profiles/r01_ncu_full_summary.md (ncu --set full; the
                     # 126 MB L2 still holds part of the 100 MB of output when the kernel ends)
                     "traffic": 55.05e6, "kernel": f"gemm_async_epi_kernel<K-major,GELU> M={Mk} N={Nk} K={Kk}", "us_per_launch": kern_us,
                     "algorithmic_bytes": kern_bytes, "peak_source": hbm_src},
        # whole-step tensor roofline (the BASELINE metric): algorithmic fwd+bwd FLOPs / step time vs measured cuBLAS peak
        "step_tensor_roofline": {"achieved_tflops": achieved_tf, "peak_tflops": peak_tf, "frac": achieved_tf / peak_tf,
                                 "peak_source": peak_src},
    }
    if not args.no_cpu_baseline:
        try:
            c_sps, cores, sec = cpu_reference_leg(cfg, args.cpu_batch, 1, 1)
            rec["cpu_baseline"] = {"value": c_sps, "unit": "samples/s", "cores": cores, "kind": "port",
                                   "sample": f"1 fwd+bwd step of batch {args.cpu_batch} (oracle port, fp32, all host threads) after 1 warm-up"}
        except Exception as ex:  # the baseline must never hide the GPU number
            rec["cpu_baseline"] = {"value": None, "error": repr(ex)[:200]}
    print(json.dumps(rec))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
