#!/usr/bin/env python
"""Benchmark of the scOT hot path (BASELINE.json: samples/sec fwd+bwd, Poseidon-B 128x128; % tensor roofline).

  python bench.py --gpus N --steps K --warmup W            # our engine (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port) on host cores

A step = zero grads + forward + backward of Poseidon-B on one synthetic batch (64 x 5 x 128 x 128 per GPU,
bf16 GEMM operands / fp32 accumulation & residual stream) [+ one NCCL all-reduce of the flat gradient
buffer when N > 1]. Prints ONE JSON line (rank 0). See DESIGN.md "Measurement".

  python bench.py --config C2|C3|C4|C5    # the other BASELINE.json configurations (C3 = default = the metric's config):
      C2 Poseidon-T b32 4ch (1 GPU) | C4 Poseidon-L b32/GPU (8 GPUs under torchrun) | C5 Poseidon-L 256x256 20-step rollout
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


CONFIGS = {  # BASELINE.json `configs`: model, per-GPU batch, channels, image size, kind
    "C2": dict(model="T", batch=32, channels=4, size=128, kind="train"),
    "C3": dict(model="B", batch=64, channels=5, size=128, kind="train"),
    "C4": dict(model="L", batch=32, channels=5, size=128, kind="train"),
    "C5": dict(model="L", batch=8, channels=5, size=256, kind="rollout", ar_steps=20),
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


def gpu_eager_baseline(cfg: dict, batch: int, dev, steps: int = 2):
    """The "library Blackwell" baseline (SURVEY.md section 8d last row): the reference's own eager code path — the
    oracle port, i.e. the same torch ops the reference issues (cuBLAS / ATen kernels) — on the GPU, fp32 and bf16
    autocast, same batch. Reported beside the engine's number; neither is the optimisation target."""
    import types

    from oracle import scot_oracle as O
    from oracle.weights import make_inputs, make_weights

    ocfg = types.SimpleNamespace(**cfg)
    ocfg.layer_norm_eps, ocfg.learn_residual = 1e-5, False
    w = {k: v.to(dev).requires_grad_(True) for k, v in make_weights(param_shapes(cfg), seed=0).items()}
    x, t, y, _ = make_inputs(batch, cfg["num_channels"], cfg["num_out_channels"], cfg["image_size"], seed=0)
    x, t, y = x.to(dev), t.to(dev), y.to(dev)
    out = {}
    for name, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        def one():
            for v in w.values():
                v.grad = None
            if ctx is None:
                loss, _ = O.scot_forward(ocfg, w, x, t, y, None)
            else:
                with ctx:
                    loss, _ = O.scot_forward(ocfg, w, x, t, y, None)
            loss.backward()
        try:
            one()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": ms, "samples_per_s": batch / (ms * 1e-3)}
        except Exception as ex:  # e.g. out of memory at a large batch: the baseline must never hide the engine's number
            out[name] = {"error": repr(ex)[:160]}
        torch.cuda.empty_cache()
    out["what"] = f"oracle port (torch eager, the reference's op sequence) on cuda, batch {batch}, {steps} timed steps after 1 warm-up"
    return out


def time_kernel(fn, n_sets: int, iters: int, dev):
    """CUDA-event timing of `fn(i)` on the launching stream; `fn` cycles through n_sets buffer sets larger than L2"""
    for i in range(2 * n_sets):
        fn(i)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / iters * 1e3  # us


def kernel_rooflines(cfg: dict, B: int, dev, peaks: dict):
    """Times the three heaviest kernel families of the step alone (inputs cycled over > L2 worth of buffers) and returns
    their roofline records sorted by (time per launch x launches per step): the first one is the step's dominant kernel."""
    import math as _m

    from poseidon_b200 import _lib as L

    C0, heads0 = cfg["embed_dim"], cfg["num_heads"][0]
    hd = C0 // heads0
    res0 = cfg["image_size"] // cfg["patch_size"]
    M0 = B * res0 * res0
    depth0, depth1 = cfg["depths"][0], cfg["depths"][1]
    recs = []
    # (1) window attention backward, stage 0 (16 x 16 windows): tensor bound (S, dP, dQ, dK, dV = 10 N^2 hd FLOP per unit)
    nset = 3
    sets = []
    for _ in range(nset):
        qkv = (torch.randn(M0, 3 * C0, device=dev) * 1.5).bfloat16()
        sets.append(dict(qkv=qkv, out=torch.empty(M0, C0, device=dev, dtype=torch.bfloat16),
                         d_o=torch.randn(M0, C0, device=dev).bfloat16(),
                         dqkv=torch.empty(M0, 3 * C0, device=dev, dtype=torch.bfloat16)))
    w1, b1 = torch.randn(512, 2, device=dev), torch.randn(512, device=dev) * 0.1
    w2 = torch.randn(heads0, 512, device=dev) / 512 ** 0.5
    ls = _m.log(10.0) + 0.3 * torch.randn(heads0, 1, 1, device=dev)
    cpb = L.CpbLayerBuffers(w1, b1, w2, ls, 16, heads0)
    cpb.forward()
    nwin = B * (res0 // 16) ** 2
    lse = torch.empty(nwin * heads0, 256, device=dev)
    partial = torch.zeros(64, device=dev)
    gq, gv = torch.zeros(C0, device=dev), torch.zeros(C0, device=dev)
    for st_ in sets:
        L.attn_fwd(st_["qkv"], st_["out"], lse, cpb.tab2, cpb.alpha, B, res0, 16, 0, heads0, hd)

    def attn_bwd(i):
        st_ = sets[i % nset]
        L.attn_bwd(st_["qkv"], st_["out"], st_["d_o"], lse, cpb.tab2, cpb.alpha, st_["dqkv"], partial, cpb.dtab, cpb.dalpha,
                   gq, gv, B, res0, 16, 0, heads0, hd)

    def attn_fwd(i):
        st_ = sets[i % nset]
        L.attn_fwd(st_["qkv"], st_["out"], lse, cpb.tab2, cpb.alpha, B, res0, 16, 0, heads0, hd)

    units = nwin * heads0
    us = time_kernel(attn_bwd, nset, 30, dev)
    fl = 10.0 * 256 * 256 * hd * units
    # stage 1 has a quarter of the tokens and twice the heads: half the units -> half the time per launch
    recs.append({"ncu_name": f"attn_tc_bwd_kernel<{hd}>",
                 "kernel": f"attn_tc_bwd_kernel<{hd}> (window attention backward, stage 0: {units} (window, head) units)",
                 "bound": "tensor", "achieved": fl / (us * 1e-6) / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                 "us_per_launch": us, "algorithmic_flops": fl, "launches_per_step": 2 * depth0 + depth1,
                 "launch_equiv_note": "stage-1 launches counted as half a stage-0 launch", "peak_source": peaks["tf_src"]})
    us = time_kernel(attn_fwd, nset, 30, dev)
    fl = 4.0 * 256 * 256 * hd * units
    recs.append({"ncu_name": f"attn_tc_fwd_kernel<{hd}>",
                 "kernel": f"attn_tc_fwd_kernel<{hd}> (window attention forward, stage 0)", "bound": "tensor",
                 "achieved": fl / (us * 1e-6) / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "us_per_launch": us,
                 "algorithmic_flops": fl, "launches_per_step": 2 * depth0 + depth1, "peak_source": peaks["tf_src"]})
    del sets
    # (2) stage-0 MLP up-projection with the fused GELU epilogue: HBM bound (K = C: ~65 FLOP/B)
    Mk, Nk, Kk = M0, 4 * C0, C0
    gsets = [(torch.randn(Mk, Kk, device=dev).bfloat16(), torch.randn(Nk, Kk, device=dev).bfloat16(), torch.randn(Nk, device=dev),
              torch.empty(Mk, Nk, device=dev, dtype=torch.bfloat16), torch.empty(Mk, Nk, device=dev, dtype=torch.bfloat16))
             for _ in range(3)]

    def gelu_gemm(i):
        a_, b_, bias_, o0, o1 = gsets[i % 3]
        L.gemm(a_, b_, Mk, Nk, Kk, mode=L.EPI_GELU, bias=bias_, out0=o0, out1=o1)

    us = time_kernel(gelu_gemm, 3, 30, dev)
    by = (Mk * Kk + Nk * Kk) * 2 + Nk * 4 + 2 * Mk * Nk * 2
    recs.append({"ncu_name": "gemm_async_epi_kernel<0, 2>",
                 "kernel": f"gemm_async_epi_kernel<K-major,GELU> M={Mk} N={Nk} K={Kk}", "bound": "hbm",
                 "achieved": by / (us * 1e-6) / 1e9, "peak": peaks["hbm"], "unit": "GB/s", "us_per_launch": us,
                 "algorithmic_bytes": by, "launches_per_step": 2 * depth0 + 2, "peak_source": peaks["hbm_src"]})
    del gsets
    # (3) the four weight-gradient GEMMs of a stage-0 block in one grouped launch: HBM bound (reads every dY / X once)
    H0 = 4 * C0
    wsets = []
    for _ in range(2):
        mk = lambda n: torch.randn(M0, n, device=dev).bfloat16()
        wsets.append([(mk(C0), mk(H0), torch.zeros(C0, H0, device=dev)), (mk(H0), mk(C0), torch.zeros(H0, C0, device=dev)),
                      (mk(C0), mk(C0), torch.zeros(C0, C0, device=dev)), (mk(3 * C0), mk(C0), torch.zeros(3 * C0, C0, device=dev))])

    def wgrad(i):
        L.wgrad_group(wsets[i % 2])

    us = time_kernel(wgrad, 2, 20, dev)
    by = M0 * (C0 + H0 + H0 + C0 + C0 + C0 + 3 * C0 + C0) * 2
    recs.append({"ncu_name": "gemm_tc_kernel<128, 1, 1, 5>",
                 "kernel": f"gemm_tc_kernel<128,MN,MN,atomic> grouped weight gradients of a stage-0 block ({M0} tokens)",
                 "bound": "hbm", "achieved": by / (us * 1e-6) / 1e9, "peak": peaks["hbm"], "unit": "GB/s", "us_per_launch": us,
                 "algorithmic_bytes": by, "launches_per_step": 2 * depth0, "peak_source": peaks["hbm_src"]})
    del wsets
    torch.cuda.empty_cache()
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")  # dram bytes per launch from ncu --set full captures
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for r in recs:
        r["frac"] = r["achieved"] / r["peak"]
        r["step_us"] = r["us_per_launch"] * r["launches_per_step"]
        key = r["kernel"].split("<")[0].split(" ")[0]
        r["traffic"] = traffic.get(r.pop("ncu_name", key), traffic.get(key))  # DRAM read + write bytes of the same launch under ncu
    recs.sort(key=lambda r: -r["step_us"])
    return recs


def run_rollout_config(args, cfg, dev, rank, world, c):
    """C5: autoregressive inference rollout (reference scOT/trainer.py:452-603): ar_steps forwards, prediction fed back"""
    from poseidon_b200.runtime import ARRollout
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    torch.manual_seed(0)
    model = ScOT(ScOTConfig(**cfg))
    realistic_init_(model)
    model = model.to(dev)
    model.precision = args.precision
    B, S, steps_ar = args.batch, cfg["image_size"], c["ar_steps"]
    ro = ARRollout(model, B, dev)
    gen = torch.Generator().manual_seed(1234 + rank)
    hx = torch.randn(B, args.channels, S, S, generator=gen).pin_memory()
    ht = torch.rand(B, generator=gen).pin_memory()
    x, t = hx.to(dev), ht.to(dev)
    for _ in range(max(args.warmup, 3)):
        ro.run(x, t, steps_ar)
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ro.run(x, t, steps_ar)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / args.steps
    t0 = time.perf_counter()
    n = max(3, args.steps // 2)
    for _ in range(n):
        out, _ = ro.run(hx.to(dev, non_blocking=True), ht.to(dev, non_blocking=True), steps_ar)
        chk = float(out.abs().mean())  # D2H read of the result
    e2e_ms = (time.perf_counter() - t0) / n * 1e3
    clocks = sampler.stop()
    f_fwd = B * flops_forward_per_sample(cfg) + flops_cpb_per_step(cfg)
    pk = load_peaks()
    rec = {"metric": "rollout samples x steps / sec (forward only)", "value": B * steps_ar / (ms * 1e-3), "unit": "sample-steps/s",
           "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "split-bf16", "data": "synthetic",
           "config": {"workload": f"{args.config}: Poseidon-{args.model} {steps_ar}-step autoregressive rollout, {args.channels}ch "
                                  f"{S}x{S}, batch {B}", "ms_per_forward": ms / steps_ar, "cuda_graph": True,
                      "l2": "working set >> L2 (no flush needed)", "mean_abs_output": chk},
           "e2e": {"value": B * steps_ar / (e2e_ms * 1e-3), "unit": "sample-steps/s", "h2d_bytes_per_step": int((hx.numel() + ht.numel()) * 4),
                   "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms},
           "gpu_launches": None, "clocks": clocks,
           "step_tensor_roofline": {"achieved_tflops": f_fwd * steps_ar / (ms * 1e-3) / 1e12, "peak_tflops": pk["tf_sustained"],
                                    "frac": f_fwd * steps_ar / (ms * 1e-3) / 1e12 / pk["tf_sustained"], "peak_source": pk["tf_sus_src"]}}
    print(json.dumps(rec))


def load_peaks():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        return {"tf_sustained": float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1590.0))),
                "tf_sus_src": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)",
                "tf_burst": float(pk.get("bf16_tflops", 1590.0)), "tf_src": "MEASURED_PEAKS.json bf16_tflops (burst, of measured)",
                "hbm": float(pk.get("hbm_gbs", 6650.0)), "hbm_src": "MEASURED_PEAKS.json hbm_gbs (of measured)"}
    return {"tf_sustained": 1400.0, "tf_sus_src": "fallback (B200_PROFILING.md sustained)", "tf_burst": 1590.0,
            "tf_src": "fallback (B200_PROFILING.md)", "hbm": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)"}


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
    ap.add_argument("--config", default=None, choices=list(CONFIGS), help="BASELINE.json configuration (default: C3)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "parity"])
    ap.add_argument("--no-eager-baseline", action="store_true")
    args = ap.parse_args()
    size = 128
    kind = "train"
    if args.config is not None:
        c = CONFIGS[args.config]
        args.model, args.batch, args.channels, size, kind = c["model"], c["batch"], c["channels"], c["size"], c["kind"]
    cfg = model_config(args.model, args.channels, size)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"Poseidon-{args.model} fwd+bwd, {args.channels}ch {size}x{size}, batch {args.batch}/GPU"

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
        sps, cores, sec = cpu_reference_leg(cfg, args.cpu_batch, steps, warm)
        print(json.dumps({
            "impl": "reference", "metric": "samples/sec (fwd+bwd)", "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Poseidon-{args.model} fwd+bwd, {args.channels}ch {size}x{size}, batch {args.cpu_batch} "
                                   f"(bounded CPU sample of the batch-{args.batch}/GPU workload)",
                       "note": "reference CPU path = oracle port (torch eager fp32) on host cores"},
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

    if kind == "rollout":
        if rank == 0:
            run_rollout_config(args, cfg, dev, rank, world, CONFIGS[args.config])
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    torch.manual_seed(0)
    model = ScOT(ScOTConfig(**cfg))
    realistic_init_(model)
    model = model.to(dev)
    model.precision = args.precision
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

    # ---- rooflines of the heaviest kernel families, timed alone right here (CUDA events on the launching stream, buffer
    # sets cycled so that no launch finds its inputs in the 126 MB L2); the first record is the step's dominant kernel
    pk = load_peaks()
    del step, opt
    torch.cuda.empty_cache()
    kernels = kernel_rooflines(cfg, B, dev, pk)
    eager = None
    if not args.no_eager_baseline and args.config in (None, "C3", "C2"):
        model._state = None
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline(cfg, B, dev)

    total_samples = B * world
    sps = total_samples / (ms * 1e-3)
    f_step = 3.0 * (B * flops_forward_per_sample(cfg) + flops_cpb_per_step(cfg))  # per GPU
    peak_tf, peak_src = pk["tf_sustained"], pk["tf_sus_src"]
    achieved_tf = f_step / (ms * 1e-3) / 1e12
    rec = {
        "metric": "samples/sec (fwd+bwd)", "value": sps, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "split-bf16", "data": "synthetic",
        "config": {"workload": (args.config + ": " if args.config else "") + workload, "global_batch": total_samples,
                   "parallelism": f"dp{world}", "precision": args.precision,
                   "cuda_graph": not args.no_graph, "l2": "working set >> L2 (no flush needed)",
                   "gflop_per_sample_fwd_bwd": 3 * flops_forward_per_sample(cfg) / 1e9, "last_loss": last_loss},
        "e2e": {"value": total_samples / (e2e_ms * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": int((hx.numel() + hy.numel() + ht.numel()) * 4), "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms},
        "gpu_launches": launches * args.steps,
        # fwd+bwd + one all-reduce + fused clip/AdamW (FlatAdamW), device-resident inputs: the full training step
        "train_step_with_optimizer": {"ms_per_step": opt_ms, "samples_per_s": total_samples / (opt_ms * 1e-3)},
        "clocks": clocks,
        # dominant kernel of the step (largest time per launch x launches per step among the families timed above)
        "roofline": {k: kernels[0].get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "us_per_launch",
                                                    "algorithmic_flops", "algorithmic_bytes", "launches_per_step", "step_us",
                                                    "peak_source") if k in kernels[0]},
        "roofline_other_kernels": kernels[1:],
        # whole-step tensor roofline (the BASELINE metric): algorithmic fwd+bwd FLOPs / step time vs measured cuBLAS peak
        "step_tensor_roofline": {"achieved_tflops": achieved_tf, "peak_tflops": peak_tf, "frac": achieved_tf / peak_tf,
                                 "peak_source": peak_src},
    }
    if eager is not None:
        rec["gpu_eager_baseline"] = eager
    if not args.no_cpu_baseline:
        try:
            c_sps, cores, sec = cpu_reference_leg(cfg, args.cpu_batch, 2, 1)
            rec["cpu_baseline"] = {"value": c_sps, "unit": "samples/s", "cores": cores, "kind": "port",
                                   "sample": f"2 fwd+bwd steps of batch {args.cpu_batch} (oracle port, fp32, all host threads, capped at 32) after 1 warm-up"}
        except Exception as ex:  # the baseline must never hide the GPU number
            rec["cpu_baseline"] = {"value": None, "error": repr(ex)[:200]}
    print(json.dumps(rec))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
