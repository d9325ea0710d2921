"""GPU bring-up: engine forward/backward vs the CPU oracle (fp64) on seeded inputs; per-parameter gradient
errors. Writes gpurun_out/model_check_<name>.json.  usage: gpu_model_check.py [tiny tiny_ln T128 B128] [--simt]"""
import json, os, sys, time, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poseidon_b200 import _lib as L
from poseidon_b200.scOT.model import ScOT, ScOTConfig
from oracle import scot_oracle as O
from oracle.weights import make_inputs, make_weights

def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))

def run(name, impl, ref_dtype=torch.float64):
    rec = torch.load(os.path.join(ROOT, "tests", "golden", f"{name}.pt"), weights_only=False)
    cfg = ScOTConfig(**rec["config"])
    w = make_weights(rec["shapes"], seed=0)
    x, t, y, pm = make_inputs(rec["batch"], cfg.num_channels, cfg.num_out_channels, cfg.image_size, seed=0,
                              mask_channels=rec["mask_channels"])
    model = ScOT(cfg)
    model.load_state_dict(w, strict=True)
    model = model.cuda()
    model.gemm_impl = impl
    t0 = time.time()
    out = model(pixel_values=x.cuda(), time=t.cuda() if cfg.use_conditioning else None, labels=y.cuda(),
                pixel_mask=pm.cuda() if rec["mask_channels"] else None)
    torch.cuda.synchronize()
    t_fwd = time.time() - t0
    rep = {"name": name, "impl": impl, "fwd_s": t_fwd}
    rep["out_rel_vs_fixture"] = rel(out.output.cpu(), rec["output"])
    rep["loss"] = float(out.loss); rep["loss_ref"] = rec["loss"]
    rep["out_finite"] = bool(torch.isfinite(out.output).all())
    t0 = time.time()
    gen = torch.Generator().manual_seed(123)
    G = torch.randn(out.output.shape, generator=gen)
    if LIN:
        # smooth objective <G, pred>: free of the sign() discontinuity of the L1 loss
        out.output.backward(G.cuda())
    else:
        out.loss.backward()
    torch.cuda.synchronize()
    rep["bwd_s"] = time.time() - t0
    grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
    # oracle gradients (fp64 CPU)
    t0 = time.time()
    wr = {k: v.to(ref_dtype).requires_grad_(True) for k, v in w.items()}
    ocfg = types.SimpleNamespace(**rec["config"]); ocfg.learn_residual = False
    if not hasattr(ocfg, "layer_norm_eps"): ocfg.layer_norm_eps = 1e-5
    loss, pred = O.scot_forward(ocfg, wr, x.to(ref_dtype), t.to(ref_dtype) if cfg.use_conditioning else None, y.to(ref_dtype),
                                pm if rec["mask_channels"] else None)
    if LIN:
        (pred * G.to(ref_dtype)).sum().backward()
    else:
        loss.backward()
    rep["oracle_s"] = time.time() - t0
    rep["out_rel_vs_oracle"] = rel(out.output.cpu(), pred.detach())
    errs = {k: rel(grads[k], wr[k].grad) for k in grads}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:25]
    rep["grad_rel_worst"] = worst
    gn = torch.sqrt(sum((grads[k].double() - wr[k].grad).pow(2).sum() for k in grads)) / torch.sqrt(sum(wr[k].grad.pow(2).sum() for k in grads))
    rep["grad_rel_global"] = float(gn)
    rep["grad_rel_median"] = float(torch.tensor(list(errs.values())).median())
    # by kind
    kinds = {}
    for k, v in errs.items():
        kind = ".".join(p for p in k.split(".") if not p.isdigit())
        kinds.setdefault(kind, []).append(v)
    rep["grad_rel_by_kind"] = {k: max(v) for k, v in sorted(kinds.items())}
    return rep

LIN = "--lin" in sys.argv
names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["tiny"]
impl = L.GEMM_SIMT if "--simt" in sys.argv else L.GEMM_TCGEN05
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for n in names:
    try:
        rep = run(n, impl)
    except Exception as ex:
        import traceback
        rep = {"name": n, "exc": traceback.format_exc()[-3000:]}
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", f"model_check_{n}_{impl}_{'lin' if LIN else 'loss'}.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in rep.items() if k != "grad_rel_by_kind"}, indent=1)[:6000], flush=True)
    if "grad_rel_by_kind" in rep:
        print("BY KIND:", json.dumps(rep["grad_rel_by_kind"], indent=0))
