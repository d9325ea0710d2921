"""Other BASELINE.json configs on the GPU: Poseidon-T training step (C2), Poseidon-L forward at 128^2 and 256^2 (C4/C5 shapes)
against the CPU oracle (fp32), plus timing of the forward pass."""
import json, os, sys, time, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from poseidon_b200.scOT.model import ScOT, ScOTConfig
from oracle import scot_oracle as O
from oracle.weights import make_inputs, make_weights

def rel(a, b): return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))
torch.set_num_threads(min(32, len(os.sched_getaffinity(0))))
for (name, ch, size, batch, do_bwd) in [("T", 4, 128, 4, True), ("L", 5, 128, 1, True), ("L", 5, 256, 1, False)]:
    cfg = bench.model_config(name, ch, size)
    with torch.device("meta"):
        shapes = {k: tuple(v.shape) for k, v in ScOT(ScOTConfig(**cfg)).state_dict().items()}
    w = make_weights(shapes, seed=0)
    model = ScOT(ScOTConfig(**cfg)); model.load_state_dict(w, strict=True); model = model.cuda()
    x, t, y, pm = make_inputs(batch, ch, ch, size, seed=0)
    rep = {"model": name, "size": size, "batch": batch}
    try:
        out = model(pixel_values=x.cuda(), time=t.cuda(), labels=y.cuda())
        if do_bwd:
            out.loss.backward()
            gn = torch.sqrt(sum(p.grad.double().pow(2).sum() for p in model.parameters()))
            rep["grad_norm"] = float(gn); rep["grad_finite"] = bool(torch.isfinite(gn))
        torch.cuda.synchronize()
        ocfg = types.SimpleNamespace(**cfg); ocfg.layer_norm_eps, ocfg.learn_residual = 1e-5, False
        t0 = time.time()
        with torch.no_grad():
            loss, pred = O.scot_forward(ocfg, w, x, t, y, None)
        rep["oracle_s"] = round(time.time() - t0, 1)
        rep["out_rel"] = rel(out.output.cpu(), pred); rep["loss"] = float(out.loss.detach()); rep["loss_ref"] = float(loss)
        with torch.no_grad():
            xs, ts, ys = x.cuda(), t.cuda(), y.cuda()
            for _ in range(2): model(pixel_values=xs, time=ts, labels=ys)
            torch.cuda.synchronize(); t0 = time.time()
            for _ in range(5): model(pixel_values=xs, time=ts, labels=ys)
            torch.cuda.synchronize(); rep["fwd_ms"] = round((time.time() - t0) / 5 * 1e3, 2)
    except Exception as ex:
        import traceback; rep["exc"] = traceback.format_exc()[-1500:]
    print(json.dumps(rep), flush=True)
    del model; torch.cuda.empty_cache()
