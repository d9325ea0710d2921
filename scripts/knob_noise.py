"""Run-to-run noise floor of the gradients vs. the effect of each scheduling knob (GPU; prints a report).

Two engines built with identical settings do not produce bit-identical gradients: split reductions (`+=` GEMMs with
split K, weight gradients, column sums) use fp32 atomics whose order changes from run to run, and an fp32 last-bit
difference occasionally flips the bf16 rounding of an operand of the next GEMM. This script measures that floor on the
small all-block-types model of tests/test_gpu_knobs.py (and optionally Poseidon-T at 128x128) and compares every knob
against it, with a per-parameter breakdown of the largest differences.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.weights import make_inputs, make_weights  # noqa: E402  (seeded inputs / weights only)

KNOBS = {"SCOT_CNX_OVERLAP": "0", "SCOT_ATTN_BWD_SPLIT": "0"}
TINY = dict(image_size=64, patch_size=4, num_channels=3, num_out_channels=3, embed_dim=32, depths=[2, 2, 2],
            num_heads=[2, 4, 8], skip_connections=[2, 1, 0], window_size=8, mlp_ratio=4.0, drop_path_rate=0.0,
            use_conditioning=True, p=1, channel_slice_list_normalized_loss=[0, 1, 3], residual_model="convnext")
T128 = dict(image_size=128, patch_size=4, num_channels=4, num_out_channels=4, embed_dim=48, depths=[4, 4, 4, 4],
            num_heads=[3, 6, 12, 24], skip_connections=[2, 2, 2, 0], window_size=16, mlp_ratio=4.0, drop_path_rate=0.0,
            use_conditioning=True, p=1, channel_slice_list_normalized_loss=[0, 1, 3, 4], residual_model="convnext")


def run(cfgd, env, use_graph, batch, nrep=2):
    from poseidon_b200.runtime import GraphedTrainStep
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    os.environ.update(KNOBS)
    os.environ.update(env)
    cfg = ScOTConfig(**cfgd)
    model = ScOT(cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(make_weights(shapes, seed=0), strict=True)
    model = model.cuda()
    x, t, y, _ = make_inputs(batch, cfg.num_channels, cfg.num_out_channels, cfg.image_size, seed=0)
    step = GraphedTrainStep(model, batch, torch.device("cuda", 0), use_graph=use_graph)
    step.load_batch(x, t, y)
    outs = []
    for _ in range(nrep):
        step.run()
        torch.cuda.synchronize()
        outs.append((step.pred.clone(), float(step.loss), step.st["gflat"].clone()))
    table = dict(step.st["engine"].table)
    return outs, table


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def worst_params(g, g0, table, k=6):
    rows = []
    for name, (off, numel, _shape) in table.items():
        a, b = g[off:off + numel], g0[off:off + numel]
        rows.append((float((a - b).norm()), rel(a, b), name))
    rows.sort(reverse=True)
    return [{"name": n, "abs": round(a, 6), "rel": round(r, 6)} for a, r, n in rows[:k]]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    cfgd, batch = (TINY, 4) if which == "tiny" else (T128, 8)
    report = {"model": which, "batch": batch}
    for use_graph in (False, True):
        tag = "graph" if use_graph else "eager"
        base, table = run(cfgd, {}, use_graph, batch, nrep=3)
        base2, _ = run(cfgd, {}, use_graph, batch, nrep=1)
        g0 = base[0][2]
        report[f"{tag}_noise_replay"] = [rel(base[i][2], g0) for i in (1, 2)]
        report[f"{tag}_noise_new_engine"] = rel(base2[0][2], g0)
        report[f"{tag}_noise_worst"] = worst_params(base2[0][2], g0, table, 4)
        for label, env in [("cnx", {"SCOT_CNX_OVERLAP": "1"}), ("attn", {"SCOT_ATTN_BWD_SPLIT": "16"}),
                           ("defaults", {"SCOT_CNX_OVERLAP": "1", "SCOT_ATTN_BWD_SPLIT": "16"})]:
            outs, _ = run(cfgd, env, use_graph, batch, nrep=2)
            report[f"{tag}_{label}"] = {
                "pred_equal": [bool(torch.equal(o[0], base[0][0])) for o in outs],
                "grad_rel": [rel(o[2], g0) for o in outs],
                "worst": worst_params(outs[0][2], g0, table, 4),
            }
    print(json.dumps(report, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"knob_noise_{which}.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
