"""A/B check of the engine's side-branch scheduling knobs on a GPU (not a test: prints a report).

  python scripts/ab_overlap.py [model=B] [batch=64] [steps=20] [settings=defaults,inline,cnx_only,attn_only]

Builds one engine per knob setting (the knobs are read from the environment when an engine first runs), feeds all of
them the same seeded batch and weights, and compares prediction (must be bit-identical: the forward pass has no
atomics), loss and the flat gradient (atomics reorder: ~1e-6 relative) against the first setting (the defaults); then times CUDA-graph
replays of zero-grad + forward + backward.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import model_config, realistic_init_  # noqa: E402
from poseidon_b200.runtime import GraphedTrainStep  # noqa: E402
from poseidon_b200.scOT.model import ScOT, ScOTConfig  # noqa: E402

OFF = {"SCOT_CNX_OVERLAP": "1", "SCOT_ATTN_BWD_SPLIT": "16"}
SETTINGS = [
    ("defaults", {}),  # must stay first: the reference everything else is compared with
    ("inline", {"SCOT_CNX_OVERLAP": "0", "SCOT_ATTN_BWD_SPLIT": "0"}),
    ("cnx_only", {"SCOT_ATTN_BWD_SPLIT": "0"}),
    ("attn_only", {"SCOT_CNX_OVERLAP": "0"}),
    # round-2 candidates (never run on hardware when this was written): run them one per process under a timeout,
]
DEFAULT_RUN = ("defaults", "inline", "cnx_only", "attn_only")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "B"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    chosen = tuple(sys.argv[4].split(",")) if len(sys.argv) > 4 else DEFAULT_RUN
    dev = torch.device("cuda", 0)
    cfg = model_config(name, 5)
    gen = torch.Generator().manual_seed(7)
    S = cfg["image_size"]
    hx = torch.randn(batch, 5, S, S, generator=gen)
    hy = torch.randn(batch, 5, S, S, generator=gen)
    ht = torch.rand(batch, generator=gen)
    ref = None
    report = []
    for label, env in [x for x in SETTINGS if x[0] in chosen]:
        os.environ.update(OFF)
        os.environ.update(env)
        torch.manual_seed(0)
        model = ScOT(ScOTConfig(**cfg))
        realistic_init_(model)
        model = model.to(dev)
        rec = {"setting": label}
        for use_graph in (False, True):
            step = GraphedTrainStep(model, batch, dev, use_graph=use_graph)
            step.load_batch(hx, ht, hy)
            step.run()
            torch.cuda.synchronize()
            out = (step.pred.clone(), float(step.loss), step.st["gflat"].clone())
            tag = "graph" if use_graph else "eager"
            if ref is None:
                ref = out
            else:
                rec[f"{tag}_pred_equal"] = bool(torch.equal(out[0], ref[0]))
                rec[f"{tag}_pred_maxdiff"] = float((out[0] - ref[0]).abs().max())
                rec[f"{tag}_loss_diff"] = abs(out[1] - ref[1])
                rec[f"{tag}_grad_rel"] = float((out[2] - ref[2]).norm() / ref[2].norm())
                rec[f"{tag}_grad_finite"] = bool(torch.isfinite(out[2]).all())
            if use_graph:
                for _ in range(3):
                    step.run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    step.run()
                e1.record()
                torch.cuda.synchronize()
                rec["ms_per_step"] = e0.elapsed_time(e1) / steps
                # repeat the comparison after the timed replays (a race may need several runs to show)
                out2 = (step.pred.clone(), float(step.loss), step.st["gflat"].clone())
                rec["replay_pred_equal"] = bool(torch.equal(out2[0], ref[0]))
                rec["replay_grad_rel"] = float((out2[2] - ref[2]).norm() / ref[2].norm())
            del step
        print(json.dumps(rec), flush=True)
        report.append(rec)
        del model
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"ab_overlap_{name}{batch}_{chosen[-1]}.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
