"""Where does the run-to-run gradient noise of the small test model start? (GPU diagnostic)

Runs N eager replays of one engine on identical inputs, picks the replay that differs most from replay 0 and prints,
in backward execution order (patch recovery -> decoder, finest stage first -> ConvNeXt skips -> encoder, deepest stage
first -> embeddings), the largest relative difference of any parameter gradient of each layer. Process-level knobs
(SCOT_PDL, SCOT_WGRAD_OVERLAP, ...) come from the environment, so run it once per setting.
"""
import json
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from knob_noise import TINY, T128  # noqa: E402
from oracle.weights import make_inputs, make_weights  # noqa: E402


def order_key(name, ns):
    m = re.match(r"decoder\.layers\.(\d+)\.blocks\.(\d+)", name)
    if name.startswith("patch_recovery"):
        return (0, 0, 0)
    if m:
        return (1, -int(m.group(1)), -int(m.group(2)))  # decoder layer ns-1 (finest) first, last block first
    m = re.match(r"decoder\.layers\.(\d+)\.upsample", name)
    if m:
        return (1, -int(m.group(1)) - 0.5, 0)  # after the blocks of the finer layer j+1
    m = re.match(r"residual_blocks\.(\d+)\.(\d+)", name)
    if m:
        return (2, int(m.group(1)), -int(m.group(2)))
    m = re.match(r"encoder\.layers\.(\d+)\.downsample", name)
    if m:
        return (3, -int(m.group(1)), 1)
    m = re.match(r"encoder\.layers\.(\d+)\.blocks\.(\d+)", name)
    if m:
        return (3, -int(m.group(1)), 2 + (100 - int(m.group(2))))
    return (4, 0, 0)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    cfgd, batch = (TINY, 4) if which == "tiny" else (T128, 8)
    from poseidon_b200.runtime import GraphedTrainStep
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    cfg = ScOTConfig(**cfgd)
    model = ScOT(cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(make_weights(shapes, seed=0), strict=True)
    model = model.cuda()
    x, t, y, _ = make_inputs(batch, cfg.num_channels, cfg.num_out_channels, cfg.image_size, seed=0)
    step = GraphedTrainStep(model, batch, torch.device("cuda", 0), use_graph=os.environ.get("TRACE_GRAPH", "0") == "1")
    step.load_batch(x, t, y)
    gs = []
    for _ in range(nrep):
        step.run()
        torch.cuda.synchronize()
        gs.append(step.st["gflat"].clone())
    g0 = gs[0]
    rels = [float((g - g0).norm() / g0.norm()) for g in gs]
    env = {k: os.environ.get(k) for k in ("SCOT_PDL", "SCOT_WGRAD_OVERLAP", "SCOT_CNX_OVERLAP", "SCOT_ATTN_BWD_SPLIT", "TRACE_GRAPH")}
    print(json.dumps({"env": env, "model": which, "rel_vs_replay0": [round(r, 9) for r in rels]}))
    worst = max(range(nrep), key=lambda i: rels[i])
    if rels[worst] < 1e-6:
        print("no replay differs by more than 1e-6")
        return
    table = step.st["engine"].table
    layers = {}
    for name, (off, numel, _s) in table.items():
        a, b = gs[worst][off:off + numel], g0[off:off + numel]
        r = float((a - b).norm() / (b.norm() + 1e-30))
        key = re.sub(r"(\.blocks\.\d+|\.downsample|\.upsample|^residual_blocks\.\d+\.\d+|^patch_recovery|^embeddings).*", r"\1", name)
        cur = layers.get(key)
        if cur is None or r > cur[0]:
            layers[key] = (r, name, order_key(name, len(cfgd["depths"])))
    print(f"replay {worst} vs replay 0, per layer in backward order (max relative difference of a parameter gradient):")
    for key, (r, name, ok) in sorted(layers.items(), key=lambda kv: kv[1][2]):
        print(f"  {r:10.3e}  {key:45s} {name}")


if __name__ == "__main__":
    main()
