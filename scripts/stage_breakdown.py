"""Step time of Poseidon-B truncated to its first 2 / 3 / 4 stages (same widths, batch 64): the differences are the cost
of the deep stages inside the CUDA graph.   python scripts/stage_breakdown.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from poseidon_b200.runtime import GraphedTrainStep  # noqa: E402
from poseidon_b200.scOT.model import ScOT, ScOTConfig  # noqa: E402

B = 64
for ns in (2, 3, 4):
    cfg = bench.model_config("B", 5)
    cfg["depths"] = cfg["depths"][:ns]
    cfg["num_heads"] = cfg["num_heads"][:ns]
    cfg["skip_connections"] = cfg["skip_connections"][:ns - 1] + [0]
    torch.manual_seed(0)
    model = ScOT(ScOTConfig(**cfg))
    bench.realistic_init_(model)
    model = model.cuda()
    step = GraphedTrainStep(model, B)
    g = torch.Generator().manual_seed(1)
    step.load_batch(torch.randn(B, 5, 128, 128, generator=g), torch.rand(B, generator=g), torch.randn(B, 5, 128, 128, generator=g))
    for _ in range(5):
        step.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step.run()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"stages": ns, "ms_per_step": e0.elapsed_time(e1) / 20, "launches": step.launches_per_step()}), flush=True)
    del step, model
    torch.cuda.empty_cache()
