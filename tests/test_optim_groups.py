"""CPU: the parameter groups of poseidon_b200.optim.build_param_groups equal the groups the UNMODIFIED reference's
Trainer.create_optimizer builds (fixture tests/golden/param_groups.json, recorded by oracle/make_golden_groups.py)."""
import json
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "param_groups.json")


@pytest.mark.parametrize("case", ["plain", "emb", "time", "emb_time"])
def test_param_groups_match_reference_trainer(case):
    from poseidon_b200.optim import build_param_groups
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    rec = json.load(open(GOLD))
    with torch.device("meta"):
        model = ScOT(ScOTConfig(**rec["config"]))
    c = rec["cases"][case]
    groups = build_param_groups(model, 0.01, c["lr_embedding_recovery"], c["lr_time_embedding"])
    names = {id(p): n for n, p in model.named_parameters()}
    assert len(groups) == len(c["groups"])
    for mine, ref in zip(groups, c["groups"]):
        assert sorted(names[id(p)] for p in mine["params"]) == ref["names"]
        assert mine["weight_decay"] == ref["weight_decay"]
        assert mine.get("lr") == ref["lr"]
    # every trainable parameter is in exactly one group
    assert sum(len(g["params"]) for g in groups) == len(names)
