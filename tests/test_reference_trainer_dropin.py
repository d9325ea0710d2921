"""CPU, authoring container only (needs /root/reference): the reference's OWN `scOT/trainer.py` runs on top of the
drop-in module. `poseidon_b200.scOT.model` is installed as `scOT.model`, the unmodified reference trainer is imported
against it, and its `Trainer.create_optimizer` (scOT/trainer.py:295-400; relies on `isinstance(module, (nn.LayerNorm,
LayerNorm, ConditionalLayerNorm))` with the classes imported from scOT.model, :230, :282-293) must build exactly the
groups it builds on the reference model (fixture tests/golden/param_groups.json).
The GPU half of the Trainer contract (model(**batch) -> loss.backward() -> clip -> AdamW) is tests/test_gpu_trainer_loop.py."""
import importlib
import json
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "param_groups.json")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scOT")), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref_trainer():
    import poseidon_b200.scOT.model as ours

    saved = {k: sys.modules.get(k) for k in ("scOT", "scOT.model", "scOT.trainer")}
    pkg = types.ModuleType("scOT")
    pkg.__path__ = [os.path.join(REF, "scOT")]  # sub-modules resolve to the reference's files ...
    sys.modules["scOT"] = pkg
    sys.modules["scOT.model"] = ours            # ... except the model, which is the B200 drop-in
    sys.modules.pop("scOT.trainer", None)
    try:
        mod = importlib.import_module("scOT.trainer")
        assert mod.__file__.startswith(REF)
        assert mod.ConditionalLayerNorm is ours.ConditionalLayerNorm and mod.LayerNorm is ours.LayerNorm
        yield mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.parametrize("case", ["plain", "emb", "time", "emb_time"])
def test_reference_create_optimizer_on_the_dropin_model(ref_trainer, case):
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    Trainer = ref_trainer.Trainer
    rec = json.load(open(GOLD))
    c = rec["cases"][case]
    with torch.device("meta"):
        model = ScOT(ScOTConfig(**rec["config"]))
    fake = object.__new__(Trainer)
    fake.model = fake.model_wrapped = model
    fake.optimizer = None
    fake.args = types.SimpleNamespace(learning_rate_embedding_recovery=c["lr_embedding_recovery"],
                                      learning_rate_time_embedding=c["lr_time_embedding"], weight_decay=0.01)
    captured = {}

    def fake_cls(grouped, **kw):
        captured["groups"] = grouped
        return types.SimpleNamespace()

    fake_cls.__name__ = "Captured"
    orig = Trainer.get_optimizer_cls_and_kwargs
    Trainer.get_optimizer_cls_and_kwargs = staticmethod(lambda args, model=None: (fake_cls, {}))
    try:
        Trainer.create_optimizer(fake)
    finally:
        Trainer.get_optimizer_cls_and_kwargs = orig
    names = {id(p): n for n, p in model.named_parameters()}
    assert len(captured["groups"]) == len(c["groups"])
    for mine, ref in zip(captured["groups"], c["groups"]):
        assert sorted(names[id(p)] for p in mine["params"]) == ref["names"]
        assert mine["weight_decay"] == ref["weight_decay"] and mine.get("lr") == ref["lr"]
