"""Records the parameter groups the UNMODIFIED reference builds in `Trainer.create_optimizer`
(/root/reference/scOT/trainer.py:295-400) for the four learning-rate settings — TEST INFRASTRUCTURE.
Run in the authoring container:  python oracle/make_golden_groups.py  -> tests/golden/param_groups.json

The reference's own method is executed on the reference's own model; only the final optimizer construction is
intercepted (`get_optimizer_cls_and_kwargs` is replaced so that the grouped parameters are returned instead of being
handed to torch.optim.AdamW). Parameter NAMES are stored, so the fixture is independent of tensor identity."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from scOT.model import ScOT, ScOTConfig  # noqa: E402  (the reference)
from scOT.trainer import Trainer  # noqa: E402  (the reference)

CFG = dict(image_size=64, patch_size=4, num_channels=3, num_out_channels=3, embed_dim=32, depths=[2, 2, 2],
           num_heads=[2, 4, 8], skip_connections=[1, 1, 0], window_size=8, mlp_ratio=4.0, drop_path_rate=0.0,
           use_conditioning=True, p=1, channel_slice_list_normalized_loss=[0, 1, 3], residual_model="convnext")


def groups_for(model, lr_emb, lr_time):
    fake = object.__new__(Trainer)
    fake.model = model
    fake.model_wrapped = model
    fake.optimizer = None
    fake.args = types.SimpleNamespace(learning_rate_embedding_recovery=lr_emb, learning_rate_time_embedding=lr_time,
                                      weight_decay=0.01)
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
    out = []
    for g in captured["groups"]:
        out.append({"names": sorted(names[id(p)] for p in g["params"]), "weight_decay": g["weight_decay"],
                    "lr": g.get("lr")})
    return out


def main():
    model = ScOT(ScOTConfig(**CFG))
    rec = {"config": CFG, "cases": {}}
    for tag, (a, b) in {"plain": (None, None), "emb": (5e-4, None), "time": (None, 1e-4), "emb_time": (5e-4, 1e-4)}.items():
        rec["cases"][tag] = {"lr_embedding_recovery": a, "lr_time_embedding": b, "groups": groups_for(model, a, b)}
    path = os.path.join(ROOT, "tests", "golden", "param_groups.json")
    with open(path, "w") as f:
        json.dump(rec, f, indent=0)
    print("wrote", path, {k: [len(g["names"]) for g in v["groups"]] for k, v in rec["cases"].items()})


if __name__ == "__main__":
    main()
