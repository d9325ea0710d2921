"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference/scOT/model.py) in the
authoring container — TEST INFRASTRUCTURE. Run:  python oracle/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as small fixtures together with
this script. Three compatibility shims are monkey-patched (no reference file is edited), all forced by
transformers 5.5.0 vs the reference's stale 4.29.2 pin:
  (i)   ScOT.get_head_mask no longer exists on PreTrainedModel,
  (ii)  Swinv2Attention.forward dropped the positional head_mask argument,
  (iii) from_pretrained's meta-device construction is bypassed by building ScOT(config) directly.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from scOT.model import ScOT, ScOTConfig  # noqa: E402  (the reference)
from transformers.models.swinv2.modeling_swinv2 import Swinv2Attention  # noqa: E402

from oracle.weights import make_inputs, make_weights  # noqa: E402

ScOT.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n
_orig = Swinv2Attention.forward
Swinv2Attention.forward = lambda self, hs, mask=None, head_mask=None, output_attentions=False: _orig(self, hs, mask, output_attentions)

GOLD = os.path.join(ROOT, "tests", "golden")

CONFIGS = {
    # small 3-stage model: stage 0 shifted 8x8 windows on a 16x16 grid, stage 1 one 8x8 window, stage 2 4x4
    "tiny": dict(cfg=dict(image_size=64, patch_size=4, num_channels=3, num_out_channels=3, embed_dim=32,
                          depths=[2, 2, 2], num_heads=[2, 4, 8], skip_connections=[1, 1, 0], window_size=8,
                          mlp_ratio=4.0, drop_path_rate=0.0, use_conditioning=True, p=1,
                          channel_slice_list_normalized_loss=[0, 1, 3], residual_model="convnext"),
                 batch=2, mask_channels=(2,), store_all=False),
    # same but unconditioned LayerNorm, plain l1 loss, learn_residual off, no pixel mask
    "tiny_ln": dict(cfg=dict(image_size=64, patch_size=4, num_channels=2, num_out_channels=2, embed_dim=32,
                             depths=[2, 2, 2], num_heads=[2, 4, 8], skip_connections=[1, 0, 0], window_size=8,
                             mlp_ratio=4.0, drop_path_rate=0.0, use_conditioning=False, p=2,
                             channel_slice_list_normalized_loss=None, residual_model="convnext"),
                    batch=2, mask_channels=(), store_all=False),
    # ScOT-T of train.py:36-44 at 128x128, 4 channels (BASELINE.json configs[0] / C1, C2)
    "T128": dict(cfg=dict(image_size=128, patch_size=4, num_channels=4, num_out_channels=4, embed_dim=48,
                          depths=[4, 4, 4, 4], num_heads=[3, 6, 12, 24], skip_connections=[2, 2, 2, 0],
                          window_size=16, mlp_ratio=4.0, drop_path_rate=0.0, use_conditioning=True, p=1,
                          channel_slice_list_normalized_loss=[0, 1, 3, 4], residual_model="convnext"),
                 batch=2, mask_channels=(3,), store_all=False),
    # Poseidon-B of train.py:54-62, 5 channels (BASELINE.json configs[2] / C3), batch 2
    "B128": dict(cfg=dict(image_size=128, patch_size=4, num_channels=5, num_out_channels=5, embed_dim=96,
                          depths=[8, 8, 8, 8], num_heads=[3, 6, 12, 24], skip_connections=[2, 2, 2, 0],
                          window_size=16, mlp_ratio=4.0, drop_path_rate=0.0, use_conditioning=True, p=1,
                          channel_slice_list_normalized_loss=[0, 1, 3, 4, 5], residual_model="convnext"),
                 batch=2, mask_channels=(), store_all=False),
}

# gradients kept for the big configs (full tensors when small, else a leading slice)
GRAD_KEYS = [
    "embeddings.patch_embeddings.projection.weight", "embeddings.norm.weight.weight", "embeddings.norm.bias.bias",
    "encoder.layers.0.blocks.1.attention.self.logit_scale",
    "encoder.layers.0.blocks.1.attention.self.continuous_position_bias_mlp.0.weight",
    "encoder.layers.0.blocks.1.attention.self.continuous_position_bias_mlp.2.weight",
    "encoder.layers.0.blocks.1.attention.self.query.bias", "encoder.layers.0.blocks.1.attention.self.key.weight",
    "encoder.layers.0.blocks.0.intermediate.dense.weight", "encoder.layers.0.blocks.0.layernorm_after.weight.weight",
    "encoder.layers.1.downsample.reduction.weight", "encoder.layers.3.blocks.0.output.dense.weight",
    "decoder.layers.0.upsample.upsample.weight", "decoder.layers.2.upsample.mixup.weight",
    "decoder.layers.3.blocks.0.attention.output.dense.weight", "decoder.layers.3.blocks.0.attention.self.value.bias",
    "residual_blocks.0.0.dwconv.weight", "residual_blocks.0.1.weight", "residual_blocks.1.0.pwconv1.weight",
    "patch_recovery.projection.weight", "patch_recovery.mixup.weight", "patch_recovery.projection.bias",
]


def run(name, spec, dtype):
    cfg = ScOTConfig(**spec["cfg"])
    torch.manual_seed(0)
    model = ScOT(cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    w = make_weights(shapes, seed=0)
    model.load_state_dict(w, strict=True)
    model = model.to(dtype)
    model.train()
    x, t, y, pm = make_inputs(spec["batch"], cfg.num_channels, cfg.num_out_channels, cfg.image_size, seed=0,
                              mask_channels=spec["mask_channels"])
    out = model(pixel_values=x.to(dtype), time=t.to(dtype) if cfg.use_conditioning else None, labels=y.to(dtype),
                pixel_mask=pm if len(spec["mask_channels"]) else None)
    out.loss.backward()
    grads = {k: p.grad.detach() for k, p in model.named_parameters()}
    return cfg, shapes, out, grads


def main():
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or list(CONFIGS)
    for name in which:
        spec = CONFIGS[name]
        cfg, shapes, out, grads = run(name, spec, torch.float64)
        rec = {"config": spec["cfg"], "batch": spec["batch"], "mask_channels": list(spec["mask_channels"]),
               "shapes": shapes, "loss": float(out.loss.detach()), "output": out.output.detach().float(),
               "grad_norms": {k: float(v.norm()) for k, v in grads.items()}}
        if spec["store_all"]:
            rec["grads"] = {k: v.float() for k, v in grads.items()}
        else:
            rec["grads"] = {k: grads[k].float().reshape(-1)[:4096].clone() for k in GRAD_KEYS if k in grads}
        # fp32 run of the reference: its own noise floor against fp64
        _, _, out32, grads32 = run(name, spec, torch.float32)
        rec["ref_fp32_vs_fp64_output_rel_l2"] = float((out32.output.double() - out.output).norm() / out.output.norm())
        rec["ref_fp32_loss"] = float(out32.loss.detach())
        torch.save(rec, os.path.join(GOLD, f"{name}.pt"))
        print(name, "loss", rec["loss"], "out rms", float(out.output.pow(2).mean().sqrt()), "fp32-vs-fp64",
              rec["ref_fp32_vs_fp64_output_rel_l2"], "params", sum(int(torch.tensor(s).prod()) for s in shapes.values()))


if __name__ == "__main__":
    main()
