"""compute-sanitizer driver: forward + backward of the small golden model through the public API in both precisions, plus one
direct launch of the 16 x 16-window tcgen05 attention kernels (the tiny model has 8 x 8 windows)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from oracle.weights import make_inputs, make_weights
from poseidon_b200.scOT.model import ScOT, ScOTConfig
rec = torch.load("tests/golden/tiny.pt", weights_only=False)
cfg = ScOTConfig(**rec["config"])
for prec in ("bf16", "parity"):
    m = ScOT(cfg); m.load_state_dict(make_weights(rec["shapes"], seed=0)); m = m.cuda(); m.precision = prec; m.use_cuda_graphs = False
    x, t, y, pm = make_inputs(rec["batch"], cfg.num_channels, cfg.num_out_channels, cfg.image_size, seed=0, mask_channels=rec["mask_channels"])
    out = m(pixel_values=x.cuda(), time=t.cuda(), labels=y.cuda(), pixel_mask=pm.cuda() if rec["mask_channels"] else None)
    out.loss.backward(); torch.cuda.synchronize()
    print(prec, "loss", float(out.loss), "ref", rec["loss"])
# the 16 x 16-window tcgen05 attention kernels are not reached by the tiny model (8 x 8 windows): run them directly
from poseidon_b200 import _lib as L
import math
Bn, res, ws, shift, heads, hd = 1, 32, 16, 8, 3, 32
C, M = heads * hd, Bn * res * res
qkv = (torch.randn(M, 3 * C, device="cuda") * 1.5).bfloat16()
cpb = L.CpbLayerBuffers(torch.randn(512, 2, device="cuda"), torch.randn(512, device="cuda") * .1, torch.randn(heads, 512, device="cuda") / 22, math.log(10.) + torch.zeros(heads, 1, 1, device="cuda"), ws, heads)
cpb.forward()
out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16); lse = torch.empty(4 * heads, 256, device="cuda")
L.attn_fwd(qkv, out, lse, cpb.tab2, cpb.alpha, Bn, res, ws, shift, heads, hd)
d_o = torch.randn(M, C, device="cuda").bfloat16(); dqkv = torch.zeros(M, 3 * C, device="cuda", dtype=torch.bfloat16)
L.attn_bwd(qkv, out, d_o, lse, cpb.tab2, cpb.alpha, dqkv, torch.zeros(64, device="cuda"), cpb.dtab, cpb.dalpha, torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"), Bn, res, ws, shift, heads, hd)
torch.cuda.synchronize(); print("attn tc ok", float(dqkv.float().abs().mean()))
