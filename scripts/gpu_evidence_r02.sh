#!/bin/bash
# Round-2 evidence in one gpurun call: launch list of one eager step, memcheck / racecheck of the small golden model,
# ncu --set full captures of the kernels the bench names. Everything lands in gpurun_out/ (summaries are copied to profiles/).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# (1) every launch of one eager fwd+bwd step of Poseidon-B batch 64 (second step: warm caches / attributes set)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1137 -c 1400 --csv --log-file gpurun_out/r02_launches_step.csv \
  python scripts/profile_step.py B 64 2 > gpurun_out/r02_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/r02_launches.log
# (2) compute-sanitizer on the small golden model (forward + backward through the public API), bf16 and parity precision
cat > /tmp/san_tiny.py <<'PY'
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
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san_tiny.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san_tiny.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_sanitizer_racecheck.log
# (3) ncu --set full: attention fwd/bwd (tcgen05), LayerNorm fwd/bwd, depthwise conv, conv5, adamw, wgrad GEMM
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'attn_tc_bwd|attn_tc_fwd' -s 4 -c 2 -f -o gpurun_out/r02_attn_tc_full python scripts/attn_bench.py both > gpurun_out/ncu_attn_tc.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'cln_fwd|cln_bwd|dwconv7|conv5|gemm_tc_kernel' -s 60 -c 40 -f -o gpurun_out/r02_misc_full python scripts/profile_step.py B 64 1 > gpurun_out/ncu_misc.log 2>&1
tail -2 gpurun_out/r02_launches.log gpurun_out/r02_sanitizer_memcheck.log gpurun_out/r02_sanitizer_racecheck.log
