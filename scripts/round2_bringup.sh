#!/bin/bash
# Round-2 bring-up of the candidates written (but never run) at the end of round 1. One gpurun call; every step has its
# own timeout because a wrong barrier protocol in a candidate kernel shows up as a hang. Reports land in gpurun_out/.
#   gpurun --timeout 900 -- 'bash scripts/round2_bringup.sh'
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/round2_bringup.log
: > $L
step() { echo "=== $*" >> $L; ( time "$@" ) >> $L 2>&1; echo "exit $?" >> $L; }
step timeout 240 python scripts/gemm_v2_check.py
for lvl in 1 3 5 7; do
  step env SCOT_GEMM_ASYNC_V2=$lvl timeout 150 python -m pytest tests/test_gpu_ops.py -x -q -k gemm
  step env SCOT_GEMM_ASYNC_V2=$lvl timeout 150 python -m pytest tests/test_gpu_model.py -x -q -k "forward_matches or smooth"
done
step env SCOT_ATTN_DQ_FOLD2=1 timeout 150 python -m pytest tests/test_gpu_ops.py -x -q -k window_attention
step env SCOT_ATTN_DQ_FOLD2=1 timeout 150 python -m pytest tests/test_gpu_model.py -x -q -k "forward_matches or smooth"
for c in gemm_v2 gemm_v2_2g gemm_v2_smallk gemm_v2_all dq_fold2; do
  step timeout 120 python scripts/ab_overlap.py B 64 20 defaults,$c
done
grep -E "^===|^exit|passed|failed|ms_per_step|equal" $L | tail -n 80
