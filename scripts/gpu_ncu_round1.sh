#!/bin/bash
# ncu --set full captures (source-level stall sampling) of the async-epilogue GEMM family and the stage-0 attention backward
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 45 ncu --set full --clock-control none --import-source on -k regex:gemm_async_epi -s 3 -c 3 -f -o gpurun_out/r01_gemm_async_full python scripts/gemm_one.py > gpurun_out/ncu_gemm_full.log 2>&1
echo "gemm ncu exit $?" >> gpurun_out/ncu_gemm_full.log
timeout 60 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -c 2 -f -o gpurun_out/r01_attn_bwd_full python scripts/profile_step.py B 64 1 > gpurun_out/ncu_attn_full.log 2>&1
echo "attn ncu exit $?" >> gpurun_out/ncu_attn_full.log
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_gemm_full.log gpurun_out/ncu_attn_full.log
