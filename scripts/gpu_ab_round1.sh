#!/bin/bash
# One gpurun call: A/B report of the scheduling / kernel-variant knobs, the knob neutrality tests, the affected per-op
# and whole-model parity tests with every knob on, and a bench line with every knob on.
# Outputs land in gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
ALL="SCOT_CNX_OVERLAP=1 SCOT_ATTN_BWD_SPLIT=8 SCOT_CLN_FWD_HOIST=1 SCOT_DWCONV_SMEM=1 SCOT_ZERO_OVERLAP=1 SCOT_CPB_FAST=1 SCOT_CPB_BWD_SPLIT=8"
( time timeout 300 python scripts/ab_overlap.py B 64 30 ) > gpurun_out/ab_overlap.log 2>&1
echo "ab_overlap exit $?" >> gpurun_out/ab_overlap.log
( time timeout 150 python -m pytest tests/test_gpu_knobs.py -x -q ) > gpurun_out/knobs_test.log 2>&1
echo "knobs exit $?" >> gpurun_out/knobs_test.log
( time env $ALL timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/bench_allknobs.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_allknobs.log
( time env $ALL timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_configs.py -x -q -k "cln or forward_matches or other_baseline or smooth" ) > gpurun_out/parity_allknobs.log 2>&1
echo "parity exit $?" >> gpurun_out/parity_allknobs.log
( time timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_default.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_default.log
( time env $ALL timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_allknobs.csv python scripts/profile_step.py B 64 1 ) > gpurun_out/ncu_allknobs.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_allknobs.log
tail -n 12 gpurun_out/ab_overlap.log gpurun_out/knobs_test.log gpurun_out/bench_allknobs.log gpurun_out/parity_allknobs.log gpurun_out/smoke_default.log gpurun_out/ncu_allknobs.log
