#!/bin/bash
# Final round-1 validation call: knob noise report, the whole GPU test suite, the bench line, smoke(), an ncu launch list.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( time timeout 90 python scripts/knob_noise.py tiny ) > gpurun_out/knob_noise_tiny.log 2>&1
echo "noise exit $?" >> gpurun_out/knob_noise_tiny.log
( time timeout 240 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log
( time timeout 120 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_final.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_final.log
( time timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_final.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_final.log
( time timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_final.csv python scripts/profile_step.py B 64 1 ) > gpurun_out/ncu_final.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_final.log
tail -n 15 gpurun_out/knob_noise_tiny.log gpurun_out/pytest_gpu_final.log gpurun_out/bench_final.log gpurun_out/smoke_final.log gpurun_out/ncu_final.log
