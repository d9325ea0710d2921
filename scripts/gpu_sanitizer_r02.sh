#!/bin/bash
# compute-sanitizer memcheck + racecheck of tests/sanitizer_driver.py (every kernel family of the engine at small sizes)
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tests/sanitizer_driver.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
timeout 700 compute-sanitizer --tool racecheck --print-limit 20 python tests/sanitizer_driver.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -3 gpurun_out/r02_sanitizer_memcheck.log gpurun_out/r02_sanitizer_racecheck.log
