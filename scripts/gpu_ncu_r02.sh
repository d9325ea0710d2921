#!/bin/bash
# ncu evidence of one Poseidon-B step (batch 64, eager launches, second step = warm attributes):
#  (1) every kernel of the step with a short metric list (duration, DRAM bytes, DRAM / tensor-pipe / issue utilisation):
#      one row per kernel name in r02_ncu_summary.md, DRAM bytes of the largest (stage-0) launch in r02_kernel_traffic.json
#  (2) --set full + source of the tcgen05 attention kernels (one launch each): stall-reason / hot-line summary
# Reports are summarised on the box (ncu reads its own reports without a GPU); only the attention report is kept.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
M=$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
M=$M,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 500 ncu --metrics $M --clock-control none -s 1137 -c 1400 -f -o gpurun_out/r02_ncu_step \
  python scripts/profile_step.py B 64 2 opt > gpurun_out/ncu_step.log 2>&1
echo "step metrics exit $?"
for k in attn_tc_bwd attn_tc_fwd; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$k -s 24 -c 1 -f -o gpurun_out/r02_ncu_$k \
    python scripts/profile_step.py B 64 2 > gpurun_out/ncu_$k.log 2>&1
  echo "$k exit $?"
  python scripts/ncu_stalls.py gpurun_out/r02_ncu_$k.ncu-rep 25 > gpurun_out/r02_ncu_${k}_stalls.txt 2>&1
done
python scripts/ncu_table.py gpurun_out/r02_ncu_summary.md gpurun_out/r02_kernel_traffic.json gpurun_out/r02_ncu_step.ncu-rep \
  gpurun_out/r02_ncu_attn_tc_bwd.ncu-rep gpurun_out/r02_ncu_attn_tc_fwd.ncu-rep > /dev/null
ls -la gpurun_out/r02_ncu_*.ncu-rep
rm -f gpurun_out/r02_ncu_step.ncu-rep gpurun_out/r02_ncu_attn_tc_fwd.ncu-rep
