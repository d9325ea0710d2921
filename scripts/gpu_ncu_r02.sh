#!/bin/bash
# ncu --set full captures of the kernels of one Poseidon-B step (one launch each of the kinds the bench / DESIGN name).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cap() {  # name regex skip count
  timeout 240 ncu --set full --clock-control none $5 -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/r02_ncu_$1 \
    python scripts/profile_step.py B 64 1 opt > gpurun_out/ncu_$1.log 2>&1
  echo "$1 exit $?"
}
cap attn_tc 'attn_tc_fwd|attn_tc_bwd' 0 4 '--import-source on'
cap cln 'cln_fwd|cln_bwd' 2 6
cap gemm_async 'gemm_async_epi' 4 6
cap gemm_tc 'gemm_tc_kernel' 6 8
cap conv 'dwconv7|conv5|scale_add|im2col|merge_|unshuffle|shuffle_grad|loss_' 0 16
cap optim 'adamw|grad_sq_norm|cast_f32' 0 4
cap attn_small 'attn_fwd_kernel|attn_bwd_dq|attn_bwd_dkv|cpb_' 0 8
# summarise on the box (ncu reads its own reports without a GPU), keep only the attention capture (source-level stalls)
python scripts/ncu_table.py gpurun_out/r02_ncu_summary.md gpurun_out/r02_kernel_traffic.json gpurun_out/r02_ncu_*.ncu-rep > /dev/null
python scripts/ncu_stalls.py gpurun_out/r02_ncu_attn_tc.ncu-rep 25 > gpurun_out/r02_ncu_attn_tc_stalls.txt 2>&1
ls -la gpurun_out/r02_ncu_*.ncu-rep
for f in gpurun_out/r02_ncu_*.ncu-rep; do case $f in *attn_tc*) ;; *) rm -f $f;; esac; done
