#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export SCOT_CNX_OVERLAP=0 SCOT_ATTN_BWD_SPLIT=0
( cd scripts
  timeout 40 python noise_trace.py tiny 10
  SCOT_PDL=0 timeout 40 python noise_trace.py tiny 10
  SCOT_WGRAD_OVERLAP=0 timeout 40 python noise_trace.py tiny 10
  SCOT_PDL=0 SCOT_WGRAD_OVERLAP=0 timeout 40 python noise_trace.py tiny 10
) > gpurun_out/noise_trace.log 2>&1
tail -n 150 gpurun_out/noise_trace.log
