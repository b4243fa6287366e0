#!/bin/bash
mkdir -p gpurun_out
for d in 4 0; do
  NSR_TC_DEBUG=$d NSR_TC_TRACE=gpurun_out/tc_trace_$d.txt timeout 120 python tools/tc_bwd_bringup.py 8192 tc 2>&1 | grep -o '"nsr_nerf_field_bwd_tc": [0-9.]*' | sed "s/^/dbg=$d /"
  python tools/tc_trace.py gpurun_out/tc_trace_$d.txt | tail -32
done
