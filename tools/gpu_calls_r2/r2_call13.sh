#!/bin/bash
mkdir -p gpurun_out
for d in 4 0; do
  NSR_TC_DEBUG=$d NSR_TC_TRACE=gpurun_out/tc_trace_$d.txt timeout 120 python tools/tc_bwd_bringup.py 8192 tc > /dev/null 2>&1
  echo "trace dbg=$d"; python tools/tc_trace.py gpurun_out/tc_trace_$d.txt | grep -A8 "steps traced"
done
