#!/bin/bash
mkdir -p gpurun_out
NSR_TC_DEBUG=0 timeout 120 python tools/tc_bwd_bringup.py 8192 tiles,tc 2>&1 | tail -1 | cut -c1-420
for d in 4 10; do
  NSR_TC_DEBUG=$d timeout 120 python tools/tc_bwd_bringup.py 8192 tc 2>&1 | grep -o '"nsr_nerf_field_bwd_tc": [0-9.]*' | sed "s/^/dbg=$d /"
done
for d in 4 0; do
  NSR_TC_DEBUG=$d NSR_TC_TRACE=gpurun_out/tc_trace_$d.txt timeout 120 python tools/tc_bwd_bringup.py 8192 tc > /dev/null 2>&1
  echo "trace dbg=$d"; python tools/tc_trace.py gpurun_out/tc_trace_$d.txt | grep -A8 "steps traced"
done
python tools/tc_trace.py gpurun_out/tc_trace_0.txt | head -10
