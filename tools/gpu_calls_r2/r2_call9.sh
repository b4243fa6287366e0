#!/bin/bash
mkdir -p gpurun_out
NSR_TC_DEBUG=0 timeout 120 python tools/tc_bwd_bringup.py 8192 tiles,tc 2>&1 | tail -1 | cut -c1-700
for d in 4 10 2 1; do
  NSR_TC_DEBUG=$d timeout 120 python tools/tc_bwd_bringup.py 8192 tc 2>&1 | grep -o '"nsr_nerf_field_bwd_tc": [0-9.]*' | sed "s/^/dbg=$d /"
done
