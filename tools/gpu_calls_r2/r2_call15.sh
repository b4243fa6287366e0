#!/bin/bash
for d in 66 34 98 2; do
  NSR_TC_DEBUG=$d timeout 120 python tools/tc_bwd_bringup.py 8192 tc 2>&1 | grep -o '"nsr_nerf_field_bwd_tc": [0-9.]*' | sed "s/^/dbg=$d /"
done
