#!/bin/bash
# whole GPU suite (the driver's command) after the test fixes + tc kernel A/B switches
mkdir -p gpurun_out
timeout 1500 python -u -m pytest tests -m gpu -x -q -rA -p no:cacheprovider > gpurun_out/r2_gputest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gputest.log
grep -E "passed|failed|^FAILED|^ERROR|pytest exit" gpurun_out/r2_gputest.log | tail -8
for d in 0 1 2 3 4; do
  NSR_TC_DEBUG=$d timeout 120 python tools/tc_bwd_bringup.py 8192 tc 2>&1 | grep -o '"nsr_nerf_field_bwd_tc": [0-9.]*' | sed "s/^/dbg=$d /"
done
