#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -u -m pytest tests -m gpu -x -q -rA -p no:cacheprovider > gpurun_out/r2_gputest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gputest_final.log
grep -E "passed|failed|^FAILED|^ERROR|pytest exit" gpurun_out/r2_gputest_final.log | tail -6
