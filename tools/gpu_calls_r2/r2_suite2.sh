#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -u -m pytest tests -m gpu -x -q -rA -p no:cacheprovider > gpurun_out/r2_gputest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gputest.log
grep -E "passed|failed|^FAILED|^ERROR|pytest exit|input-gradient" gpurun_out/r2_gputest.log | tail -30
timeout 600 python tools/gather_bench.py > gpurun_out/gather_bench.log 2>&1; tail -5 gpurun_out/gather_bench.log
