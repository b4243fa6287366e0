#!/bin/bash
# whole GPU suite exactly as the driver runs it (plus -rA into a log), then the default bench
mkdir -p gpurun_out
timeout 1700 python -u -m pytest tests -m gpu -x -q -rA -p no:cacheprovider > gpurun_out/r2_gputest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gputest.log
grep -E "passed|failed|^FAILED|^ERROR|pytest exit|input-gradient" gpurun_out/r2_gputest.log | tail -30
grep -E "slowest|s call" gpurun_out/r2_gputest.log | head
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_suite.json 2> gpurun_out/bench_suite.err; tail -c 600 gpurun_out/bench_suite.json
