#!/bin/bash
# re-entry call: tc bring-up (own processes: a trap kills the context), whole GPU suite minus the tc cases, the tc cases, backward A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 180 python tools/tc_bwd_bringup.py 512 tiles,tiles_split,tc > gpurun_out/tc_bringup_512.log 2>&1; echo "exit $?" >> gpurun_out/tc_bringup_512.log; tail -5 gpurun_out/tc_bringup_512.log | cut -c1-1200
timeout 180 python tools/tc_bwd_bringup.py 8192 tiles,tiles_split,tc > gpurun_out/tc_bringup_8192.log 2>&1; echo "exit $?" >> gpurun_out/tc_bringup_8192.log; tail -5 gpurun_out/tc_bringup_8192.log | cut -c1-1200
timeout 1500 python -u -m pytest tests -m gpu -q -rA -p no:cacheprovider -k "not per_ray_tc" > gpurun_out/r2_gputest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gputest.log
grep -E "passed|failed|^FAILED|^ERROR|pytest exit" gpurun_out/r2_gputest.log | tail -30
timeout 600 python -u -m pytest tests/test_gpu_nerf.py -m gpu -q -rA -p no:cacheprovider -k "per_ray_tc" > gpurun_out/r2_gputest_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gputest_tc.log
grep -E "passed|failed|^FAILED|^ERROR|pytest exit|Error|assert " gpurun_out/r2_gputest_tc.log | tail -20
for k in tiles tiles_split tc; do
  NSR_BWD_KERNEL=$k timeout 300 python bench.py --steps 50 --warmup 10 --no-extra > gpurun_out/bench_$k.json 2> gpurun_out/bench_$k.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_$k.json').read().strip().splitlines()[-1])
    print('$k', d['ms_per_step'], d['value'], d.get('kernels_ms'))
except Exception as e:
    print('$k', 'failed', e)
PY
done
