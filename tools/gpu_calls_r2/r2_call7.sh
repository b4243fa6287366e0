#!/bin/bash
# fixed tests re-run, forward CTA A/B, ncu --set full of the backward variants and the forward
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_gpu_z_vanilla.py tests/test_gpu_neus.py -m gpu -q -rA -p no:cacheprovider -k "vanilla or full_size" > gpurun_out/r2_gputest_fix.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gputest_fix.log
grep -E "passed|failed|^FAILED|^ERROR|pytest exit" gpurun_out/r2_gputest_fix.log | tail -12
NSR_FWD_CTAS=3 NSR_BWD_KERNEL=tiles_split timeout 300 python bench.py --steps 50 --warmup 10 --no-extra > gpurun_out/bench_fwd3.json 2> gpurun_out/bench_fwd3.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_fwd3.json').read().strip().splitlines()[-1])
print('fwd3+split', d['ms_per_step'], d['value'], d.get('kernels_ms'))
PY
NSR_BWD_KERNEL=tiles_split timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nerf_table_scatter_kernel|nerf_bwd_kernel|nerf_rays_fwd_kernel" \
    -s 6 -c 3 -o gpurun_out/r2_split -f python tools/ncu_target.py 4 > gpurun_out/r2_split.log 2>&1
NSR_BWD_KERNEL=tc timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nerf_bwd_tc_kernel" \
    -s 2 -c 1 -o gpurun_out/r2_tc -f python tools/ncu_target.py 4 > gpurun_out/r2_tc.log 2>&1
tail -3 gpurun_out/r2_split.log gpurun_out/r2_tc.log
ls -la gpurun_out | tail -8
