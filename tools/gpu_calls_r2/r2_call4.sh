#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_nerf.py -x -q -rA -p no:cacheprovider > gpurun_out/r2_nerf_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_nerf_tests.log
grep -E "passed|failed|^FAILED|^ERROR|pytest exit" gpurun_out/r2_nerf_tests.log | tail -12
for k in tiles tiles_split; do
  NSR_BWD_KERNEL=$k timeout 300 python bench.py --steps 50 --warmup 10 --no-extra > gpurun_out/bench_$k.json 2> gpurun_out/bench_$k.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python - <<'PY'
import json
for name in ('tiles', 'tiles_split', 'full'):
    try:
        d = json.loads(open(f'gpurun_out/bench_{name}.json').read().strip().splitlines()[-1])
        print(name, d['ms_per_step'], d['value'], d.get('kernels_ms'))
        if 'extra' in d:
            for k, v in d['extra'].items():
                print(' ', k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'error')}, v.get('kernels_ms'))
            print('  cpu', d.get('cpu_baseline'), d.get('cpu_baseline_secondary'))
    except Exception as e:
        print(name, 'failed', e)
PY
tail -3 gpurun_out/bench_full.err
# in-graph per-kernel times of the two variants (cold cache, serialised: shares only)
for k in tiles tiles_split; do
  NSR_BWD_KERNEL=$k ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nerf_|march_|scan_|pack_|ray_bwd|loss" -s 150 -c 40 --csv \
    --log-file gpurun_out/r2_launches_$k.csv python tools/ncu_target.py 8 > gpurun_out/r2_launches_$k.log 2>&1
done
tail -22 gpurun_out/r2_launches_tiles_split.csv | cut -d, -f5,12-
