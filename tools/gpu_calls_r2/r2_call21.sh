#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_nerf.py tests/test_gpu_z_training.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
for a in 1 0; do
NSR_MARCH_ALLOC=$a timeout 300 python bench.py --steps 50 --warmup 10 --no-extra > gpurun_out/bench_alloc$a.json 2> gpurun_out/bench_alloc$a.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_alloc$a.json').read().strip().splitlines()[-1])
    print('alloc=$a', d['ms_per_step'], d['value'], d.get('kernels_ms'))
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_alloc$a.err').read()[-2000:])
PY
done
