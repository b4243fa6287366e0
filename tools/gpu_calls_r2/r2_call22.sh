#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_nerf.py tests/test_gpu_ops.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --warmup 10 --no-extra > gpurun_out/bench_mod.json 2> gpurun_out/bench_mod.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_mod.json').read().strip().splitlines()[-1])
print('C2', d['ms_per_step'], d['value'], d.get('kernels_ms'))
PY
