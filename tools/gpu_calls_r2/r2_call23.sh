#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python bench.py --steps 50 --warmup 10 --no-extra > gpurun_out/bench_mod2.json 2> gpurun_out/bench_mod2.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_mod2.json').read().strip().splitlines()[-1])
print('C2', d['ms_per_step'], d['value'], d.get('kernels_ms'))
PY
done
