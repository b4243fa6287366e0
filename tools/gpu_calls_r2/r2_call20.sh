#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print('C2', d['ms_per_step'], d['value'], d.get('kernels_ms'))
for k, v in d.get('extra', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'error')}, v.get('kernels_ms'))
print('cpu', d.get('cpu_baseline'), d.get('cpu_baseline_secondary'))
PY
tail -3 gpurun_out/bench_full.err
