#!/bin/bash
mkdir -p gpurun_out
NSR_NEUS_FWD=tc timeout 600 python -u -m pytest tests/test_gpu_neus.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
for v in tc scalar; do
NSR_NEUS_FWD=$v timeout 600 python bench.py --config C3 --steps 20 --warmup 5 > gpurun_out/bench_c3_$v.json 2> gpurun_out/bench_c3_$v.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_c3_$v.json').read().strip().splitlines()[-1])
    k = d.get('kernels_ms', {})
    print('C3 $v', round(d['ms_per_step'], 4), round(d['value'] / 1e6, 3), 'field fwd', k.get('nsr_neus_field_fwd'), 'bwd', k.get('nsr_neus_field_bwd'))
except Exception as e:
    print('C3 $v failed', e); print(open('gpurun_out/bench_c3_$v.err').read()[-1500:])
PY
done
