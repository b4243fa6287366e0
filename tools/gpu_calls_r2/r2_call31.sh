#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_gpu_neus.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
timeout 600 python bench.py --config C3 --steps 20 --warmup 5 > gpurun_out/bench_c3_b.json 2> gpurun_out/bench_c3_b.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_c3_b.json').read().strip().splitlines()[-1])
k = d.get('kernels_ms', {})
print('C3 batched', round(d['ms_per_step'], 4), round(d['value'] / 1e6, 3), 'field fwd', k.get('nsr_neus_field_fwd'), 'bwd', k.get('nsr_neus_field_bwd'))
PY
