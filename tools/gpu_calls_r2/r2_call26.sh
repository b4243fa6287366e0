#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 50 --warmup 10 --no-extra > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1])
    k = d.get('kernels_ms')
    print('$name', round(d['ms_per_step'], 4), round(d['value'] / 1e6, 2), 'fwd', k.get('nsr_nerf_rays_fwd'), 'scatter', k.get('nsr_nerf_table_scatter'))
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/bench_$name.err').read()[-1500:])
PY
}
run base A=1
run batch8 NSR_FWD_BATCH=8
run sc6 NSR_SCATTER_CTAS=6
run sc8 NSR_SCATTER_CTAS=8
run base2 A=1
