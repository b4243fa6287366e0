#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 50 --warmup 10 --no-extra > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_8gpu.json').read().strip().splitlines()[-1])
    print('8gpu', d['ms_per_step'], d['value'], d['config'].get('step','')[-160:])
except Exception as e:
    print('8gpu failed', e); print(open('gpurun_out/bench_8gpu.err').read()[-2500:])
PY
