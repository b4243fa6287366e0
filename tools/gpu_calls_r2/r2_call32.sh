#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 10 --no-extra > gpurun_out/bench_2gpu_last.json 2> gpurun_out/bench_2gpu_last.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_2gpu_last.json').read().strip().splitlines()[-1])
    print('2gpu', d['ms_per_step'], d['value'], d['config'].get('step','')[-120:])
except Exception as e:
    print('2gpu failed', e); print(open('gpurun_out/bench_2gpu_last.err').read()[-2500:])
PY
