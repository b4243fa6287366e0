#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_nerf.py tests/test_gpu_z_training.py tests/test_gpu_shade_optim.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
for z in 1 0; do
NSR_PREZERO_GRADS=$z timeout 300 python bench.py --steps 50 --warmup 10 --no-extra > gpurun_out/bench_pz$z.json 2> gpurun_out/bench_pz$z.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_pz$z.json').read().strip().splitlines()[-1])
    print('prezero=$z', d['ms_per_step'], d['value'], d['e2e']['value'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_pz$z.err').read()[-2000:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 10 --no-extra > gpurun_out/bench_2gpu_pz.json 2> gpurun_out/bench_2gpu_pz.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_2gpu_pz.json').read().strip().splitlines()[-1])
    print('2gpu', d['ms_per_step'], d['value'])
except Exception as e:
    print('2gpu failed', e); print(open('gpurun_out/bench_2gpu_pz.err').read()[-2500:])
PY
