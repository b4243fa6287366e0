#!/bin/bash
# 2 GPUs: nerf tests on the new default path, 1-GPU bench, 2-GPU bench (new one-launch exchange, zero-copy) vs legacy exchange
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_gpu_nerf.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --warmup 10 --no-extra > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
print('1gpu', d['ms_per_step'], d['value'], d.get('kernels_ms'), d['roofline']['kernel'], round(d['roofline']['frac'],3))
PY
for mode in fused legacy; do
NSR_P2P_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 10 --no-extra > gpurun_out/bench_2gpu_$mode.json 2> gpurun_out/bench_2gpu_$mode.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_2gpu_$mode.json').read().strip().splitlines()[-1])
    print('2gpu $mode', d['ms_per_step'], d['value'], d['config'].get('step','')[-120:])
except Exception as e:
    print('2gpu $mode failed', e); print(open('gpurun_out/bench_2gpu_$mode.err').read()[-1500:])
PY
done
