#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py 2>gpurun_out/p2p_check.err | grep '"rank": 0' | cut -c1-900
tail -3 gpurun_out/p2p_check.err
for ov in 1 0; do
NSR_P2P_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 10 --no-extra > gpurun_out/bench_2gpu_ov$ov.json 2> gpurun_out/bench_2gpu_ov$ov.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_2gpu_ov$ov.json').read().strip().splitlines()[-1])
    print('2gpu overlap=$ov', d['ms_per_step'], d['value'], d['config'].get('step','')[-150:])
except Exception as e:
    print('2gpu overlap=$ov failed', e); print(open('gpurun_out/bench_2gpu_ov$ov.err').read()[-2500:])
PY
done
