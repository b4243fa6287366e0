#!/bin/bash
# 4 GPUs: the NVSwitch multicast form of the exchange (ranges / channels, pipelined with the scatter) against NCCL, then the bench
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py 2>gpurun_out/p2p_check4.err | grep -o '"rank": 0.*' | python -c "
import sys, json, re
s = sys.stdin.read()
m = re.search(r'\{.*?\}', '{' + s)
print(m.group(0)[:900] if m else s[:900])"
for ov in 1 0; do
NSR_P2P_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 50 --warmup 10 --no-extra > gpurun_out/bench_4gpu_ov$ov.json 2> gpurun_out/bench_4gpu_ov$ov.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_4gpu_ov$ov.json').read().strip().splitlines()[-1])
    print('4gpu overlap=$ov', d['ms_per_step'], d['value'], d['config'].get('step','')[-140:])
except Exception as e:
    print('4gpu overlap=$ov failed', e); print(open('gpurun_out/bench_4gpu_ov$ov.err').read()[-2500:])
PY
done
