#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 2" "8 4" "16 1" "16 2" "4 4" "4 8"; do
  set -- $cfg
  NSR_P2P_UNROLL=$1 NSR_P2P_CTAS_PER_SM=$2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py 2>/dev/null | grep -o '"rank": 0.*' | python -c "
import sys, json, re
s = sys.stdin.read()
m = re.search(r'\{.*?\}', '{' + s)
d = json.loads(m.group(0))
print('U=$1 ctas=$2', {k: d[k] for k in d if k.endswith('no_copy_ms') or k.endswith('max_err')})"
done
