#!/bin/bash
# end of round 2: the driver's three commands on the final tree
mkdir -p gpurun_out
timeout 1500 python -u -m pytest tests -m gpu -x -q -rA -p no:cacheprovider > gpurun_out/r2_gputest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gputest_final.log
grep -E "passed|failed|^FAILED|^ERROR|pytest exit" gpurun_out/r2_gputest_final.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('C2', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['kernel'], round(d['roofline']['frac'], 3), 'traffic', d['roofline'].get('traffic'), 'whole', round(d['roofline']['whole_step']['frac'], 3))
print('clocks', d.get('clocks'))
for k, v in d.get('extra', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'error')})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline_secondary']['value'] if d.get('cpu_baseline_secondary') else None)
PY
tail -2 gpurun_out/bench_final.err
