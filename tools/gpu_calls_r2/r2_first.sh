#!/bin/bash
# round 2, first GPU call: the whole GPU suite with every gate open (no -x: see every failure), then timings
mkdir -p gpurun_out
export NSR_EXPERIMENTAL=1
timeout 1200 python -u -m pytest tests -m gpu -q -rA --timeout 180 -p no:cacheprovider > gpurun_out/r2_gputest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_gputest_all.log
grep -E "passed|failed|^FAILED|^ERROR|pytest exit" gpurun_out/r2_gputest_all.log | tail -40
NSR_EXPERIMENTAL=0 timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
NSR_EXPERIMENTAL=pack_scan timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_pack_scan.json 2> gpurun_out/bench_pack_scan.err
NSR_FWD_CTAS=3 NSR_EXPERIMENTAL=0 timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_fwd_ctas3.json 2> gpurun_out/bench_fwd_ctas3.err
python - <<'PY'
import json
for name in ('default', 'pack_scan', 'fwd_ctas3'):
    try:
        d = json.loads(open(f'gpurun_out/bench_{name}.json').read().strip().splitlines()[-1])
        print(name, d['ms_per_step'], d['value'], d.get('kernels_ms'))
    except Exception as e:
        print(name, 'failed', e)
PY
NSR_EXPERIMENTAL=0 timeout 300 python tools/neus_times.py > gpurun_out/neus_times_torch_mlps.json 2> gpurun_out/neus_times_torch_mlps.err
NSR_EXPERIMENTAL=1 timeout 300 python tools/neus_times.py > gpurun_out/neus_times_fused_mlps.json 2> gpurun_out/neus_times_fused_mlps.err
tail -c 1500 gpurun_out/neus_times_torch_mlps.json; echo; tail -c 1500 gpurun_out/neus_times_fused_mlps.json
