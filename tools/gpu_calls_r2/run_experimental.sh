#!/bin/bash
# First GPU call of the next round: run the GPU tests that are still behind NSR_EXPERIMENTAL (see nsr_b200/config.py) and time C3 / C4 with
# the torch VanillaMLP layers and with the fused VanillaMLP kernels.  Logs land in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/run_experimental.sh'
mkdir -p gpurun_out
export NSR_EXPERIMENTAL=1
python -u -m pytest tests/test_gpu_z_frontend.py tests/test_gpu_z_vanilla.py tests/test_gpu_z_export.py tests/test_gpu_zy_configs.py tests/test_gpu_z_training.py tests/test_gpu_zz_export_colours.py -q -rA --timeout 120 \
  -p no:cacheprovider > gpurun_out/exp_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/exp_tests.log
tail -40 gpurun_out/exp_tests.log
NSR_EXPERIMENTAL=0 timeout 300 python tools/neus_times.py > gpurun_out/neus_times_torch_mlps.json 2> gpurun_out/neus_times_torch_mlps.err
NSR_EXPERIMENTAL=1 timeout 300 python tools/neus_times.py > gpurun_out/neus_times_fused_mlps.json 2> gpurun_out/neus_times_fused_mlps.err
tail -c 1500 gpurun_out/neus_times_torch_mlps.json; echo; tail -c 1500 gpurun_out/neus_times_fused_mlps.json
# end-to-end training on the analytic scene (PSNR before / after, steps/s): NeRF then NeuS
NSR_EXPERIMENTAL=0 timeout 600 python tools/train_synthetic.py --model nerf --steps 2000 --export > gpurun_out/train_nerf.json 2> gpurun_out/train_nerf.err
NSR_EXPERIMENTAL=0 timeout 600 python tools/train_synthetic.py --model neus --steps 2000 --export > gpurun_out/train_neus.json 2> gpurun_out/train_neus.err
tail -c 800 gpurun_out/train_nerf.json; tail -c 600 gpurun_out/train_nerf.err; tail -c 800 gpurun_out/train_neus.json; tail -c 600 gpurun_out/train_neus.err
# A/B of the opt-in step variants on the bench workload (C2, one GPU)
NSR_EXPERIMENTAL=0 timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
NSR_EXPERIMENTAL=pack_scan timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_pack_scan.json 2> gpurun_out/bench_pack_scan.err
python - <<'PY'
import json
for name in ('default', 'pack_scan'):
    try:
        d = json.loads(open(f'gpurun_out/bench_{name}.json').read().strip().splitlines()[-1])
        print(name, d['ms_per_step'], d['value'], d.get('kernels_ms'))
    except Exception as e:
        print(name, 'failed', e)
PY
# memcheck of the kernels added at the end of round 1 (small test cases only; compute-sanitizer slows kernels down 10-100x)
NSR_EXPERIMENTAL=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_z_frontend.py \
  "tests/test_gpu_z_export.py::test_marching_cubes_matches_oracle_exactly" "tests/test_gpu_z_vanilla.py::test_vanilla_mlp_matches_oracle_forward_and_backward" \
  "tests/test_gpu_z_vanilla.py::test_vanilla_radiance_matches_oracle_forward_and_backward" -q -x -p no:cacheprovider > gpurun_out/memcheck_new_kernels.log 2>&1
echo "memcheck exit $?" >> gpurun_out/memcheck_new_kernels.log
tail -15 gpurun_out/memcheck_new_kernels.log
# forward kernel sized for three CTAs per SM (80 registers, some spills, 24 warps/SM): parity of the NeRF model tests, then the bench A/B
NSR_FWD_CTAS=3 timeout 600 python -m pytest tests/test_gpu_nerf.py -q -x -p no:cacheprovider > gpurun_out/fwd_ctas3_tests.log 2>&1; echo "exit $?" >> gpurun_out/fwd_ctas3_tests.log
NSR_FWD_CTAS=3 NSR_EXPERIMENTAL=0 timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_fwd_ctas3.json 2> gpurun_out/bench_fwd_ctas3.err
tail -3 gpurun_out/fwd_ctas3_tests.log; python -c "
import json; d=json.loads(open('gpurun_out/bench_fwd_ctas3.json').read().strip().splitlines()[-1]); print('fwd_ctas3', d['ms_per_step'], d['value'], d.get('kernels_ms'))"
