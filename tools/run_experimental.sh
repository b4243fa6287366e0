#!/bin/bash
# Runs the GPU tests of the kernels that are still behind NSR_EXPERIMENTAL (see nsr_b200/config.py) and leaves the log in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/run_experimental.sh'
mkdir -p gpurun_out
export NSR_EXPERIMENTAL=1
python -u -m pytest tests/test_gpu_frontend.py tests/test_gpu_vanilla.py tests/test_gpu_export.py tests/test_gpu_c1_vanilla.py -q -rA --timeout 120 -p no:cacheprovider \
  > gpurun_out/exp_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/exp_tests.log
tail -40 gpurun_out/exp_tests.log
