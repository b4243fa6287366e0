#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"neus_field_fwd_tc_kernel" -s 2 -c 1 -o gpurun_out/r2_c3_tc -f python tools/neus_times.py > gpurun_out/r2_c3_tc.log 2>&1
ls -la gpurun_out/r2_c3_tc.ncu-rep
