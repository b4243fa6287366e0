#!/bin/bash
mkdir -p gpurun_out
timeout 180 python tools/tc_bwd_bringup.py 512 tiles,tc > gpurun_out/tc_bringup_512.log 2>&1; echo "exit $?" >> gpurun_out/tc_bringup_512.log; tail -6 gpurun_out/tc_bringup_512.log | cut -c1-1500
timeout 180 python tools/tc_bwd_bringup.py 8192 tiles,tiles_split,tc > gpurun_out/tc_bringup_8192.log 2>&1; echo "exit $?" >> gpurun_out/tc_bringup_8192.log; tail -6 gpurun_out/tc_bringup_8192.log | cut -c1-1500
