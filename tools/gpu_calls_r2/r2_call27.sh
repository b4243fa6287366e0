#!/bin/bash
# end-of-round profiles: launch list of the bench command, --set full of the C2 step's kernels and of the NeuS field kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 6 --warmup 3 --no-extra > gpurun_out/r2_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nerf_rays_fwd_kernel|nerf_bwd_kernel|nerf_table_scatter_kernel|march_rays_mask_kernel|pack_kept_kernel|ray_bwd_loose_kernel" \
    -s 18 -c 6 -o gpurun_out/r2_final -f python tools/ncu_target.py 5 > gpurun_out/r2_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"neus_field_fwd_kernel|neus_field_bwd_kernel" \
    -s 4 -c 2 -o gpurun_out/r2_c3 -f python tools/neus_times.py > gpurun_out/r2_c3.log 2>&1
tail -2 gpurun_out/r2_final.log gpurun_out/r2_c3.log
ls -la gpurun_out | grep -E "ncu-rep|launches_final"
