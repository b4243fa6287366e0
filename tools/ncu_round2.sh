#!/bin/bash
# Profiling call for round 2 (one GPU, never multi-rank):  /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/ncu_round2.sh'
# Leaves in gpurun_out/: the launch list of the bench command and ncu --set full captures of the dominant C2 kernels, the NeuS field kernels
# (C3) and the kernels of the C4 step with the fused VanillaMLP kernels on.  Read them on the CPU box:
#   ncu -i gpurun_out/r2_c2.ncu-rep --page raw --csv | grep -E 'gpu__time_duration|dram__bytes_(read|write)\.sum|l1tex__data_pipe_lsu_wavefronts|sm__pipe_tensor|sm__warps_active|smsp__average_warp.*stall'
#   ncu -i gpurun_out/r2_c2.ncu-rep --page source --csv        # per-line stall samples (the library is built with -lineinfo)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nerf_bwd_kernel|nerf_rays_fwd_kernel|march_rays_mask_kernel|scan_counts_kernel|pack_kept_kernel|ray_bwd_loose_kernel" \
    -s 12 -c 8 -o gpurun_out/r2_c2 -f python tools/ncu_target.py 5 > gpurun_out/r2_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"neus_field_fwd_kernel|neus_field_bwd_kernel|radiance_bwd_kernel" \
    -s 4 -c 6 -o gpurun_out/r2_c3 -f python tools/neus_times.py > gpurun_out/r2_c3.log 2>&1
NSR_EXPERIMENTAL=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_c4_launches_fused_mlps.csv \
    python tools/neus_times.py > gpurun_out/r2_c4_launches_fused_mlps.log 2>&1
ls -la gpurun_out | tail -12
