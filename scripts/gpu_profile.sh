#!/bin/bash
# plain run must exit 0 before ncu; then the launch list and one --set full capture of the hot kernels
mkdir -p gpurun_out
python scripts/profile_step.py 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 3 > gpurun_out/ncu_list.log 2>&1
python scripts/profile_step.py 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"corr_lookup|warp_frame|corr_tc|voxel_scatter|voxel_stats|voxel_normalise" -s 20 -c 12 -o gpurun_out/prof_r1a python scripts/profile_step.py 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/plain.log gpurun_out/ncu_list.log gpurun_out/ncu_full.log
ls -la gpurun_out/
