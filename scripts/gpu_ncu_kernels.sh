#!/bin/bash
# usage: gpu_ncu_kernels.sh <regex> <count> <outname>
mkdir -p gpurun_out
python scripts/profile_step.py 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$1" -c $2 -o gpurun_out/$3 python scripts/profile_step.py 1 > gpurun_out/ncu_$3.log 2>&1
tail -n 3 gpurun_out/ncu_$3.log
