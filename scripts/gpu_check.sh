#!/bin/bash
# usage: gpu_check.sh [tests] [timeline] [bench] [ncu] [scale]
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc"; tail -n ${TAILN:-15} gpurun_out/$name.log; return $rc; }
for stage in "$@"; do
  case $stage in
    tests) TAILN=25 run tests 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ;;
    smoke) run smoke 600 python -c "import __graft_entry__ as e; e.smoke()" ;;
    timeline) TAILN=40 run timeline 300 python scripts/tc_timeline.py ;;
    bench) TAILN=2 run bench 900 python bench.py --steps 200 --warmup 10 ;;
    benchref) TAILN=2 run benchref 900 python bench.py --impl reference --steps 5 --warmup 1 ;;
    ncu)
      python scripts/profile_step.py 1 > gpurun_out/plain.log 2>&1 &&
      ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 1 > gpurun_out/ncu_list.log 2>&1
      python scripts/profile_step.py 1 > gpurun_out/plain2.log 2>&1 &&
      ncu --set full --clock-control none --import-source on -k regex:"corr_|warp_|voxel_|avg_pool" -c 22 -o gpurun_out/prof_step python scripts/profile_step.py 1 > gpurun_out/ncu_full.log 2>&1
      tail -n 3 gpurun_out/ncu_full.log ;;
  esac
done
