#!/bin/bash
# First GPU pass: parity tests group by group (a trap in one kernel must not hide the others), smoke, short bench.
mkdir -p gpurun_out
run() { # name, timeout, command...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/summary.txt
run build 600 python __graft_entry__.py
run warp 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "warp" -p no:cacheprovider
run voxel 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "voxel or preprocess" -p no:cacheprovider
run corr_fp32 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fp32 or odd or lookup_on_reference" -p no:cacheprovider
for f in 0 1 2 3 4; do CF_TC_FLAGS=$f run tf32_flags$f 300 python scripts/tf32_experiment.py; done
run corr_tf32 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tf32 or full_size_properties or trace" -p no:cacheprovider
run smoke 600 python -c "import __graft_entry__ as e; e.smoke()"
run bench 900 python bench.py --steps 50 --warmup 5
