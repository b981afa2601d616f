#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
for f in 0 1 2 3 4; do CF_TC_FLAGS=$f run tf32_flags$f 300 python scripts/tf32_experiment.py; done
run corr_all 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "corr or lookup or trace" -p no:cacheprovider
run smoke 600 python -c "import __graft_entry__ as e; e.smoke()"
TAILN=3 run bench 900 python bench.py --steps 50 --warmup 5
