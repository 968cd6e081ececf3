#!/bin/bash
set -u
OUT=gpurun_out/r2_fused2
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-400} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=300 run python -m pytest tests/test_gpu_painn.py -q -x -rA -p no:cacheprovider -k "fused or golden or cfg2_slice or edge_cases"
TMO=200 run python bench.py --steps 30 --warmup 5 --no-cpu-baseline --streams 1
TMO=200 run python bench.py --steps 30 --warmup 5 --no-cpu-baseline
# role timing (separate build with -DNF_PROF; the box copy is discarded afterwards)
NB200_NVCC_EXTRA=-DNF_PROF timeout 300 python -m nabladft_b200.build --force >> $OUT/log.txt 2>&1
TMO=200 run python tools/nf_prof.py
grep -v "^{" $OUT/log.txt | tail -40
