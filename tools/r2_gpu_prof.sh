#!/bin/bash
# role / phase timing of the fused node kernels + one ncu --set full capture of each
set -u
OUT=gpurun_out/r2_prof
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-400} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=400 run ncu --set full --clock-control none --import-source on -k regex:k_node -s 8 -c 4 -o $OUT/node_fused python bench.py --steps 4 --warmup 3 --streams 1 --skip-e2e --no-cpu-baseline
NB200_NVCC_EXTRA=-DNF_PROF timeout 300 python -m nabladft_b200.build --force >> $OUT/log.txt 2>&1
TMO=200 run python tools/nf_prof.py
grep -v "^{" $OUT/log.txt | tail -60
