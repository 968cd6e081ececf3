#!/bin/bash
# GPU call: first device run of the fused node kernels (painn_fused.cu): parity, then A/B bench with the per-category breakdown.
set -u
OUT=gpurun_out/r2_fused
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-400} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=300 run python -m pytest tests/test_gpu_painn.py -q -x -rA -p no:cacheprovider -k "fused or golden or cfg2_slice or spk_painn or edge_cases"
TMO=200 run python bench.py --steps 30 --warmup 5 --no-cpu-baseline --streams 1
TMO=200 run python bench.py --steps 30 --warmup 5 --no-cpu-baseline --streams 1 --node unfused
TMO=200 run python bench.py --steps 30 --warmup 5 --no-cpu-baseline
tail -30 $OUT/log.txt
