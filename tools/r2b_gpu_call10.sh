#!/bin/bash
# call 10: half-operand hand-over with warp-uniform halves: parity + A/B
set -u
OUT=gpurun_out/r2b_call10
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-600} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=600 run python -m pytest tests/test_gpu_painn.py -q -m gpu -x -p no:cacheprovider
TMO=300 run python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline
NB200_NF_XSPLIT=0 TMO=300 run python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline
TMO=400 run python bench_gemnet.py --steps 3 --warmup 1
NB200_GEMM_XSPLIT=0 TMO=400 run python bench_gemnet.py --steps 2 --warmup 1
grep -E "^\{|passed|failed|FAILED|Error|rc=|===" $OUT/log.txt | cut -c1-300 | tail -30
