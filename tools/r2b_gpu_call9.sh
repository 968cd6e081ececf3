#!/bin/bash
# call 9: activation operand handed over in two K halves (tc_pipe.cuh xsplit): parity, then A/B on the headline bench, GemNet-OC, QHNet
set -u
OUT=gpurun_out/r2b_call9
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-600} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=600 run python -m pytest tests/test_gpu_painn.py -q -m gpu -x -p no:cacheprovider
TMO=300 run python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline
NB200_NF_XSPLIT=0 TMO=300 run python bench.py --steps 30 --warmup 5 --no-train --no-cpu-baseline
TMO=400 run python bench_gemnet.py --steps 3 --warmup 1
NB200_GEMM_XSPLIT=0 TMO=400 run python bench_gemnet.py --steps 2 --warmup 1
TMO=300 run python bench_qhnet.py --steps 5 --warmup 2
TMO=600 run python -m pytest tests/test_zz_gpu_first_runs.py tests/test_gpu_qhnet.py -q -m gpu -p no:cacheprovider
grep -E "^\{|passed|failed|FAILED|Error|rc=|===" $OUT/log.txt | cut -c1-400 | tail -30
