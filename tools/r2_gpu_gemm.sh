#!/bin/bash
set -u
OUT=gpurun_out/r2_gemm
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-600} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=300 run python -m pytest tests/test_gpu_painn.py -q -x -rA -p no:cacheprovider -k "gemm_tf32x3_matches or backends_agree"
TMO=300 run python tools/gemm_microbench.py
NB200_GEMM_VARIANT=wide TMO=300 run python tools/gemm_microbench.py
TMO=300 run python bench_qhnet.py
TMO=300 run python bench_gemnet.py --batch 512 --steps 3 --warmup 2
TMO=300 run python -m pytest tests/test_gpu_qhnet.py tests/test_zz_gpu_first_runs.py -q -x -p no:cacheprovider
grep -E "passed|failed|rc=|^\{'M'|^\{\"metric|rel err" $OUT/log.txt | cut -c1-330
