#!/bin/bash
# round 2, second session, call 1: parity of everything touched (GEMM K < 128 / lm batching, two-call training step, balanced filter
# weight gradients, GemNet-OC warp-per-edge quadruplet kernel + row-blocked rbf product), then A/B timings.
set -u
OUT=gpurun_out/r2b_call1
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-900} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=1200 run python -m pytest tests -q -m gpu -rA -p no:cacheprovider
# training step: new default, then the two switches back to the old paths
TMO=300 run python bench_train.py --steps 10 --warmup 3
NB200_TRAIN_KEEP=0 TMO=300 run python bench_train.py --steps 10 --warmup 3
NB200_FWGRAD=old TMO=300 run python bench_train.py --steps 10 --warmup 3
TMO=300 run python bench_qhnet.py --steps 5 --warmup 2
NB200_GEMM_VARIANT=wide TMO=300 run python bench_qhnet.py --steps 3 --warmup 1
TMO=600 run python bench_gemnet.py --steps 3 --warmup 1
NB200_GOC_QUAD=functor TMO=600 run python bench_gemnet.py --steps 2 --warmup 1
grep -E "^\{|passed|failed|FAILED|Error|rc=|===" $OUT/log.txt | cut -c1-600 | tail -60
