#!/bin/bash
# call 5: weight-gradient leaves on a side stream (A/B), full GPU suite
set -u
OUT=gpurun_out/r2b_call5
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-900} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=900 run python -m pytest tests/test_gpu_train.py -q -m gpu -rA -p no:cacheprovider
TMO=300 run python bench_train.py --steps 10 --warmup 3
NB200_TRAIN_SIDE=0 TMO=300 run python bench_train.py --steps 10 --warmup 3
TMO=300 run python bench_train.py --steps 10 --warmup 3 --storage bf16
TMO=300 run python bench_train.py --steps 10 --warmup 3 --loss e
grep -E "^\{|passed|failed|FAILED|Error|rc=|===|bf16 edge|kept vs" $OUT/log.txt | cut -c1-330 | tail -30
