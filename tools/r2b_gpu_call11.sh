#!/bin/bash
# call 11: one filter row per undirected pair in the TRAINING path (parity + A/B); first GemNet-OC training-step number
set -u
OUT=gpurun_out/r2b_call11
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-600} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=600 run python -m pytest tests/test_gpu_train.py tests/test_gpu_painn.py -q -m gpu -p no:cacheprovider
TMO=300 run python bench_train.py --steps 10 --warmup 3 --storage bf16
TMO=300 run python bench_train.py --steps 10 --warmup 3
NB200_TRAIN_HALF_ROWS=0 TMO=300 run python bench_train.py --steps 10 --warmup 3
TMO=400 run python bench_gemnet.py --train --batch 16 --steps 5 --warmup 2
grep -E "^\{|passed|failed|FAILED|Error|rc=|===" $OUT/log.txt | cut -c1-500 | tail -30
