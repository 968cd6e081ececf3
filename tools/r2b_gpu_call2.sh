#!/bin/bash
# call 2: stacked [primal ; tangent] Linear backward + bf16 edge storage: parity, timings; launch lists of the three secondary workloads
set -u
OUT=gpurun_out/r2b_call2
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-900} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=900 run python -m pytest tests/test_gpu_train.py tests/test_gpu_painn.py -q -m gpu -rA -p no:cacheprovider
TMO=300 run python bench_train.py --steps 10 --warmup 3
TMO=300 run python bench_train.py --steps 10 --warmup 3 --storage bf16
TMO=300 run python bench_train.py --steps 10 --warmup 3 --loss e
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/train_bf16_launches.csv python bench_train.py --steps 1 --warmup 1 --storage bf16 > $OUT/ncu_train.log 2>&1
echo "ncu train rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/qhnet_launches.csv python bench_qhnet.py --steps 1 --warmup 1 > $OUT/ncu_qhnet.log 2>&1
echo "ncu qhnet rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/gemnet_launches.csv python bench_gemnet.py --batch 64 --steps 1 --warmup 1 > $OUT/ncu_gemnet.log 2>&1
echo "ncu gemnet rc=$?"
grep -E "^\{|passed|failed|FAILED|Error|rc=|===|bf16 edge|kept vs" $OUT/log.txt | cut -c1-500 | tail -40
