#!/bin/bash
set -u
OUT=gpurun_out/r2_misc
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-600} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=400 run python -m pytest tests/test_zz_gpu_first_runs.py -q -x -rA -p no:cacheprovider
TMO=300 run python bench_gemnet.py --batch 512 --steps 3 --warmup 2
TMO=300 run python bench_train.py --steps 10 --warmup 3
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/train_launches.csv python bench_train.py --steps 1 --warmup 1 > $OUT/ncu_train.log 2>&1
echo "ncu rc=$?" >> $OUT/log.txt
grep -E "passed|failed|rc=|^\{" $OUT/log.txt | cut -c1-400
