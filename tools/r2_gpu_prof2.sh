#!/bin/bash
# launch lists (ncu, per-kernel durations) of one training step and one QHNet forward
set -u
OUT=gpurun_out/r2_prof2
mkdir -p $OUT
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/train_launches.csv python bench_train.py --steps 1 --warmup 1 > $OUT/ncu_train.log 2>&1
echo "ncu train rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/qhnet_launches.csv python bench_qhnet.py --steps 1 --warmup 1 > $OUT/ncu_qhnet.log 2>&1
echo "ncu qhnet rc=$?"
tail -2 $OUT/ncu_train.log $OUT/ncu_qhnet.log | cut -c1-400
