#!/bin/bash
# call 4: cp.async ring in the filter weight-gradient kernel, 16-byte loads in MulRbfRowsK
set -u
OUT=gpurun_out/r2b_call4
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-900} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=900 run python -m pytest tests/test_gpu_train.py tests/test_zz_gpu_first_runs.py -q -m gpu -rA -p no:cacheprovider -k "not schnet"
TMO=300 run python bench_train.py --steps 10 --warmup 3
TMO=300 run python bench_train.py --steps 10 --warmup 3 --storage bf16
TMO=600 run python bench_gemnet.py --steps 3 --warmup 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/train_f32_launches.csv python bench_train.py --steps 1 --warmup 1 > $OUT/ncu_train.log 2>&1
echo "ncu train rc=$?"
grep -E "^\{|passed|failed|FAILED|Error|rc=|===|bf16 edge|kept vs" $OUT/log.txt | cut -c1-420 | tail -30
