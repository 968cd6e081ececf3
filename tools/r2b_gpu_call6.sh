#!/bin/bash
# call 6: QHNet pair tensor product / expansion with bulk-copied weight rows (A/B), refreshed GemNet-OC launch list
set -u
OUT=gpurun_out/r2b_call6
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-900} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=900 run python -m pytest tests/test_gpu_qhnet.py tests/test_gpu_phisnet.py -q -m gpu -rA -p no:cacheprovider
TMO=300 run python bench_qhnet.py --steps 5 --warmup 2
NB200_QH_TP_PAIR=plain TMO=300 run python bench_qhnet.py --steps 5 --warmup 2
NB200_QH_EXPAND=plain TMO=300 run python bench_qhnet.py --steps 5 --warmup 2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/gemnet_launches.csv python bench_gemnet.py --batch 64 --steps 1 --warmup 1 > $OUT/ncu_gemnet.log 2>&1
echo "ncu gemnet rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/qhnet_launches.csv python bench_qhnet.py --steps 1 --warmup 1 > $OUT/ncu_qhnet.log 2>&1
echo "ncu qhnet rc=$?"
grep -E "^\{|passed|failed|FAILED|Error|rc=|===" $OUT/log.txt | cut -c1-330 | tail -30
