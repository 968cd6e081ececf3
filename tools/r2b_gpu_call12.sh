#!/bin/bash
# call 12 (last GPU minutes of the round): GemNet-OC device parity + cfg-5 forward after the 16-byte loads in AggAtomRbfK / QuadXtK
set -u
OUT=gpurun_out/r2b_call12
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-300} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=300 run python -m pytest tests/test_zz_gpu_first_runs.py -q -m gpu -p no:cacheprovider -k gemnet
TMO=200 run python bench_gemnet.py --steps 3 --warmup 1
grep -E "^\{|passed|failed|FAILED|Error|rc=|===" $OUT/log.txt | cut -c1-300 | tail -8
