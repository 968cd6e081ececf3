#!/bin/bash
# full GPU suite + the bench line as the driver runs it
set -u
OUT=gpurun_out/r2_full
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-900} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=900 run python -m pytest tests -q -x -m gpu -rA -p no:cacheprovider
TMO=300 run python bench.py
TMO=300 run python bench.py --streams 1 --no-cpu-baseline
grep -v "^{" $OUT/log.txt | grep -E "passed|failed|FAILED|Error|error|rc=" | tail -30
