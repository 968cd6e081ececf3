#!/bin/bash
# last call of round 2: the whole GPU suite and the bench line on the final tree
set -u
OUT=gpurun_out/r2b_final2
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-600} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=900 run python -m pytest tests -q -m gpu -p no:cacheprovider
TMO=400 run python bench.py
python - <<'PY' > $OUT/smoke.txt 2>&1
import __graft_entry__ as g
g.smoke()
PY
echo "smoke rc=$?"; tail -1 $OUT/smoke.txt
grep -E "^\{|passed|failed|FAILED|Error|rc=|===" $OUT/log.txt | cut -c1-300 | tail -12
