#!/bin/bash
# final call of round 2: A/B of the warp-per-edge triplet kernel, the whole GPU suite, the bench line as the driver runs it, secondary benches
set -u
OUT=gpurun_out/r2b_final
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-600} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=900 run python -m pytest tests -q -m gpu -p no:cacheprovider
TMO=400 run python bench.py
TMO=400 run python bench_gemnet.py --steps 3 --warmup 1
NB200_GOC_TRIP=functor TMO=400 run python bench_gemnet.py --steps 2 --warmup 1
NB200_GEMM_2G=0 TMO=400 run python bench_gemnet.py --steps 2 --warmup 1
TMO=300 run python bench_qhnet.py --steps 5 --warmup 2
TMO=300 run python bench_train.py --steps 10 --warmup 3 --storage bf16
TMO=300 run python bench_train.py --steps 10 --warmup 3
TMO=300 run python bench_opt.py
python - <<'PY' > $OUT/smoke.txt 2>&1
import __graft_entry__ as g
g.smoke()
PY
echo "smoke rc=$?"; cat $OUT/smoke.txt | tail -2
grep -E "^\{|passed|failed|FAILED|Error|rc=|===" $OUT/log.txt | cut -c1-2500 | tail -40
