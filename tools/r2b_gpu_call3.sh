#!/bin/bash
# call 3: merged 3-term weight-gradient launches, GemNet-OC fused GEMM tails + cp.async quadruplet kernel; ncu --set full of three kernels
set -u
OUT=gpurun_out/r2b_call3
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-900} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=900 run python -m pytest tests/test_gpu_train.py tests/test_zz_gpu_first_runs.py -q -m gpu -rA -p no:cacheprovider
TMO=300 run python bench_train.py --steps 10 --warmup 3
TMO=300 run python bench_train.py --steps 10 --warmup 3 --storage bf16
TMO=600 run python bench_gemnet.py --steps 3 --warmup 1
NB200_GOC_TAILS=separate TMO=600 run python bench_gemnet.py --steps 2 --warmup 1
for k in MulRbfRowsK k_quad_edges k_filter_wgrad_bal; do
  if [ $k = k_filter_wgrad_bal ]; then CMD="python bench_train.py --steps 1 --warmup 1"; else CMD="python bench_gemnet.py --batch 64 --steps 1 --warmup 0"; fi
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $OUT/ncu_$k -f $CMD > $OUT/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
grep -E "^\{|passed|failed|FAILED|Error|rc=|===|bf16 edge|kept vs" $OUT/log.txt | cut -c1-420 | tail -40
