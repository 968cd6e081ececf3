#!/bin/bash
# A/B of the node-kernel worker layout: ab/g1.so = one worker group, ab/g2.so = loader + epilogue groups (-DNF_TWO_GROUPS); *p.so = with -DNF_PROF
set -u
OUT=gpurun_out/r2_ab
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-400} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
for v in g2 g1; do
  cp ab/$v.so nabladft_b200/libnabla_b200.so
  echo "##### variant $v" | tee -a $OUT/log.txt
  TMO=300 run python -m pytest tests/test_gpu_painn.py -q -x -p no:cacheprovider -k "fused or golden or cfg2_slice or gemm"
  TMO=200 run python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train --streams 1
  TMO=200 run python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train
  TMO=200 run python tools/gemm_microbench.py
  cp ab/${v}p.so nabladft_b200/libnabla_b200.so
  TMO=200 run python tools/nf_prof.py
done
cp ab/g1.so nabladft_b200/libnabla_b200.so
echo "##### full checks (g1)" | tee -a $OUT/log.txt
TMO=200 run python -m pytest tests/test_gpu_painn.py -q -x -p no:cacheprovider -k "wgrad" -rP
TMO=700 run python -m pytest tests/test_gpu_train.py tests/test_gpu_painn.py tests/test_gpu_qhnet.py tests/test_zz_gpu_first_runs.py -q -x -p no:cacheprovider
TMO=300 run python bench_train.py --steps 10 --warmup 3
TMO=300 run python bench_train.py --steps 10 --warmup 3 --loss e
TMO=300 run env NB200_WGRAD=cublas python bench_train.py --steps 10 --warmup 3
TMO=300 run python bench_qhnet.py --profile
grep -E "#####|wgrad M|PASSED|FAILED|passed|failed|cycles|CTA runs|rc=|\"value\"|ms|Error" $OUT/log.txt | cut -c1-900 | tail -120
