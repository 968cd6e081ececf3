#!/bin/bash
# call 8: role timing inside k_gemm_ps (profiling build made on the box)
set -u
OUT=gpurun_out/r2b_call8
mkdir -p $OUT
NB200_NVCC_EXTRA=-DNF_PROF timeout 600 python -m nabladft_b200.build --force > $OUT/build.log 2>&1
echo "build rc=$?"
for shape in "16384 8320 128" "76600 512 512" "100096 640 640" "65536 5376 32"; do
  timeout 120 python tools/gemm_ps_prof.py $shape >> $OUT/gemm_ps_prof.txt 2>&1
done
cat $OUT/gemm_ps_prof.txt
