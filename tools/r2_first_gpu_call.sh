#!/bin/bash
# First GPU call of the next round, ONE gpurun invocation (budget: ~12 box-minutes):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r2_first_gpu_call.sh'
# Validates on a device what round 1 could only check under host emulation (GemNet-OC forward, SchNet training, the GEMM at the new shapes) and
# takes the first measurements of both, plus a launch list of one GemNet-OC forward.  Everything lands in gpurun_out/r2_first/.
set -u
OUT=gpurun_out/r2_first
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout 600 "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
# 1. device parity of the two functor engines (xfail(strict=False): look for XPASS / xfail and the printed numbers)
run python -m pytest tests/test_zz_gpu_first_runs.py -q -rA -p no:cacheprovider
# 2. GemNet-OC: small batch first (cheap failure), then BASELINE config 5; tcgen05 vs functor GEMM
run python bench_gemnet.py --batch 32 --steps 3 --warmup 3 --cpu
run python bench_gemnet.py --batch 512 --steps 3 --warmup 3
run python bench_gemnet.py --batch 64 --steps 2 --warmup 1 --simt
run python bench_gemnet.py --train --batch 16 --steps 3 --warmup 2
# 3. SchNet training step (E+F and E-only)
run python bench_train.py --model schnet --batch 256 --steps 5 --warmup 3
run python bench_train.py --model schnet --batch 256 --steps 5 --warmup 3 --loss e
# 4. launch list of one GemNet-OC forward (times under ncu are NOT bench values)
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/gemnet_launches.csv \
    python bench_gemnet.py --batch 64 --steps 1 --warmup 1 > $OUT/ncu_gemnet.log 2>&1
echo "ncu rc=$?" >> $OUT/log.txt
tail -5 $OUT/log.txt
