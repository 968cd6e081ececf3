#!/bin/bash
# ncu --set full, one launch each, of the kernels the first capture window (tools/r2_gpu_ncu.sh) missed
set -u
OUT=gpurun_out/r2_ncu
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
PB="python bench.py --steps 2 --warmup 2 --streams 1 --skip-e2e --no-cpu-baseline --no-train"
timeout 300 $NCU -k regex:'k_qh_expand$' -s 2 -c 1 -o $OUT/qh_expand -f python bench_qhnet.py --steps 1 --warmup 0 > $OUT/qh_expand.log 2>&1; echo "expand rc=$?"
timeout 300 $NCU -k regex:k_node_bwd -s 8 -c 1 -o $OUT/node_bwd -f $PB > $OUT/node_bwd.log 2>&1; echo "node_bwd rc=$?"
timeout 300 $NCU -k regex:k_painn_msg_bwd -s 8 -c 1 -o $OUT/msg_bwd -f $PB > $OUT/msg_bwd.log 2>&1; echo "msg_bwd rc=$?"
timeout 300 $NCU -k regex:k_filter -s 2 -c 1 -o $OUT/filter -f $PB > $OUT/filter.log 2>&1; echo "filter rc=$?"
timeout 300 $NCU -k regex:k_wgrad_tc -s 80 -c 2 -o $OUT/wgrad -f python bench_train.py --steps 1 --warmup 1 > $OUT/wgrad.log 2>&1; echo "wgrad rc=$?"
ls -la $OUT | grep rep
