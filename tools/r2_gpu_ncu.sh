#!/bin/bash
# one ncu --set full capture per hot kernel family (QHNet tensor products / expansion, PaiNN node + message + filter kernels, split-K GEMM, weight gradients)
set -u
OUT=gpurun_out/r2_ncu
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:'k_qh_tp_pair|k_qh_tp_conv|k_qh_expand$' -s 6 -c 6 -o $OUT/qhnet -f python bench_qhnet.py --steps 1 --warmup 0 > $OUT/qhnet.log 2>&1; echo "qhnet rc=$?"
timeout 600 $NCU -k regex:'k_node_fwd|k_node_bwd|k_painn_msg_fwd|k_painn_msg_bwd|k_filter' -s 30 -c 10 -o $OUT/painn -f python bench.py --steps 2 --warmup 2 --streams 1 --skip-e2e --no-cpu-baseline --no-train > $OUT/painn.log 2>&1; echo "painn rc=$?"
timeout 600 $NCU -k regex:'k_wgrad_tc|k_gemm_ps' -s 40 -c 8 -o $OUT/train -f python bench_train.py --steps 1 --warmup 0 > $OUT/train.log 2>&1; echo "train rc=$?"
ls -la $OUT
