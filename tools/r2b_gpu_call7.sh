#!/bin/bash
# call 7: what bounds the big k_gemm_ps launches (L2 throughput of the weight-tile stream?) + QHNet back on the streaming kernels + conv staged A/B
set -u
OUT=gpurun_out/r2b_call7
mkdir -p $OUT
run() { echo "=== $*" | tee -a $OUT/log.txt; timeout -s KILL ${TMO:-900} "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" | tee -a $OUT/log.txt; }
TMO=300 run python bench_qhnet.py --steps 5 --warmup 2
NB200_QH_TP_CONV=staged TMO=300 run python bench_qhnet.py --steps 5 --warmup 2
timeout 500 ncu --metrics gpu__time_duration.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.max --clock-control none -k regex:k_gemm_ps -c 140 --csv --log-file $OUT/gemm_ps_qhnet.csv python bench_qhnet.py --steps 1 --warmup 0 > $OUT/ncu_gemm.log 2>&1
echo "ncu rc=$?"
grep -E "^\{|rc=|===" $OUT/log.txt | cut -c1-300 | tail
