#!/bin/bash
# usage: r2_gpu_scale.sh N   -- bench.py and bench_train.py at N GPUs (torchrun, one rank per GPU), JSON lines into gpurun_out/r2_scale/
set -u
N=$1
OUT=gpurun_out/r2_scale
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_n$N.out 2> $OUT/bench_n$N.err
echo "bench rc=$?"
grep '^{' $OUT/bench_n$N.out > $OUT/bench_n$N.json
timeout 300 $TR bench_train.py --steps 10 --warmup 3 > $OUT/train_n$N.out 2> $OUT/train_n$N.err
echo "train rc=$?"
grep '^{' $OUT/train_n$N.out > $OUT/train_n$N.json
cut -c1-1500 $OUT/bench_n$N.json; cut -c1-600 $OUT/train_n$N.json; tail -3 $OUT/bench_n$N.err
