#!/bin/bash
# multi-GPU check on one box: bench at N ranks over NCCL (global threshold; sharded e2e), reference arm under torchrun
N=${1:-2}
TAG=${2:-multi$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?" >> $OUT/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err; echo "rc=$?" >> $OUT/bench_ref_n$N.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --skip-extra > $OUT/bench_n1.json 2> $OUT/bench_n1.err
cat $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err; cat $OUT/bench_ref_n$N.json; tail -3 $OUT/bench_ref_n$N.err; cat $OUT/bench_n1.json
