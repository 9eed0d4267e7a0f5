#!/bin/bash
# final-tree check on 2 GPUs: NCCL parity test + bench both arms under torchrun
OUT=gpurun_out/r02am
mkdir -p $OUT
python -m semiuhpe_b200._build > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k nccl > $OUT/pytest_nccl.log 2>&1; echo "rc=$?" >> $OUT/pytest_nccl.log; tail -3 $OUT/pytest_nccl.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err; echo "ref rc=$?"
cat $OUT/bench_n2.json | cut -c1-1500; cat $OUT/bench_ref_n2.json | cut -c1-300; tail -3 $OUT/bench_n2.err
