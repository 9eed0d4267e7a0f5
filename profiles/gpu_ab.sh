#!/bin/bash
# A/B iteration on the GPU box: GPU parity tests, K2 alone (cut off / on), a short bench
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python profiles/time_fisher.py 23 > $OUT/time_fisher.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
tail -5 $OUT/pytest_gpu.log; cat $OUT/time_fisher.log; cat $OUT/bench.json; tail -3 $OUT/bench.err
