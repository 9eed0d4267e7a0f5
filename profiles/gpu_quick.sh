#!/bin/bash
# quick iteration loop on the GPU box: fisher parity tests, a short bench, optional ncu of the fused kernel
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
if [ "$2" == "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fisher_fused -s 1 -c 1 -f -o $OUT/fisher_full \
    python profiles/run_kernels.py fisher 21 > $OUT/ncu_fisher.log 2>&1
fi
tail -5 $OUT/pytest_gpu.log; cat $OUT/bench.json; tail -3 $OUT/bench.err
