#!/bin/bash
# K2: second UY accumulator for the second pair of a 64-pair pass (no register-rename MOVs at the back edge)
OUT=gpurun_out/r02ap
mkdir -p $OUT
python -m semiuhpe_b200._build --force > /dev/null 2>&1
timeout 300 python profiles/time_fisher.py 23 2>&1 | grep -v Warning | tee $OUT/time_fisher.txt
timeout 900 python -m pytest tests/test_gpu_fisher.py tests/test_gpu_round2.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -3 | tee $OUT/tests.log
timeout 300 python profiles/time_small.py 2>&1 | grep "K2 n=" | tee -a $OUT/time_fisher.txt
