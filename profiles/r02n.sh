#!/bin/bash
OUT=gpurun_out/r02n
mkdir -p $OUT
python -m semiuhpe_b200._build --force > /dev/null 2>&1
BITS=0,26 timeout 300 python profiles/time_fisher.py 23 2>&1 | grep -v Warning | tee $OUT/time_fisher.log
timeout 200 python profiles/time_small.py 2>&1 | tail -6 | tee $OUT/time_small.log
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; grep -v "^\s*$" $OUT/pytest_gpu.log | grep -v DEBUG | tail -4
