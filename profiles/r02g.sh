#!/bin/bash
OUT=gpurun_out/r02g
mkdir -p $OUT
python -m semiuhpe_b200._build > /dev/null 2>&1
timeout 300 python profiles/diag_eager_after_cpu.py 2>&1 | tee $OUT/diag_eager.log
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; grep -v "^\s*$" $OUT/pytest_gpu.log | grep -v DEBUG | tail -6
