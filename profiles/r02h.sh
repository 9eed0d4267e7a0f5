#!/bin/bash
OUT=gpurun_out/r02h
mkdir -p $OUT
for FLAGS in "-DSUHPE_FAST_JACOBI=0" ""; do
  echo "== K2 $FLAGS" | tee -a $OUT/ab_fisher.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>> $OUT/ab_fisher.log
  BITS=0,26 timeout 300 python profiles/time_fisher.py 23 2>&1 | grep -v Warning | tee -a $OUT/ab_fisher.log
  timeout 200 python profiles/time_small.py 2>&1 | tail -6 | tee -a $OUT/ab_fisher.log
done
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; grep -v "^\s*$" $OUT/pytest_gpu.log | grep -v DEBUG | tail -8
