#!/bin/bash
OUT=gpurun_out/r02m
mkdir -p $OUT
for FLAGS in "-DSUHPE_K2L_MBLOCK=0" ""; do
  echo "== K2L $FLAGS" | tee -a $OUT/ab_k2l.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>&1
  timeout 300 python profiles/time_k2l.py 2>&1 | tee -a $OUT/ab_k2l.log
done
timeout 600 python -m pytest tests/test_gpu_laplace_metrics.py -m gpu -q 2>&1 | tail -3 | tee -a $OUT/ab_k2l.log
