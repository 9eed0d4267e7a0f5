#!/bin/bash
OUT=gpurun_out/r02i
mkdir -p $OUT
for FLAGS in "-DSUHPE_FAST_JACOBI=1" "-DSUHPE_FAST_JACOBI=2"; do
  echo "== K2 $FLAGS" | tee -a $OUT/ab.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>> $OUT/ab.log
  BITS=26 timeout 300 python profiles/time_fisher.py 23 2>&1 | grep -v Warning | tee -a $OUT/ab.log
  timeout 600 python -m pytest tests/test_gpu_fisher.py tests/test_gpu_fisher_ce.py -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|worst" | tee -a $OUT/ab.log
done
