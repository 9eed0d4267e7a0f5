#!/bin/bash
OUT=gpurun_out/r02p
mkdir -p $OUT
for FLAGS in "" "-DSUHPE_K2_PIPE=1"; do
  echo "== K2 $FLAGS" | tee -a $OUT/ab.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>&1
  BITS=0,26 timeout 300 python profiles/time_fisher.py 23 2>&1 | grep -v Warning | tee -a $OUT/ab.log
  timeout 200 python profiles/time_small.py 2>&1 | grep "K2 n=" | tee -a $OUT/ab.log
done
timeout 900 python -m pytest tests/test_gpu_fisher.py tests/test_gpu_fisher_ce.py tests/test_gpu_pipeline.py -m gpu -q 2>&1 | tail -3 | tee -a $OUT/ab.log
python -m semiuhpe_b200._build --force > /dev/null 2>&1
