#!/bin/bash
# K2L sample-packed kernel with its 16 MUFU per trip replaced by FMULs (timing diagnostic, wrong results) + sanitizer run
OUT=gpurun_out/r02ad
mkdir -p $OUT
for FLAGS in "-DSUHPE_K2L_DIAG_NOMUFU=1" ""; do
  echo "== $FLAGS" | tee -a $OUT/ab_k2l.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>> $OUT/ab_k2l.log
  timeout 300 python profiles/time_k2l.py 2>&1 | grep -v Warning | tee -a $OUT/ab_k2l.log
done
timeout 600 python -m pytest tests/test_gpu_laplace_metrics.py -x -q -m gpu 2>&1 | tail -2 | tee -a $OUT/ab_k2l.log
bash profiles/r02ac.sh
