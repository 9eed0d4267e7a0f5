#!/bin/bash
OUT=gpurun_out/r02ae
mkdir -p $OUT
for FLAGS in "-DSUHPE_K2L_PACK_SAMPLES=0" "-DSUHPE_K2L_PACK_SAMPLES=2" "-DSUHPE_K2L_PACK_SAMPLES=1"; do
  echo "== $FLAGS" | tee -a $OUT/sweep.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>&1
  timeout 300 python profiles/sweep_k2l.py 2>&1 | grep -v Warning | tee -a $OUT/sweep.log
done
timeout 600 python -m pytest tests/test_gpu_laplace_metrics.py -x -q -m gpu 2>&1 | tail -2 | tee -a $OUT/sweep.log
