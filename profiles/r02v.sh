#!/bin/bash
# K2L: sample-packed stream kernel (two samples per thread, one comparison per trip) against the point-packed one
OUT=gpurun_out/r02v
mkdir -p $OUT
for FLAGS in "" "-DSUHPE_K2L_PACK_SAMPLES=1 -DSUHPE_K2L_S2_NP=2" "-DSUHPE_K2L_PACK_SAMPLES=1 -DSUHPE_K2L_S2_NP=4"; do
  echo "== $FLAGS" | tee -a $OUT/ab_k2l_b.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>> $OUT/ab_k2l_b.log
  timeout 300 python profiles/time_k2l.py 2>&1 | grep -v Warning | tee -a $OUT/ab_k2l_b.log
  if [ -n "$FLAGS" ]; then
    timeout 900 python -m pytest tests/test_gpu_laplace_metrics.py tests/test_gpu_round2.py tests/test_torch_ops.py -x -q -m gpu 2>&1 | tail -5 | tee -a $OUT/ab_k2l_b.log
  fi
done
python -m semiuhpe_b200._build --force > /dev/null
