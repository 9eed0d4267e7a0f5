#!/bin/bash
OUT=gpurun_out/r02ah
mkdir -p $OUT
for FLAGS in "-DSUHPE_K2L_BLOCK_KERNEL=0" "-DSUHPE_K2L_BLOCK_KERNEL=100000" ""; do
  echo "== $FLAGS" | tee -a $OUT/sweep_small.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>&1
  timeout 300 python profiles/sweep_k2l_small.py 2>&1 | grep -v Warning | tee -a $OUT/sweep_small.log
done
timeout 900 python -m pytest tests/test_gpu_laplace_metrics.py tests/test_torch_ops.py -x -q -m gpu 2>&1 | tail -3 | tee $OUT/tests.log
