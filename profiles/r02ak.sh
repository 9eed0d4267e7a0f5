#!/bin/bash
# K2L per-sample set-up: fp32 Jacobi + 2 fp64 polishing sweeps against 7 fp64 sweeps from scratch
OUT=gpurun_out/r02ak
mkdir -p $OUT
for FLAGS in "-DSUHPE_K2L_SETUP_F64_ONLY=1" ""; do
  echo "== $FLAGS" | tee -a $OUT/ab_setup.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>&1
  timeout 300 python profiles/time_k2l.py 2>&1 | grep -v Warning | tee -a $OUT/ab_setup.log
  timeout 300 python profiles/sweep_k2l_small.py 2>&1 | grep -v Warning | tee -a $OUT/ab_setup.log
  timeout 300 python profiles/sweep_k2l.py 2>&1 | grep -v Warning | head -8 | tee -a $OUT/ab_setup.log
done
timeout 900 python -m pytest tests/test_gpu_laplace_metrics.py tests/test_torch_ops.py tests/test_gpu_round2.py -x -q -m gpu 2>&1 | tail -3 | tee $OUT/tests.log
