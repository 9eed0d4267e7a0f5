#!/bin/bash
# K2L sample-packed kernel as thread-block clusters (grid sliced across the CTAs of a cluster, DSMEM merge): parity + batch sweep
OUT=gpurun_out/r02af
mkdir -p $OUT
python -m semiuhpe_b200._build --force > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_laplace_metrics.py -x -q -m gpu 2>&1 | tail -15 | tee $OUT/tests.log
for FLAGS in "-DSUHPE_K2L_PACK_SAMPLES=0" "-DSUHPE_K2L_PACK_SAMPLES=2" "-DSUHPE_K2L_PACK_SAMPLES=1"; do
  echo "== $FLAGS" | tee -a $OUT/sweep.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>&1
  timeout 300 python profiles/sweep_k2l.py 2>&1 | grep -v Warning | tee -a $OUT/sweep.log
done
python -m semiuhpe_b200._build --force > /dev/null 2>&1
timeout 300 python profiles/time_k2l.py 2>&1 | grep -v Warning | tee $OUT/time_k2l.txt
