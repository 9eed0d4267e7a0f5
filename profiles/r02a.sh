#!/bin/bash
# r02a: correctness of the K2 schedule changes (round-robin Horner, 4-value butterfly, tile base re-derived)
# + A/B timing of each flag, K2L without the Newton step (parity + timing)
TAG=r02a
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python -m semiuhpe_b200._build --force > /dev/null
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
for FLAGS in "" "-DSUHPE_K2_RR=0 -DSUHPE_K2_BFLY=0" "-DSUHPE_K2_RR=0" "-DSUHPE_K2_BFLY=0" "-DSUHPE_K2_WARPS=20"; do
  echo "== K2 $FLAGS" | tee -a $OUT/ab_fisher.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>> $OUT/ab_fisher.log
  BITS=0,26 timeout 300 python profiles/time_fisher.py 23 2>&1 | tee -a $OUT/ab_fisher.log
done
for FLAGS in "" "-DSUHPE_K2L_NEWTON=0"; do
  echo "== K2L $FLAGS" | tee -a $OUT/ab_k2l.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>> $OUT/ab_k2l.log
  timeout 300 python profiles/time_k2l.py 2>&1 | tee -a $OUT/ab_k2l.log
done
# parity of the Newton-free K2L (library still built with the flag)
timeout 600 python -m pytest tests/test_gpu_laplace_metrics.py -m gpu -q > $OUT/pytest_k2l_nonewton.log 2>&1; echo "rc=$?" >> $OUT/pytest_k2l_nonewton.log
tail -15 $OUT/pytest_k2l_nonewton.log
python -m semiuhpe_b200._build --force > /dev/null
