#!/bin/bash
OUT=gpurun_out/r02k
mkdir -p $OUT
for FLAGS in "" "-DSUHPE_K2L_STAGGER=40" "-DSUHPE_K2L_STAGGER=80" "-DSUHPE_K2L_STAGGER=160" "-DSUHPE_K2L_DIAG=1" "-DSUHPE_K2L_DIAG=2" "-DSUHPE_K2L_DIAG=3"; do
  echo "== K2L $FLAGS" | tee -a $OUT/ab_k2l.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>&1
  timeout 300 python profiles/time_k2l.py 2>&1 | tee -a $OUT/ab_k2l.log
done
python -m semiuhpe_b200._build --force > /dev/null 2>&1
