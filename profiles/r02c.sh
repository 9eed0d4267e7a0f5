#!/bin/bash
OUT=gpurun_out/r02c
mkdir -p $OUT
for FLAGS in "" "-DSUHPE_K2_RR=0" "-DSUHPE_K2_BFLY=0" "-DSUHPE_K2_RR=0 -DSUHPE_K2_BFLY=0"; do
  echo "== $FLAGS" | tee -a $OUT/diag.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>&1
  timeout 300 python profiles/diag_fwd_only.py 2>&1 | tee -a $OUT/diag.log
done
python -m semiuhpe_b200._build --force > /dev/null 2>&1
