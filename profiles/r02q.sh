#!/bin/bash
OUT=gpurun_out/r02q
mkdir -p $OUT
for FLAGS in "" "-DSUHPE_K2_UNROLL=2"; do
  echo "== K2 $FLAGS" | tee -a $OUT/ab.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>&1
  BITS=0,26 timeout 300 python profiles/time_fisher.py 23 2>&1 | grep -v Warning | tee -a $OUT/ab.log
done
python -m semiuhpe_b200._build --force > /dev/null 2>&1
