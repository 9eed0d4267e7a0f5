#!/bin/bash
# A/B builds of K2 on the GPU box: bash profiles/ab_fisher.sh <tag> "<flags A>" "<flags B>" ...
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for FLAGS in "$@"; do
  echo "== $FLAGS" | tee -a $OUT/ab_fisher.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>> $OUT/ab_fisher.log
  BITS=26 timeout 300 python profiles/time_fisher.py 23 2>&1 | tee -a $OUT/ab_fisher.log
done
python -m semiuhpe_b200._build --force > /dev/null
