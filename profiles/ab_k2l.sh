#!/bin/bash
# A/B builds of K2L on the GPU box: bash profiles/ab_k2l.sh <tag> "<flags A>" "<flags B>" ...
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for FLAGS in "$@"; do
  echo "== $FLAGS" | tee -a $OUT/ab_k2l.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>> $OUT/ab_k2l.log
  timeout 300 python profiles/time_k2l.py 2>&1 | tee -a $OUT/ab_k2l.log
done
python -m semiuhpe_b200._build --force > /dev/null
