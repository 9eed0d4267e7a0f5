#!/bin/bash
# A/B the fused Fisher kernel geometry on the GPU box: bash profiles/ab_fisher.sh <tag> <warps...>
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for W in "$@"; do
  sed -i "s/constexpr int kWarpsPerBlock = [0-9]*;/constexpr int kWarpsPerBlock = $W;/" semiuhpe_b200/csrc/fisher_kernels.cu
  python -m semiuhpe_b200._build --force > /dev/null 2>$OUT/build_$W.err || { echo "build failed W=$W"; cat $OUT/build_$W.err; continue; }
  echo "== warps per block $W" | tee -a $OUT/ab.log
  python profiles/time_fisher.py 23 2>&1 | tee -a $OUT/ab.log
done
