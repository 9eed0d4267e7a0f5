#!/bin/bash
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for W in "$@"; do
  sed -i "s/constexpr int kWarpsPerBlock = [0-9]*;/constexpr int kWarpsPerBlock = $W;/" semiuhpe_b200/csrc/fisher_kernels.cu
  python -m semiuhpe_b200._build --force > /dev/null 2>$OUT/build_$W.err || { echo "build failed W=$W"; cat $OUT/build_$W.err; continue; }
  echo "== warps per block $W (cycles column assumes 16 warps: scale by W/16)" | tee -a $OUT/ab.log
  python profiles/body_probe.py 2>&1 | grep full | tee -a $OUT/ab.log
done
sed -i "s/constexpr int kWarpsPerBlock = [0-9]*;/constexpr int kWarpsPerBlock = 16;/" semiuhpe_b200/csrc/fisher_kernels.cu
