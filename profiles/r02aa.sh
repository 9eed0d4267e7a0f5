#!/bin/bash
# ncu --set full of the sample-packed K2L kernel (2^18 samples x 4608 points), with the source page
OUT=gpurun_out/r02aa
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:laplace_stream2 -s 1 -c 1 -f -o $OUT/k2l_full \
    python profiles/run_kernels.py laplace 21 > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
