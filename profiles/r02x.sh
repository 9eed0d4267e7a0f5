#!/bin/bash
OUT=gpurun_out/r02x
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_laplace_metrics.py -x -q -m gpu 2>&1 | tail -30 | tee $OUT/tests.log
timeout 300 python profiles/time_k2l.py 2>&1 | grep -v Warning | tee $OUT/time_k2l.txt
