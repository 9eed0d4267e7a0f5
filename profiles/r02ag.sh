#!/bin/bash
OUT=gpurun_out/r02ag
mkdir -p $OUT
timeout 300 python profiles/sweep_k2l.py 2>&1 | grep -v Warning | tee $OUT/sweep.log
timeout 900 python -m pytest tests/test_gpu_laplace_metrics.py tests/test_gpu_round2.py tests/test_torch_ops.py -x -q -m gpu 2>&1 | tail -3 | tee $OUT/tests.log
bash profiles/r02ac.sh
