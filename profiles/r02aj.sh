#!/bin/bash
OUT=gpurun_out/r02aj
mkdir -p $OUT
python -m semiuhpe_b200._build --force > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_laplace_metrics.py tests/test_torch_ops.py tests/test_gpu_round2.py -x -q -m gpu 2>&1 | tail -3 | tee $OUT/tests.log
timeout 300 python profiles/sweep_k2l.py 2>&1 | grep -v Warning | tee $OUT/sweep.log
timeout 300 python profiles/sweep_k2l_small.py 2>&1 | grep -v Warning | tee -a $OUT/sweep.log
timeout 300 python profiles/time_k2l.py 2>&1 | grep -v Warning | tee -a $OUT/sweep.log
