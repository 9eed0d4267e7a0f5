#!/bin/bash
OUT=gpurun_out/r02j
mkdir -p $OUT
python -m semiuhpe_b200._build > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_r02.py > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_r02.py > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/racecheck.log
tail -5 $OUT/memcheck.log; tail -5 $OUT/racecheck.log
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "ema or differentiable or quats" 2>&1 | tail -5 | tee $OUT/pytest.log
for T in 512 640; do
  echo "== K2L threads $T" | tee -a $OUT/ab_k2l.log
  SUHPE_NVCC_EXTRA="-DSUHPE_K2L_THREADS=$T" python -m semiuhpe_b200._build --force > /dev/null 2>&1
  timeout 300 python profiles/time_k2l.py 2>&1 | tee -a $OUT/ab_k2l.log
done
python -m semiuhpe_b200._build --force > /dev/null 2>&1
