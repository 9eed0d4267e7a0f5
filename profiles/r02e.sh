#!/bin/bash
# r02e: butterfly diag (NFAM=1 through the butterfly too), scalar-tail A/B, failing tests again
OUT=gpurun_out/r02e
mkdir -p $OUT
echo "== BFLY=2" | tee -a $OUT/diag.log
SUHPE_NVCC_EXTRA="-DSUHPE_K2_BFLY=2" python -m semiuhpe_b200._build --force > /dev/null 2>&1
timeout 300 python profiles/diag_fwd_only.py 2>&1 | tee -a $OUT/diag.log
for FLAGS in "" "-DSUHPE_K2_STAIL=1"; do
  echo "== K2 $FLAGS" | tee -a $OUT/ab_fisher.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>> $OUT/ab_fisher.log
  BITS=0,26 timeout 300 python profiles/time_fisher.py 23 2>&1 | tee -a $OUT/ab_fisher.log
done
# parity of the scalar-tail build (library still built with the flag)
timeout 600 python -m pytest tests/test_gpu_fisher.py -m gpu -q 2>&1 | tail -4 | tee -a $OUT/ab_fisher.log
python -m semiuhpe_b200._build --force > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_torch_ops.py -m gpu -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
grep -v "^\s*$" $OUT/pytest.log | grep -v DEBUG | tail -25
