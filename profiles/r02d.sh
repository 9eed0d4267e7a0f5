#!/bin/bash
# r02d: forward-only vs full launch diag on K2 variants, host overhead of the small-batch steps, failing tests again
OUT=gpurun_out/r02d
mkdir -p $OUT
for FLAGS in "-DSUHPE_K2_RR=0" "-DSUHPE_K2_BFLY=0" ""; do
  echo "== $FLAGS" | tee -a $OUT/diag.log
  SUHPE_NVCC_EXTRA="$FLAGS" python -m semiuhpe_b200._build --force > /dev/null 2>&1
  timeout 300 python profiles/diag_fwd_only.py 2>&1 | tee -a $OUT/diag.log
done
timeout 600 python profiles/host_overhead.py > $OUT/host_overhead.log 2>&1; cat $OUT/host_overhead.log
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_torch_ops.py -m gpu -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
grep -v "^\s*$" $OUT/pytest.log | tail -40
