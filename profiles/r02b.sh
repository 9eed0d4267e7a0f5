#!/bin/bash
# r02b: ABI v2 (per-call cut_bits, in-kernel first histogram, SSL step entry, masked backward) -- tests, diag, bench
TAG=r02b
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m semiuhpe_b200._build --force > /dev/null 2> $OUT/build.err
timeout 300 python profiles/diag_fwd_only.py > $OUT/diag.log 2>&1; cat $OUT/diag.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
grep -v "^\s*$" $OUT/pytest_gpu.log | tail -60
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
cat $OUT/bench.json; tail -5 $OUT/bench.err
