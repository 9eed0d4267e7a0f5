#!/bin/bash
# Full round-2 session on one B200 (gpurun): parity tests, smoke, bench (both arms), ncu launch list of the bench command,
# --set full captures of K2 AT THE BENCH LAUNCH SIZE (2^23) and of the other kernels, K2 DRAM traffic, per-kernel timings.
# usage: bash profiles/gpu_session_r02.sh <tag>
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
lscpu | head -20 > $OUT/lscpu.txt
python -m semiuhpe_b200._build --force > /dev/null 2> $OUT/build.err
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
# launch list of the same bench command (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --gpus 1 --steps 2 --warmup 3 --skip-extra --no-cpu --no-e2e > $OUT/bench_under_ncu.log 2>&1
# full capture of the dominant kernel at the bench launch size
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fisher_fused -s 1 -c 1 -f -o $OUT/fisher_full \
    python profiles/run_kernels.py fisher 23 > $OUT/ncu_fisher.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'laplace|metrics|select|mask|hist|ce_close|scale_rows|ssl_finalize' -c 16 -f -o $OUT/others_full \
    python profiles/run_kernels.py others 21 > $OUT/ncu_others.log 2>&1
# DRAM traffic of one K2 launch at the bench launch size (roofline.traffic)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fisher_fused -s 1 -c 1 --csv --log-file $OUT/k2_traffic.csv \
    python profiles/run_kernels.py fisher 23 > $OUT/ncu_traffic.log 2>&1
timeout 300 python profiles/time_fisher.py 23 > $OUT/time_fisher.log 2>&1
timeout 300 python profiles/time_k4.py > $OUT/time_k4.log 2>&1
timeout 300 python profiles/time_k2l.py > $OUT/time_k2l.log 2>&1
timeout 300 python profiles/time_ce.py > $OUT/time_ce.log 2>&1
timeout 300 python profiles/time_small.py > $OUT/time_small.log 2>&1
timeout 300 python profiles/host_overhead.py > $OUT/host_overhead.log 2>&1
cat $OUT/time_fisher.log $OUT/time_k4.log $OUT/time_k2l.log $OUT/time_ce.log $OUT/time_small.log; tail -3 $OUT/k2_traffic.csv
grep -v "^\s*$" $OUT/pytest_gpu.log | grep -v DEBUG | tail -3; tail -2 $OUT/smoke.log; cat $OUT/bench.json; tail -3 $OUT/bench.err; cat $OUT/bench_ref.json
