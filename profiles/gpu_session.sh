#!/bin/bash
# Run on the GPU box via gpurun: parity tests, bench (both arms), ncu launch list + full capture of the top kernel.
# usage: bash profiles/gpu_session.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
lscpu | head -20 > $OUT/lscpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
# launch list of the same bench command (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --gpus 1 --steps 2 --warmup 3 --skip-extra --no-cpu --no-e2e > $OUT/bench_under_ncu.log 2>&1
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fisher_fused -s 1 -c 2 -f -o $OUT/fisher_full \
    python profiles/run_kernels.py fisher 21 > $OUT/ncu_fisher.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'laplace|metrics|select|mask|hist|ce_close' -c 16 -f -o $OUT/others_full \
    python profiles/run_kernels.py others 21 > $OUT/ncu_others.log 2>&1
# DRAM traffic of one K2 launch at the bench launch size (roofline.traffic)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fisher_fused -s 1 -c 1 --csv --log-file $OUT/k2_traffic.csv \
    python profiles/run_kernels.py fisher 23 > $OUT/ncu_traffic.log 2>&1
timeout 300 python profiles/time_fisher.py 23 > $OUT/time_fisher.log 2>&1
timeout 300 python profiles/time_k4.py > $OUT/time_k4.log 2>&1
timeout 300 python profiles/time_k2l.py > $OUT/time_k2l.log 2>&1
timeout 300 python profiles/time_ce.py > $OUT/time_ce.log 2>&1
cat $OUT/time_fisher.log $OUT/time_k4.log $OUT/time_k2l.log $OUT/time_ce.log; cat $OUT/k2_traffic.csv | tail -3
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/bench.json; tail -3 $OUT/bench.err; cat $OUT/bench_ref.json
