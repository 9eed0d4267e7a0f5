#!/bin/bash
# K4 iteration loop on the GPU box: parity tests, K4 timing, optional ncu of the metrics kernel, short bench
TAG=${1:-k4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python profiles/time_k4.py > $OUT/time_k4.log 2>&1
if [ "$2" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:metrics_kernel -s 3 -c 1 -f -o $OUT/metrics_full \
    python profiles/time_k4.py 10000000 > $OUT/ncu_metrics.log 2>&1
fi
if [ "$3" == "bench" ]; then
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
for c in 524288 2097152; do
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --skip-extra --e2e-chunk $c > $OUT/bench_chunk$c.json 2>> $OUT/bench.err
done
fi
tail -5 $OUT/pytest_gpu.log; cat $OUT/time_k4.log; cat $OUT/bench*.json 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'e2e', d['e2e'], 'extra', d.get('extra'))
"
