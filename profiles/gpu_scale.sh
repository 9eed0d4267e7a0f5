#!/bin/bash
# scaling run on one 8-GPU box: N = 8, 4 (the driver runs 1,2,4,8 at round end); short steps
TAG=${1:-scale}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?" >> $OUT/bench_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 \
    bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $OUT/bench_ref_n8.json 2> $OUT/bench_ref_n8.err
for N in 8 4; do python - <<PY
import json
d = json.load(open("$OUT/bench_n$N.json"))
print($N, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "thr", d["threshold"], d["e2e"]["threshold"], "clocks", d["clocks"])
PY
done
tail -2 $OUT/bench_n8.err; cat $OUT/bench_ref_n8.json | cut -c1-300
