#!/bin/bash
# r02f (gpurun --gpus 8): what bounds the N-GPU end-to-end leg?  topology, N-rank copy-only ceiling (pinned / unpinned
# ranks), the NCCL threshold-parity test on 2 GPUs, bench at N=8 (threshold_check; e2e with and without core pinning), N=2
OUT=gpurun_out/r02f
mkdir -p $OUT
{ nvidia-smi topo -m; echo; nproc; lscpu | head -25; echo; (numactl -H 2>/dev/null || echo "numactl: not installed");
  echo; for d in /sys/bus/pci/devices/*; do v=$(cat $d/vendor 2>/dev/null); c=$(cat $d/class 2>/dev/null);
    if [ "$v" = "0x10de" ]; then echo "$(basename $d) class=$c numa_node=$(cat $d/numa_node) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done;
  echo; lspci -tv 2>/dev/null | head -60; echo; free -g | head -3; } > $OUT/topo.txt 2>&1
python -m semiuhpe_b200._build > /dev/null 2>&1
timeout 600 python profiles/probe_link.py --ranks 1,2,4,8 > $OUT/link_unpinned.jsonl 2> $OUT/link_unpinned.err
timeout 400 python profiles/probe_link.py --ranks 4,8 --pin > $OUT/link_pinned.jsonl 2> $OUT/link_pinned.err
cat $OUT/link_unpinned.jsonl $OUT/link_pinned.jsonl
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k nccl > $OUT/pytest_nccl.log 2>&1; echo "rc=$?" >> $OUT/pytest_nccl.log; tail -4 $OUT/pytest_nccl.log
run_bench() { # N extra-flags tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $1 --steps 5 --warmup 3 --skip-extra --no-cpu $2 > $OUT/bench_$3.json 2> $OUT/bench_$3.err
  echo "bench $3 rc=$?"; cat $OUT/bench_$3.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('n_gpus','value','ms_per_step')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('host_cores_per_rank')); print(json.dumps(d['threshold_check']))"
}
run_bench 8 "" n8
run_bench 8 "--no-pin" n8_nopin
run_bench 2 "" n2
run_bench 4 "" n4
