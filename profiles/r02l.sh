#!/bin/bash
OUT=gpurun_out/r02l
mkdir -p $OUT
python -m semiuhpe_b200._build > /dev/null 2>&1
timeout 300 python profiles/probe.py 2>&1 | tee $OUT/probes.json
