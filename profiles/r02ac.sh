#!/bin/bash
# compute-sanitizer on every kernel path incl. the three K2L decompositions
OUT=gpurun_out/r02ac
mkdir -p $OUT
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_r02.py > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_r02.py > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/racecheck.log
tail -5 $OUT/memcheck.log; tail -5 $OUT/racecheck.log
