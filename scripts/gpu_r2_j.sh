#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2j.log
: > $L
run() { echo "=== $1" >> $L; env $2 timeout 300 python scripts/pipe_bench.py --in-flight 24 --graph 1 --steps 144 --tag "$1" >> $L 2>&1; }
run current "X=1"
run pb16 "TPLB_LIB_OVERRIDE=/root/repo/variants_pb16.so"
run sweepblock64 "TPLB_PROBLEM_BLOCK=64"
run sweepblock128 "TPLB_PROBLEM_BLOCK=128"
run current "X=1"
run pb16 "TPLB_LIB_OVERRIDE=/root/repo/variants_pb16.so"
echo "=== single 32768" >> $L
timeout 300 python scripts/quick_bench.py --batch 32768 --rounds 2 --keep-previous 0 --keep-records 0 --no-fp32 --reps 3 >> $L 2>&1
TPLB_LIB_OVERRIDE=/root/repo/variants_pb16.so timeout 300 python scripts/quick_bench.py --batch 32768 --rounds 2 --keep-previous 0 --keep-records 0 --no-fp32 --reps 3 >> $L 2>&1
grep -E "===|PIPE|backward |rollout  |solves/s" $L
