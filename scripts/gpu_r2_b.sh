#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2b.log
: > $L
for D in 8 16; do
  timeout 300 python scripts/pipe_bench.py --in-flight $D >> $L 2>&1
  timeout 300 python scripts/pipe_bench.py --in-flight $D --graph 1 >> $L 2>&1
done
timeout 300 python scripts/pipe_bench.py --in-flight 24 --graph 1 >> $L 2>&1
timeout 300 python scripts/pipe_bench.py --batch 16384 --in-flight 4 --graph 1 >> $L 2>&1
timeout 300 python scripts/pipe_bench.py --batch 16384 --in-flight 6 --graph 1 >> $L 2>&1
timeout 300 python scripts/pipe_bench.py --batch 32768 --in-flight 3 --graph 1 >> $L 2>&1
grep -E "PIPE|Error|error" $L
