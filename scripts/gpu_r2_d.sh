#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2d.log
: > $L
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -3 gpurun_out/r2d_pytest.log >> $L
timeout 300 python scripts/quick_bench.py --batch 32768 --rounds 2 --keep-previous 0 --keep-records 0 --no-fp32 --reps 3 >> $L 2>&1
timeout 300 python scripts/pipe_bench.py --in-flight 16 --graph 1 --steps 96 >> $L 2>&1
timeout 300 python scripts/pipe_bench.py --in-flight 24 --graph 1 --steps 96 >> $L 2>&1
timeout 300 python scripts/pipe_bench.py --batch 16384 --in-flight 6 --graph 1 >> $L 2>&1
grep -E "PIPE|solves/s|rc=|passed|failed|backward |rollout  " $L
