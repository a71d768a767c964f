#!/bin/bash
# round-2 experiment batch A: fused sweep vs separate kernels
mkdir -p gpurun_out
L=gpurun_out/r2a.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $L 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $L
tail -5 gpurun_out/r2a_pytest.log >> $L
for prev in 0 1; do
  timeout 300 python scripts/quick_bench.py --batch 32768 --rounds 2 --keep-previous $prev --keep-records 0 --no-fp32 --reps 3 >> $L 2>&1
  TPLB_NO_FUSED_SWEEP=1 timeout 300 python scripts/quick_bench.py --batch 32768 --rounds 2 --keep-previous $prev --keep-records 0 --no-fp32 --reps 3 >> $L 2>&1
done
for D in 8 12; do
  timeout 300 python scripts/pipe_bench.py --in-flight $D >> $L 2>&1
  TPLB_NO_FUSED_SWEEP=1 timeout 300 python scripts/pipe_bench.py --in-flight $D >> $L 2>&1
done
timeout 300 python scripts/pipe_bench.py --in-flight 8 --keep-previous 1 --keep-records 1 >> $L 2>&1
timeout 300 python scripts/pipe_bench.py --batch 8192 --in-flight 4 >> $L 2>&1
timeout 300 python scripts/pipe_bench.py --batch 16384 --in-flight 4 >> $L 2>&1
timeout 300 python scripts/quick_bench.py --batch 65536 --rounds 2 --keep-previous 0 --keep-records 0 --no-fp32 --reps 3 >> $L 2>&1
grep -E "PIPE|solves/s|rc=|passed|failed" $L
