#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2f.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -3 gpurun_out/r2f_pytest.log >> $L
for cfg in "TPLB_ROLLOUT_PAIR=0" "TPLB_ROLLOUT_PAIR=1" "TPLB_ROLLOUT_PAIR=1 TPLB_ROLLOUT_PAIR_BLOCK=64" "TPLB_ROLLOUT_PAIR=0" "TPLB_ROLLOUT_PAIR=1"; do
  echo "=== $cfg" >> $L
  env $cfg timeout 300 python scripts/pipe_bench.py --in-flight 24 --graph 1 --steps 144 --tag "$cfg" >> $L 2>&1
  env $cfg timeout 300 python scripts/quick_bench.py --batch 32768 --rounds 2 --keep-previous 0 --keep-records 0 --no-fp32 --reps 3 >> $L 2>&1
done
grep -E "===|PIPE|solves/s|backward |rollout  |rc=|passed|failed" $L
