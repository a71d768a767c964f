#!/bin/bash
# one batch alone: latency sequence (rounds 1) against throughput sequence (rounds 2), default keep flags
for m in lateral velocity mpc; do for B in 4096 8192; do
  for r in 1 2; do
    echo -n "$m B=$B rounds=$r: "; python scripts/quick_bench.py --model $m --batch $B --rounds $r --no-fp32 --reps 6 --horizon 100 2>&1 | grep "solves/s" | head -1
  done
done; done
