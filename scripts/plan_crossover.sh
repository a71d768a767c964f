#!/bin/bash
# One batch alone: latency sequence (line_search_rounds 1) against throughput sequence (2), default
# keep flags — the measurement behind throughput_sequence() in csrc/cabi.cu.
for m in mpc_time lateral mpc; do
  for B in 2048 4096 8192 16384; do
    for r in 1 2; do
      echo -n "$m B=$B rounds=$r: "
      python scripts/quick_bench.py --model $m --batch $B --rounds $r --no-fp32 --reps 6 --horizon 100 2>&1 | grep "solves/s" | head -1
    done
  done
done
