#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2h.log
: > $L
for v in variants_base.so "" variants_c.so variants_base.so "" variants_c.so; do
  export TPLB_LIB_OVERRIDE=${v:+/root/repo/$v}
  [ -z "$v" ] && unset TPLB_LIB_OVERRIDE
  echo "=== variant ${v:-current(A)}" >> $L
  timeout 300 python scripts/pipe_bench.py --in-flight 24 --graph 1 --steps 144 --tag "${v:-A}" >> $L 2>&1
  timeout 300 python scripts/quick_bench.py --batch 32768 --rounds 2 --keep-previous 0 --keep-records 0 --no-fp32 --reps 3 >> $L 2>&1
done
grep -E "===|PIPE|backward " $L
