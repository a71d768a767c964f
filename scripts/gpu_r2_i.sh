#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2i.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -3 gpurun_out/r2i_pytest.log >> $L
for v in variants_c.so "" variants_c.so ""; do
  export TPLB_LIB_OVERRIDE=${v:+/root/repo/$v}
  [ -z "$v" ] && unset TPLB_LIB_OVERRIDE
  echo "=== variant ${v:-current(pipelined loads)}" >> $L
  timeout 300 python scripts/pipe_bench.py --in-flight 24 --graph 1 --steps 144 --tag "${v:-new}" >> $L 2>&1
  timeout 300 python scripts/quick_bench.py --batch 32768 --rounds 2 --keep-previous 0 --keep-records 0 --no-fp32 --reps 3 >> $L 2>&1
done
grep -E "===|PIPE|backward |rc=|passed|failed" $L
