#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2g.log
: > $L
echo "--- 3 steps" >> $L; python scripts/ulp_check.py tpl_b200/lib/libtplb200_lateral_profile.so >> $L 2>&1
echo "--- 2 steps" >> $L; python scripts/ulp_check.py /root/repo/variants_n2_lateral_profile.so >> $L 2>&1
for v in "" variants_n2_trajectory_tracking_mpc_time.so "" variants_n2_trajectory_tracking_mpc_time.so; do
  export TPLB_LIB_OVERRIDE=${v:+/root/repo/$v}
  [ -z "$v" ] && unset TPLB_LIB_OVERRIDE
  echo "=== variant ${v:-current}" >> $L
  timeout 300 python scripts/pipe_bench.py --in-flight 24 --graph 1 --steps 144 --tag "${v:-current}" >> $L 2>&1
done
grep -E "===|PIPE|ulp|---" $L
