#!/bin/bash
mkdir -p gpurun_out
M=$(python - <<'PY'
import sys; sys.path.insert(0, "scripts")
import summarize_ncu as s
print(",".join(k for k, _ in s.KEEP) + ",smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio")
PY
)
ncu --profile-from-start off --metrics $M --clock-control none --csv --page raw --log-file gpurun_out/r2_ncu_32768.csv \
    python scripts/ncu_update.py --batch 32768 --iters 2 --rounds 2 > gpurun_out/r2_ncu_a.log 2>&1
tail -2 gpurun_out/r2_ncu_a.log
