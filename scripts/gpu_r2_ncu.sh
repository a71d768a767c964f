#!/bin/bash
# round-2 ncu captures: metric table of one update at 32768 problems, --set full of the sweep, the
# round-1 rollout and the single-launch kernel, and the launch list of a short bench run
mkdir -p gpurun_out
M=$(python - <<'PY'
import sys; sys.path.insert(0, "scripts")
import summarize_ncu as s
print(",".join(k for k, _ in s.KEEP))
PY
)
ncu --profile-from-start off --metrics $M --clock-control none --csv --page raw --log-file gpurun_out/r2_ncu_32768.csv \
    python scripts/ncu_update.py --batch 32768 --iters 2 --rounds 2 > gpurun_out/r2_ncu_a.log 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:sweep_kernel --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r2_sweep -f python scripts/ncu_update.py --batch 32768 --iters 2 --rounds 2 > gpurun_out/r2_ncu_b.log 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:rollout_kernel --launch-skip 2 --launch-count 1 \
    -o gpurun_out/r2_rollout -f python scripts/ncu_update.py --batch 32768 --iters 2 --rounds 2 > gpurun_out/r2_ncu_c.log 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:solo_update_kernel --launch-count 1 \
    -o gpurun_out/r2_solo -f python scripts/ncu_update.py --batch 1 --iters 10 --single-launch 1 > gpurun_out/r2_ncu_d.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 8 --warmup 3 --in-flight 4 --skip-configs > gpurun_out/r2_ncu_e.log 2>&1
ls -la gpurun_out/ | tail -12
tail -3 gpurun_out/r2_ncu_a.log gpurun_out/r2_ncu_d.log gpurun_out/r2_ncu_e.log
