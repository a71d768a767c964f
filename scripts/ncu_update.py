#!/usr/bin/env python
"""One update() of the bench workload between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --set full --clock-control none -o gpurun_out/x python scripts/ncu_update.py --batch 65536
(first update untimed/unprofiled as warm-up)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpl_b200 import build, scenarios as sc  # noqa: E402
from tpl_b200.batched import BatchedOptim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="trajectory_tracking_mpc_time")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--horizon", type=int, default=100)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--rounds", type=int, default=0)
ap.add_argument("--keep", type=int, default=0, help="keep_previous / keep_records")
ap.add_argument("--single-launch", type=int, default=-1)
a = ap.parse_args()

pb = sc.mpc_time(batch=a.batch, horizon=a.horizon, max_iterations=a.iters, forced=True)
opt = sc.apply_to_batched(BatchedOptim(build.zoo_library_path(a.model), batch=a.batch, horizon_max=a.horizon), pb)
opt.line_search_rounds = a.rounds
opt.keep_previous = opt.keep_records = bool(a.keep)
opt.single_launch = a.single_launch
x0, u0 = opt._x[0].clone(), opt._u.clone()
opt.update()
torch.cuda.synchronize()
opt._x[0].copy_(x0)
opt._u.copy_(u0)
opt.lagrange_multiplier = 0.0
opt.mu, opt.mu_step = 0.0, 0
torch.cuda.synchronize()
torch.cuda.profiler.start()
opt.update()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("cost", float(opt.traj_costs.sum()))
