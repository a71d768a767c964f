"""BASELINE.json configs[0]: one solve per planning / control cycle (batch = 1) for the three
solvers the default tplsim scenario runs — latency of the B200 path next to the reference's own
CPU solver on this host.  Prints a markdown table (profiles/rNN_latency_config1.md)."""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tpl_b200 import build, scenarios as sc
from tpl_b200.batched import BatchedOptim
from oracle import ref, oracle

CASES = [
    ("lateral_profile (PathOptim, N=250, max_iterations=5, EULER)", sc.lateral, dict(batch=1, horizon=250, max_iterations=5, forced=False)),
    ("velocity_profile_space (VelocityOptim, N=250, max_iterations=20, EULER)", sc.velocity, dict(batch=1, horizon=250, max_iterations=20, forced=False)),
    ("trajectory_tracking_mpc (MPC, N=60, max_iterations=20, HEUN)", sc.mpc, dict(batch=1, horizon=60, max_iterations=20, forced=False)),
    ("trajectory_tracking_mpc_time (MPC time, N=40, max_iterations=20, HEUN)", sc.mpc_time, dict(batch=1, horizon=40, max_iterations=20, forced=False)),
    ("trajectory_tracking_mpc_time (bench shape: N=100, 10 forced iterations, HEUN)", sc.mpc_time, dict(batch=1, horizon=100, max_iterations=10, forced=True)),
]
print("| solver (caller settings) | iterations | B200 one launch p50 ms | B200 launch sequence p50 ms | reference CPU p50 ms (1 core) | CPU kind |")
print("|---|---:|---:|---:|---:|---|")
for label, gen, kw in CASES:
    pb = gen(**kw)
    res = {}
    for mode in (1, -1):
        q = sc.apply_to_batched(BatchedOptim(build.zoo_library_path(pb.model), batch=1, horizon_max=pb.horizon), pb)
        q.single_launch = mode
        x0, u0 = q._x[0].clone(), q._u.clone()
        ts = []
        for i in range(110):
            q._x[0].copy_(x0); q._u.copy_(u0); q.mu = 0.0; q.mu_step = 0
            if q.C: q.lagrange_multiplier = 0.0
            torch.cuda.synchronize(); q.update()
            if i >= 10: ts.append(q.runtime)
        res[mode] = float(np.median(ts))
    Ref = ref.load(pb.model, "fast")
    kind = "reference"
    if Ref is None:
        oracle.build_libs(); Ref = lambda: oracle.OracleOptim(pb.model); kind = "port"
    base = sc.apply_to_single(Ref(), pb, 0)
    rt = []
    for _ in range(50):
        o = copy.deepcopy(base); t0 = time.perf_counter(); o.update(); rt.append((time.perf_counter() - t0) * 1e3)
    print(f"| {label} | {int(q.iterations[0])} (ref {int(o.iterations)}) | {res[1]:.3f} | {res[-1]:.3f} | {np.median(rt):.3f} | {kind} |")
