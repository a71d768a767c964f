"""Development probe: one update() of B problems through the single-launch kernel (one thread block
per problem) and through the batched launch sequences — where does the automatic choice flip?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tpl_b200 import build, scenarios as sc
from tpl_b200.batched import BatchedOptim

print("| model | B | one launch ms | launch sequences ms |\n|---|---:|---:|---:|")
for name, gen, kw in (("mpc_time N=100 it=10", sc.mpc_time, dict(horizon=100, max_iterations=10, forced=True)),
                      ("lateral N=250 it=5", sc.lateral, dict(horizon=250, max_iterations=5, forced=False))):
    for B in (1, 37, 148, 296, 444, 592, 888, 1184, 2368):
        pb = gen(batch=B, **kw)
        res = {}
        for mode in (1, -1):
            q = sc.apply_to_batched(BatchedOptim(build.zoo_library_path(pb.model), batch=B, horizon_max=pb.horizon), pb)
            q.single_launch = mode
            x0, u0 = q._x[0].clone(), q._u.clone()
            ts = []
            for i in range(8):
                q._x[0].copy_(x0); q._u.copy_(u0); q.mu = 0.0; q.mu_step = 0
                if q.C: q.lagrange_multiplier = 0.0
                torch.cuda.synchronize(); q.update()
                if i >= 2: ts.append(q.runtime)
            res[mode] = float(np.median(ts))
        print(f"| {name} | {B} | {res[1]:.3f} | {res[-1]:.3f} |")
