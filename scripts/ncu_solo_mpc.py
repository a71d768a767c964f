"""One single-launch update of the 7x2 model (batch 1) for `ncu --set full --import-source on`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tpl_b200 import build, scenarios as sc
from tpl_b200.batched import BatchedOptim
pb = sc.mpc(batch=1, horizon=60, max_iterations=20, forced=False)
q = sc.apply_to_batched(BatchedOptim(build.zoo_library_path(pb.model), batch=1, horizon_max=60), pb)
q.single_launch = 1
x0, u0 = q._x[0].clone(), q._u.clone()
q.update(); torch.cuda.synchronize()
q._x[0].copy_(x0); q._u.copy_(u0); q.mu = 0.0; q.mu_step = 0; q.lagrange_multiplier = 0.0
torch.cuda.synchronize()
torch.cuda.profiler.start(); q.update(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("iterations", int(q.iterations[0]), "ms", q.runtime)
