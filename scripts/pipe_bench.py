"""Development probe: device-resident throughput of the headline workload with several
batches in flight (what bench.py times, without its CPU legs).  One line per run."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tpl_b200 import build, scenarios as sc
from tpl_b200.batched import BatchedOptim
from tpl_b200.streaming import SolverPipeline

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--horizon", type=int, default=100)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--in-flight", type=int, default=8)
ap.add_argument("--steps", type=int, default=48)
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--keep-previous", type=int, default=0)
ap.add_argument("--keep-records", type=int, default=0)
ap.add_argument("--tag", default="")
ap.add_argument("--graph", type=int, default=0, help="1: replay one CUDA graph per slot instead of re-enqueueing")
a = ap.parse_args()

lib = os.environ.get("TPLB_LIB_OVERRIDE") or build.zoo_library_path("trajectory_tracking_mpc_time")
pb = sc.mpc_time(batch=a.batch, horizon=a.horizon, max_iterations=a.iters, forced=True)

def make():
    o = sc.apply_to_batched(BatchedOptim(lib, batch=a.batch, horizon_max=a.horizon), pb)
    o.line_search_rounds = a.rounds
    o.keep_previous = bool(a.keep_previous)
    o.keep_records = bool(a.keep_records)
    return o

pipe = SolverPipeline(make, depth=a.in_flight)
o0 = pipe.slots[0].opt
x0, u0 = o0._x[0].clone(), o0._u.clone()
torch.cuda.synchronize()

def step():
    with pipe.next() as slot:
        o = slot.opt
        o._x[0].copy_(x0); o._u.copy_(u0)
        o.lagrange_multiplier = 0.0; o.mu = 0.0; o.mu_step = 0
        o.update()

for _ in range(max(3, a.in_flight)):
    step()
pipe.join(); torch.cuda.synchronize()

if a.graph:
    graphs = []
    for slot in pipe.slots:
        o = slot.opt
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=slot.stream):
            o._x[0].copy_(x0); o._u.copy_(u0)
            o.lagrange_multiplier = 0.0; o.mu = 0.0; o.mu_step = 0
            o.update()
        graphs.append(g)
    torch.cuda.synchronize()

    def step():
        with pipe.next() as slot:
            graphs[slot.index].replay()

import time
best = None
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pipe.fork()
    h0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    host_ms = (time.perf_counter() - h0) * 1e3
    pipe.join(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = ms if best is None else min(best, ms)
print(f"PIPE {a.tag} B={a.batch} D={a.in_flight} rounds={a.rounds} prev={a.keep_previous} rec={a.keep_records} "
      f"fused={'0' if os.environ.get('TPLB_NO_FUSED_SWEEP') else '1'}: {best / a.steps:.4f} ms/step "
      f"{a.batch * a.steps / best * 1e3:.4e} solves/s  host_enqueue={host_ms / a.steps:.4f} ms/step graph={a.graph} cost={float(o0.traj_costs.sum()):.9e}")
