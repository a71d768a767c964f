"""Development timing probe (not the contract bench): one workload, CUDA-event timing."""
import argparse, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tpl_b200 import build, scenarios as sc
from tpl_b200.batched import BatchedOptim
import copy

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="mpc_time")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--horizon", type=int, default=100)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--rounds", type=int, default=0)
ap.add_argument("--keep-previous", type=int, default=1)
ap.add_argument("--keep-records", type=int, default=1)
ap.add_argument("--no-fp32", action="store_true")
a = ap.parse_args()
t0 = time.time()
pb = getattr(sc, a.model)(batch=a.batch, horizon=a.horizon, max_iterations=a.iters, forced=True)
print(f"generated {a.batch} problems in {time.time()-t0:.1f}s")
libpath = os.environ.get("TPLB_LIB_OVERRIDE") or build.zoo_library_path(pb.model)
q0 = sc.apply_to_batched(BatchedOptim(libpath, batch=pb.batch, horizon_max=pb.horizon), pb)
q0.line_search_rounds = a.rounds
q0.keep_previous = bool(a.keep_previous)
q0.keep_records = bool(a.keep_records)
print("fp64 peak TFLOP/s:", q0.measure_fp64_tflops())
x0 = q0._x.clone(); u0 = q0._u.clone()
st = {k: v.clone() for k, v in q0._status.items()}
times = []
for r in range(a.reps):
    q0._x.copy_(x0); q0._u.copy_(u0)
    for k, v in st.items(): q0._status[k].copy_(v)
    torch.cuda.synchronize()
    q0.update()
    times.append(q0.runtime)
print("ms per update:", [f"{t:.3f}" for t in times])
best = min(times)
print(f"{a.model} B={a.batch} N={a.horizon} it={a.iters}: {best:.3f} ms -> {a.batch/best*1e3:.3e} solves/s")
print("iterations", q0.iterations[:8].tolist(), "cost", q0.traj_costs[:4].tolist())
q0._x.copy_(x0); q0._u.copy_(u0)
for k, v in st.items(): q0._status[k].copy_(v)
prof = q0.update_profiled()
tot = sum(ms for ms, _ in prof.values())
print("profiled (sync after every launch): total %.3f ms" % tot)
for k, (ms, n) in prof.items():
    print(f"  {k:14s} {ms:8.3f} ms  {n:3d} launches  {ms/max(n,1)*1e3:8.1f} us/launch  {ms/tot:6.1%}")
lin, bwd, roll = q0.work_counters()
print("work per solve: lin %.2f bwd %.2f rollouts %.2f" % (lin.float().mean(), bwd.float().mean(), roll.float().mean()))
if a.model == "mpc_time" and not a.no_fp32:
    q0.precision = "fp32"
    ts = []
    for r in range(4):
        q0._x.copy_(x0); q0._u.copy_(u0)
        for k, v in st.items(): q0._status[k].copy_(v)
        torch.cuda.synchronize(); q0.update(); ts.append(q0.runtime)
    print(f"fp32 mode: {min(ts):.3f} ms -> {a.batch/min(ts)*1e3:.3e} solves/s")
    q0.precision = "fp64"
