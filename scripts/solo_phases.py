"""Development probe: cycles per phase of the single-launch kernel (needs a library built with
-DTPLB_SOLO_TIMING, path in argv[1])."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tpl_b200 import scenarios as sc
from tpl_b200.batched import BatchedOptim
lib = sys.argv[1]
gen, kw = (sc.mpc, dict(batch=1, horizon=60, max_iterations=20, forced=False)) if "mpc.so" in lib else \
          (sc.mpc_time, dict(batch=1, horizon=100, max_iterations=10, forced=True))
pb = gen(**kw)
q = sc.apply_to_batched(BatchedOptim(lib, batch=1, horizon_max=pb.horizon), pb)
q.single_launch = 1
x0, u0 = q._x[0].clone(), q._u.clone()
h = q._lib
h.tplb_debug_solo_cycles.argtypes = [C.POINTER(C.c_longlong)]
buf = (C.c_longlong * 8)()
for i in range(5):
    q._x[0].copy_(x0); q._u.copy_(u0); q.mu = 0.0; q.mu_step = 0
    torch.cuda.synchronize(); q.update(); torch.cuda.synchronize()
    if i == 0: h.tplb_debug_solo_cycles(buf)
h.tplb_debug_solo_cycles(buf)
names = ["load+init rollout+cost", "multiplier+linearize", "backward", "rollout", "stage costs", "sums+select+accept", "write back"]
tot = sum(buf[:7])
print(f"runtime {q.runtime:.3f} ms; cycles over 4 updates: total {tot}")
for n, c in zip(names, buf[:7]):
    print(f"  {n:28s} {c/4:12.0f} cycles/update  {c/4/1.965e3:8.1f} us  {c/tot:6.1%}")
