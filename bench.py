#!/usr/bin/env python
"""Headline benchmark: batched iLQR solves/sec on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): 4096 independent kinematic-bicycle MPC problems
(`trajectory_tracking_mpc_time`, X=6, U=2, C=4), N=100 stages, HEUN, 10 forced
iLQR iterations, fp64, synthetic road / reference-path data seeded per problem.
A *step* is one `update()` of one batch of 4096 from the same cold start.  `--in-flight D`
(default 8) solver instances, each on its own CUDA stream, take the steps round-robin
(tpl_b200.streaming): one batch alone cannot fill the GPU, consecutive batches overlap.
With N GPUs every rank solves its own batches (weak scaling, no data-path collective)
and every step ends with an all_gather of the per-group best costs.

Prints ONE JSON line (see the driver contract): `value` is device-resident throughput,
`e2e` the same through the public API with pinned host buffers (uploads and downloads of
every step inside the timed region), `roofline` the HBM fraction of the dominant kernel
(algorithmic bytes / launch duration, timed with the same number of problems in one
batch because kernels of different streams overlap), `roofline_step` the same for the
whole step, `roofline_fp64` the algorithmic FP64 fraction, `cpu_baseline` the reference's
own CPU solver on this host, `latency` the single-solve p50, `profile_shaping` row f2.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL = "trajectory_tracking_mpc_time"
METRIC = "batched iLQR solves/sec (N=100, fp64)"
UNIT = "solves/s"

# Algorithmic flops per stage (add/sub/mul/div = 1,
# interpolation call = 9; transcendental calls are listed separately as "special").
FLOPS = {
    # fwd = feedback (X+2U+2UX = 34) + HEUN (2 F_ct + 5X = 70) + stage cost (84 + 4 lookups x 9 + 1 = 121)
    "trajectory_tracking_mpc_time": dict(lin=366, bwd=1717, fwd=225, fwd_chain=104, fwd_cost=121, con=9,
                                         sp_lin=22, sp_fwd=12),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="problems per GPU")
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--line-search-rounds", type=int, default=2, choices=[0, 1, 2],
                    help="tplb_batch.line_search_rounds for the throughput legs (2: least work, for a full GPU)")
    ap.add_argument("--in-flight", type=int, default=8,
                    help="batches in flight: solver instances (one CUDA stream each) the steps alternate between")
    return ap.parse_args()


# ---------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# ---------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.marks = index, [], None, []

    def mark(self):
        """Bracket the timed region: samples outside [first mark, last mark + one period] are dropped."""
        self.marks.append(time.time())

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lo = self.marks[0] if self.marks else 0.0
        hi = (self.marks[-1] + 0.05) if len(self.marks) > 1 else float("inf")
        inside = [r for ts, r in self.rows if lo <= ts <= hi] or [r for _, r in self.rows[-3:]]
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------
# CPU arm: the reference's own solver on the host cores
# ---------------------------------------------------------------------------------
def cpu_solver_factory(model):
    """(factory, kind): the real reference build if oracle/_ref travelled here,
    else the C restatement."""
    from oracle import ref
    Ref = ref.load(model, "fast")
    if Ref is not None:
        return Ref, "reference"
    from oracle import oracle
    oracle.build_libs()
    return (lambda: oracle.OracleOptim(model)), "port"


def cpu_throughput(pb, n, threads, repeats=1):
    """solves/s of `n` problems of `pb` on `threads` host threads.  Objects are
    prepared beforehand; only update() (GIL released, optim.c:1487) is timed."""
    import copy
    from concurrent.futures import ThreadPoolExecutor
    from tpl_b200 import scenarios as sc
    factory, kind = cpu_solver_factory(pb.model)
    n = min(n, pb.batch)
    objs = [sc.apply_to_single(factory(), pb, i) for i in range(n)]
    best = None
    chunks = [list(range(i, n, threads)) for i in range(threads)]
    for _ in range(repeats):
        work = [copy.deepcopy(o) for o in objs]

        def run(idx):
            for i in idx:
                work[i].update()
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(run, chunks))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n / best, kind, n


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tpl_b200 import scenarios as sc
    cores = os.cpu_count() or 1
    n = max(cores * 32, 512)
    pb = sc.mpc_time(batch=n, horizon=a.horizon, max_iterations=a.iterations, forced=True)
    for _ in range(a.warmup):
        cpu_throughput(pb, max(cores, 16), cores)
    t0 = time.perf_counter()
    vals = []
    for _ in range(a.steps):
        v, kind, used = cpu_throughput(pb, n, cores)
        vals.append(v)
    elapsed = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * n / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{n}-problem sample per step of: 4096 independent {MODEL} problems, "
                               f"N={a.horizon}, {a.iterations} forced iLQR iterations, HEUN, fp64",
                   "flush": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{n} problems per step x {a.steps} steps, one Optim object per problem, "
                                   f"{cores} threads, update() only"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": elapsed,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from tpl_b200 import build, dist as tdist, scenarios as sc
    from tpl_b200.batched import BatchedOptim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; stdout is
        # reserved for the one JSON line, so fd 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    B, T, I = a.batch, a.horizon, a.iterations
    lib = build.zoo_library_path(MODEL)
    if not os.path.exists(lib):
        build.build_zoo([MODEL])
    pb = sc.mpc_time(batch=B, horizon=T, max_iterations=I, forced=True, seed0=rank * B)
    # `in_flight` solver instances, each on its own CUDA stream, take the batches round-robin
    # (tpl_b200.streaming): one batch of 4096 leaves most SMs idle in the two serial phases, so
    # consecutive batches overlap; with host buffers their copies overlap the kernels as well.
    from tpl_b200.streaming import SolverPipeline
    depth = max(1, a.in_flight)
    def make_solver():
        o = sc.apply_to_batched(BatchedOptim(lib, batch=B, horizon_max=T), pb)
        o.line_search_rounds = a.line_search_rounds
        return o

    pipe = SolverPipeline(make_solver, depth=depth)
    opt = pipe.slots[0].opt
    group = 64                                           # problems per argmin group (multi-start style)

    x0 = opt._x[0].clone()
    u0 = opt._u.clone()
    torch.cuda.synchronize()                             # set-up ran on the default stream

    def reset(o=None):
        o = opt if o is None else o
        o._x[0].copy_(x0)
        o._u.copy_(u0)
        o.lagrange_multiplier = 0.0                      # a fresh problem, as in the reference arm
        o.mu = 0.0
        o.mu_step = 0

    def step():
        with pipe.next() as slot:
            reset(slot.opt)
            slot.opt.update()
            if world > 1:
                mn, am = slot.opt.argmin_groups(group)
                tdist.gather_best(mn, am, rank * B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the sampler starts before the warm-up (nvidia-smi needs ~0.1 s to deliver its first
    # line); only samples taken between the two barriers of the timed region are kept
    with ClockSampler(local) as clk:
        for _ in range(max(a.warmup, depth)):
            step()
        pipe.join()
        barrier()
        clk.mark()
        e0.record()
        pipe.fork()
        for _ in range(a.steps):
            step()
        pipe.join()
        e1.record()
        barrier()
        clk.mark()
        time.sleep(0.05)
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * B * a.steps / (elapsed_ms * 1e-3)
    resident_cost_check = float(opt.traj_costs.sum())

    # ---- end to end through the public API with host buffers --------------------------
    names = opt.params.scalar_names
    host = {
        "x0": torch.from_numpy(pb.x0).pin_memory(),
        "u0": torch.from_numpy(pb.u0).pin_memory(),
        "arrays": {k: torch.from_numpy(v).pin_memory() for k, v in pb.arrays.items()},
        # all scalar parameters as one (num_scalars, S) block in the solver's order
        "scalars": torch.from_numpy(np.stack([np.broadcast_to(np.asarray(pb.scalars[k], dtype=np.float64),
                                                              (opt.scenes,)) for k in names])).pin_memory(),
    }
    # same pipeline; every step uploads all of its inputs and downloads all of its results
    # inside the timed region
    outs = [{
        "x": torch.empty((B, T + 1, opt.X), dtype=torch.float64).pin_memory(),
        "u": torch.empty((B, T, opt.U), dtype=torch.float64).pin_memory(),
        "c": torch.empty((B,), dtype=torch.float64).pin_memory(),
        "i": torch.empty((2, B), dtype=torch.int32).pin_memory(),
    } for _ in range(depth)]
    h2d = (host["x0"].numel() + host["u0"].numel() + host["scalars"].numel()) * 8 \
        + sum(v.numel() * 8 for v in host["arrays"].values())
    d2h = (outs[0]["x"].numel() + outs[0]["u"].numel() + outs[0]["c"].numel()) * 8 + outs[0]["i"].numel() * 4

    def step_e2e():
        with pipe.next() as slot:
            o, out = slot.opt, outs[slot.index]
            o.params.set_scalars(host["scalars"])
            for k, v in host["arrays"].items():
                setattr(o.params, k, v)
            o.set_initial_state(host["x0"])
            o.u = host["u0"]
            o.lagrange_multiplier = 0.0
            o.mu = 0.0
            o.mu_step = 0
            o.update()
            out["x"].copy_(o.x, non_blocking=True)
            out["u"].copy_(o.u, non_blocking=True)
            out["c"].copy_(o.traj_costs, non_blocking=True)
            out["i"][0].copy_(o.iterations, non_blocking=True)
            out["i"][1].copy_(o.termination_condition, non_blocking=True)
            if world > 1:
                mn, am = o.argmin_groups(group)
                tdist.gather_best(mn, am, rank * B)

    for _ in range(max(depth, a.warmup // 2)):
        step_e2e()
    pipe.join()
    barrier()
    e0.record()
    pipe.fork()
    for _ in range(a.steps):
        step_e2e()
    pipe.join()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    # the downloaded results are the solver's: same costs as the device-resident run
    pipe.slots[0].wait()
    e2e_cost_check = float(outs[0]["c"].sum())
    if abs(e2e_cost_check - resident_cost_check) > 1e-9 * abs(resident_cost_check):
        raise SystemExit(f"end-to-end results differ from the device-resident run: "
                         f"{e2e_cost_check!r} vs {resident_cost_check!r}")
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B * a.steps / (e2e_ms * 1e-3)

    # ---- attribution: time per kernel class and algorithmic work (rank 0) ----------------
    line = None
    if rank == 0:
        reset()
        prof = opt.update_profiled()
        torch.cuda.synchronize()
        lin, bwd, roll = (int(c.sum().item()) for c in opt.work_counters())
        F = FLOPS[MODEL]
        # algorithmic flops by kernel class; the forward pass (F_fwd per stage and rollout the
        # reference would run sequentially) splits into the dynamics chain and the stage cost
        flops = {
            "linearize": lin * T * F["lin"],
            "backward": bwd * T * F["bwd"],
            "rollout": (roll - B) * T * F["fwd_chain"],
            "stage_cost": (roll - B) * T * F["fwd_cost"],
            "rollout_init": B * T * F["fwd"],
            "multiplier": B * T * F["con"] * pb.max_lg_iterations,
        }
        flops_step = sum(flops.values())
        launches = sum(c for _, c in prof.values())
        total_prof_ms = sum(ms for ms, _ in prof.values())
        peak = opt.measure_fp64_tflops()
        ms_per_step = elapsed_ms / a.steps
        step_tf = flops_step / (ms_per_step * 1e-3) / 1e12
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fd:
                hbm_peak = json.load(fd)["hbm_gbs"]
            hbm_src = "MEASURED_PEAKS.json hbm_gbs (copy bandwidth, burst)"
        except (OSError, KeyError, ValueError):
            hbm_peak = 6650.0
            hbm_src = "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"

        # ---- roofline: HBM -------------------------------------------------------------
        # With the GPU full (`depth` batches in flight) every hot kernel streams its operands
        # from HBM: the working set of ONE batch (derivative records, K, k, x, u, candidates)
        # is already larger than L2.  Kernels of different streams overlap, so they cannot be
        # timed one by one inside the timed region; their durations are taken from the same
        # number of problems (depth x B) in one batch, one launch at a time (CUDA events on the
        # launching stream).  Algorithmic bytes per problem and stage:
        info = opt._info
        X_, U_, C_, REC = opt.X, opt.U, opt.C, info["deriv_compact"]
        R1 = 2                                           # candidates of the first line-search round
        two_round = a.line_search_rounds == 2 or (a.line_search_rounds == 0 and B * depth >= 16384)
        # pure penalty (lg_mult_limit = 0, the shipped callers' setting): the multipliers are
        # identically 0 after the multiplier update and are not read again inside update()
        LAM = 0 if float(np.max(np.abs(np.asarray(pb.lg_mult_limit)))) == 0.0 else C_
        step_copy = 4 * (X_ + U_)        # accept: winner -> x,u and x,k -> prev_x,prev_k (read + write)
        abytes = {                       # doubles moved per (problem, stage, launch)
            # x, u, multipliers in, derivative record out (+ the accept when it is folded in)
            "linearize": (X_ + U_) + LAM + REC if two_round else step_copy + LAM + REC,
            "accept": step_copy,
            # record + u + bounds in, K and k out
            "backward": REC + 3 * U_ + U_ * X_ + U_,
            # u, k, bounds, K, x in; R1 candidates out
            "rollout": 4 * U_ + U_ * X_ + X_ + R1 * (X_ + U_),
            # R1 candidates + multipliers in, R1 terms out
            "stage_cost": R1 * (X_ + U_) + C_ + R1,
        }
        if two_round:                    # the rollouts add up the stage costs themselves (+ multipliers in)
            abytes["rollout"] += LAM
            abytes["stage_cost"] = 0
        sat = None
        Bs = B * depth
        if depth > 1 and Bs <= 131072:
            pbs = sc.mpc_time(batch=Bs, horizon=T, max_iterations=I, forced=True, seed0=rank * B)
            big = sc.apply_to_batched(BatchedOptim(lib, batch=Bs, horizon_max=T), pbs)
            big.line_search_rounds = a.line_search_rounds
            big.update()                                  # warm-up
            sc.apply_to_batched(big, pbs)                 # fresh problems again
            big.mu, big.mu_step = 0.0, 0
            torch.cuda.synchronize()
            sat = big.update_profiled()
            torch.cuda.synchronize()
            del big
            torch.cuda.empty_cache()
        src_prof, src_B = (sat, Bs) if sat is not None else (prof, B)
        hot = {k: v for k, v in src_prof.items() if k in abytes and v[0] > 0}
        dom = max(hot, key=lambda k: hot[k][0])
        dom_ms, dom_n = hot[dom]
        # rollout / stage_cost: round 2 touches only the few pending problems, so the bytes are
        # those of round 1 while the time is that of both rounds.  linearize: the first launch of
        # an update has no step to accept.
        iters = I * pb.max_lg_iterations
        doubles_update = {k: v * iters for k, v in abytes.items()}
        if not two_round:                # folded: the first linearize of an update has nothing to accept,
            doubles_update["linearize"] -= 3 * (X_ + U_) * pb.max_lg_iterations
            doubles_update["accept"] = step_copy * pb.max_lg_iterations   # ... the last step has its own launch
        gbs = {k: doubles_update[k] * 8.0 * src_B * T / (ms * 1e-3) / 1e9 for k, (ms, n) in hot.items()}
        bytes_solve = sum(doubles_update[k] for k in hot) * 8.0 * T
        step_gbs = bytes_solve * value / world / 1e9
        # DRAM bytes of the hot kernels from the committed ncu capture (profiles/ncu_traffic.json),
        # per problem and stage, scaled to this launch
        ncu = {}
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fd:
                ncu = json.load(fd)["kernels"]
        except (OSError, KeyError, ValueError):
            ncu = {}
        traffic = None
        if dom in ncu and (T, MODEL) == (100, "trajectory_tracking_mpc_time"):
            traffic = ncu[dom]["dram_bytes_per_problem_stage"] * src_B * T
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{B} independent {MODEL} problems per GPU (X=6,U=2,C=4), N={T}, "
                            f"{I} forced iLQR iterations, HEUN, fp64 (BASELINE.json configs[1])",
                "problems_per_gpu": B, "stages": T, "iterations": I, "batches_in_flight": depth,
                "line_search_rounds": a.line_search_rounds,
                "flush": "working set per step (derivative blocks + 8 line-search candidates, "
                         f"{opt._workspace_bytes / 1e6:.0f} MB) exceeds the 126 MB L2; no explicit flush",
                "final_collective": "all_gather of per-group (min cost, argmin)" if world > 1 else "none (1 GPU)",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / a.steps, "batches_in_flight": depth,
                    "cost_checksum": e2e_cost_check},
            "gpu_launches": launches * a.steps,
            "gpu_launches_per_step": launches,
            "roofline": {
                "bound": "hbm", "kernel": dom, "achieved": gbs[dom], "peak": hbm_peak, "unit": "GB/s",
                "frac": gbs[dom] / hbm_peak, "traffic": traffic,
                "peak_source": hbm_src,
                "algorithmic_bytes_per_launch": doubles_update[dom] * 8.0 * src_B * T / iters,
                "avg_launch_ms": dom_ms / iters,
                "timed_on": f"{src_B} problems in one batch (the {depth} x {B} in flight overlap and cannot be "
                            "timed per kernel), CUDA events around every launch",
                "all_kernels_gbs": {k: round(v, 1) for k, v in gbs.items()},
                "all_kernels_frac": {k: round(v / hbm_peak, 4) for k, v in gbs.items()},
            },
            "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": hbm_peak, "unit": "GB/s",
                              "frac": step_gbs / hbm_peak,
                              "algorithmic_bytes_per_solve": bytes_solve},
            "roofline_fp64": {"achieved": step_tf, "peak": peak, "unit": "TFLOP/s",
                              "frac": step_tf / peak if peak > 0 else None,
                              "algorithmic_flops_per_solve": flops_step / B,
                              "peak_source": "DFMA loop measured live by tplb_measure_fp64_tflops",
                              "special_function_calls_excluded": True},
            "kernel_ms_saturated": ({k: round(ms, 4) for k, (ms, _) in sat.items()} if sat else None),
            "kernel_ms": {k: round(ms, 4) for k, (ms, _) in prof.items()},
            "kernel_share": {k: round(ms / total_prof_ms, 4) for k, (ms, _) in prof.items()},
            "work_per_solve": {"linearisations": lin / B, "backward_sweeps": bwd / B, "rollouts": roll / B},
            "hbm": {"peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json"},
            "clocks": clk.summary(),
        }
        if world == 1:
            cores = os.cpu_count() or 1
            cpu_v, kind, used = cpu_throughput(pb, a.cpu_sample, cores)
            line["cpu_baseline"] = {
                "value": cpu_v, "unit": UNIT, "cores": cores, "kind": kind,
                "sample": f"first {used} problems of the same batch, one Optim object per problem, "
                          f"{cores} threads, update() only (GIL released)"}
            # second half of the metric: p50 latency of ONE solve (batch = 1), same problem shape
            line["latency"] = single_solve_latency(lib, pb, cpu=True)
            # the same kernels once the batch fills the chip (not the headline workload)
            line["large_batch"] = large_batch_throughput(lib, a, 65536)
            # row f2: the profile shaping that precedes the lateral / velocity solves
            line["profile_shaping"] = profile_shaping()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def single_solve_latency(lib, pb, reps=200, cpu=True):
    """p50 of one update() with batch = 1 (CUDA events around the call, parameters
    resident), next to the reference's own `opt.runtime` for the same problem."""
    import copy
    import numpy as np
    import torch
    from tpl_b200 import scenarios as sc
    from tpl_b200.batched import BatchedOptim
    one = pb.subset([0])
    opt = sc.apply_to_batched(BatchedOptim(lib, batch=1, horizon_max=one.horizon), one)
    x0, u0 = opt._x[0].clone(), opt._u.clone()
    times = []
    for i in range(reps + 10):
        opt._x[0].copy_(x0); opt._u.copy_(u0); opt.mu = 0.0; opt.mu_step = 0
        torch.cuda.synchronize()
        opt.update()
        if i >= 10:
            times.append(opt.runtime)
    out = {"gpu_p50_ms": float(np.median(times)), "gpu_p90_ms": float(np.percentile(times, 90)),
           "reps": reps, "what": "batch=1, one update() = 10 forced iterations, N=100, CUDA events around the call"}
    if cpu:
        factory, kind = cpu_solver_factory(pb.model)
        base = sc.apply_to_single(factory(), pb, 0)
        rt = []
        for _ in range(50):
            q = copy.deepcopy(base)
            t0 = time.perf_counter()
            q.update()
            rt.append((time.perf_counter() - t0) * 1e3)
        out["cpu_p50_ms"] = float(np.median(rt))
        out["cpu_kind"] = kind
    return out


def profile_shaping(batch=16384, reps=20, cpu_sample=256):
    """Row f2 (tpl_b200.prep): the two rampify kernels at config #4's batch size, device-resident
    inputs, CUDA events; beside them the C restatement (oracle) on one host core."""
    import numpy as np
    import torch
    from oracle import prep as oprep
    from tpl_b200 import prep, prep_scenarios as ps

    dev = torch.device("cuda", torch.cuda.current_device())
    out = {}
    vb = ps.velocity_batch(64, n=250)
    tile = lambda x: torch.as_tensor(np.resize(x, (batch,) + x.shape[1:]), device=dev)   # noqa: E731
    lim, v0, a0 = tile(vb["lim_v"]), tile(vb["v0"]), tile(vb["a0"])
    args = (vb["a_min"], vb["a_max"], vb["j_min"], vb["j_max"], vb["v_min"], vb["step"])
    lb = ps.lateral_batch(64, n=200)
    pv, lo, up, pj = tile(lb["path_v"]), tile(lb["lower"]), tile(lb["upper"]), tile(lb["proj_distance"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms = timed(lambda: prep.rampify_velocity_profile(v0, a0, lim, *args))
    t0 = time.perf_counter()
    for k in range(cpu_sample):
        oprep.rampify_velocity(vb["v0"][k % 64], vb["a0"][k % 64], vb["lim_v"][k % 64], *args)
    cpu = cpu_sample / (time.perf_counter() - t0)
    out["velocity"] = {"problems": batch, "samples": 250, "ms": ms, "profiles_per_s": batch / (ms * 1e-3),
                       "cpu_port_profiles_per_s_1core": cpu,
                       "what": "includes the (B,N)->[N][B] transposes of the wrapper"}
    ms = timed(lambda: prep.rampify_lateral_profile(lb["step"], 200, lb["evasion_sharpness"], pj, pv, lb["gap"], lo, up))
    path = np.zeros((200, 6))
    t0 = time.perf_counter()
    for k in range(cpu_sample):
        path[:, 5] = lb["path_v"][k % 64]
        oprep.rampify_lateral(lb["step"], 200, lb["evasion_sharpness"], lb["proj_distance"][k % 64], path, lb["gap"],
                              lb["lower"][k % 64], lb["upper"][k % 64])
    cpu = cpu_sample / (time.perf_counter() - t0)
    out["lateral"] = {"problems": batch, "samples": 200, "ms": ms, "profiles_per_s": batch / (ms * 1e-3),
                      "cpu_port_profiles_per_s_1core": cpu,
                      "divisions_per_profile": 2 * 200 * 201 // 2 + 200}
    return out


def large_batch_throughput(lib, a, batch):
    import torch
    from tpl_b200 import scenarios as sc
    from tpl_b200.batched import BatchedOptim
    pb = sc.mpc_time(batch=batch, horizon=a.horizon, max_iterations=a.iterations, forced=True)
    opt = sc.apply_to_batched(BatchedOptim(lib, batch=batch, horizon_max=a.horizon), pb)
    x0, u0 = opt._x[0].clone(), opt._u.clone()
    best = None
    for i in range(4):
        opt._x[0].copy_(x0); opt._u.copy_(u0); opt.mu = 0.0; opt.mu_step = 0
        torch.cuda.synchronize()
        opt.update()
        if i:
            best = opt.runtime if best is None else min(best, opt.runtime)
    return {"problems": batch, "value": batch / (best * 1e-3), "unit": UNIT, "ms_per_step": best}


def main():
    a = parse()
    world_env = int(os.environ.get("WORLD_SIZE", "0"))
    if a.gpus > 1 and world_env == 0:
        # convenience: re-launch one rank per GPU like the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
