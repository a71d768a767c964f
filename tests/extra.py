"""Parity cases beyond the per-iteration traces of tests/common.py: the solver entry points and
corner semantics that round 1 compared with the C restatement only (`ilr`, `shift`, the point
evaluations incl. the negative-index wrap, sticky state, the two zoo models without a workload
generator, a user-defined problem with RK4 + augmented Lagrangian + end cost, prev_x / prev_k).

``run_single(make)`` drives any CPU solver with the reference's ``Optim`` interface
(``make(model)`` returns a fresh object) and returns ``{key: array}``;
``run_batched(make)`` produces the same keys from the CUDA solver (``make(model, batch,
horizon_max, scenes)`` returns a ``BatchedOptim``).  tests/golden/make_golden_extra.py records
``run_single`` of the REAL reference into tests/golden/ref_extra.npz.
"""

import numpy as np

from tpl_b200 import scenarios as sc

CUSTOM = "custom_unicycle"
TRACK = "custom_track"
S_R = (0.25, 0.0, -1e-9, -0.25, 100.0)
ZOO = (("ref_line_smoother_dk", 120), ("velocity_profile_time", 80))
STICKY = ((2, 1), (0, 1), (1, 1), (3, 0))      # (max_iterations, max_lg_iterations) of consecutive update() calls


def custom_definition(genopt, spx):
    """A unicycle that has to reach a target pose: lookup array, constraint, end cost (RK4 and
    two augmented-Lagrangian iterations are solver settings).  `genopt`, `spx`: the reference's
    modules or tpl_b200's — same definition for both."""
    import sympy as sp
    x, y, phi, v, w, t, dt = sp.symbols("x y phi v w t dt")
    wx, wu, r_max, ref_step = sp.symbols("wx wu r_max ref_step")
    lane = spx.ArraySymbol("lane")
    y_ref = spx.lerp(0.0, ref_step, t * dt, lane)
    return genopt.Config(
        [x, y, phi], [v, w], {wx: 2.0, wu: 0.1, r_max: 4.0, ref_step: 0.1, lane: None},
        sp.Matrix([v * sp.cos(phi), v * sp.sin(phi), w]),
        wx * ((x - 3.0)**2 + (y - y_ref)**2) + wu * (v**2 + w**2),
        end_costs=10.0 * ((x - 3.0)**2 + (y - y_ref)**2 + phi**2),
        constraints=[x**2 + y**2 - r_max**2])


def custom_inputs():
    B, T = 16, 50
    rng = np.random.default_rng(3)
    lanes = 0.5 * np.sin(np.linspace(0, 3, 60))[None, :] + rng.normal(0, 0.05, (B, 1))
    x0 = rng.normal(0.0, 0.3, (B, 3))
    return B, T, lanes, x0


def configure_custom(o, T, lane):
    o.horizon = T; o.step = 0.1; o.integrator_type = o.RK4
    o.max_iterations = 15; o.max_lg_iterations = 2
    o.barrier_weight = 50.0; o.lg_mult_limit = 5.0
    o.u_min = -1.5; o.u_max = 1.5
    for n, v in (("wx", 2.0), ("wu", 0.1), ("r_max", 4.0), ("ref_step", 0.1)):
        setattr(o.params, n, v)
    o.params.lane = lane


def track_definition(genopt, spx):
    """A car on a closed track over a terrain map: the two lookups no shipped problem uses.
    `lerp_wrap` (optim.c:410-455) gives the track's curvature at the position x (in the dynamics,
    argument = a state) and the target speed along the lap (in the cost, argument = the stage:
    a per-stage constant); `blerp` (optim.c:457-486) reads the slope of a 2-D map at (x, y)."""
    import sympy as sp
    x, y, phi, v, a, w, t, dt = sp.symbols("x y phi v a w t dt")
    wy, wv, wu, period, ds, gx0, gy0, gdx, gdy, k_slope = sp.symbols("wy wv wu period ds gx0 gy0 gdx gdy k_slope")
    track_s, track_k, track_v = (spx.ArraySymbol(n) for n in ("track_s", "track_k", "track_v"))
    slope = spx.ArraySymbol("slope")
    v_ref = spx.lerp_wrap(period, ds, 2.0 * t * dt, track_s, track_v)
    return genopt.Config(
        [x, y, phi, v], [a, w],
        {wy: 1.0, wv: 0.5, wu: 0.1, period: 10.5, ds: 0.5, gx0: -5.0, gy0: -3.0, gdx: 2.0, gdy: 0.5, k_slope: 2.0,
         track_s: None, track_k: None, track_v: None, slope: None},
        sp.Matrix([v * sp.cos(phi), v * sp.sin(phi),
                   w + spx.lerp_wrap(period, ds, x, track_s, track_k),
                   a - k_slope * spx.blerp(gx0, gy0, gdx, gdy, x, y, slope)]),
        wy * y**2 + wv * (v - v_ref)**2 + wu * (a**2 + w**2),
        end_costs=5.0 * y**2 + phi**2)


TRACK_SCALARS = dict(wy=1.0, wv=0.5, wu=0.1, period=10.5, ds=0.5, gx0=-5.0, gy0=-3.0, gdx=2.0, gdy=0.5, k_slope=2.0)
TRACK_PICK = (0, 5, 11)


def track_inputs():
    """x0 spread over several laps on both sides of the origin (the wrap, its `x < 0` branch, the
    closing segment between the last and the first sample) and beyond the map on every side (the
    clamped cells and the negative-index quirk of initInterp, optim.c:347-355)."""
    B, T, n = 12, 40, 20
    rng = np.random.default_rng(21)
    track_s = 1.0 + 0.5 * np.arange(n)                       # first = 1, last = 10.5, period 10.5: gap = 1
    track_k = 0.2 * np.sin(2 * np.pi * np.arange(n) / n)[None, :] + rng.normal(0, 0.02, (B, n))
    track_v = 3.0 + np.cos(2 * np.pi * np.arange(n) / n)[None, :] + rng.normal(0, 0.1, (B, n))
    slope = 0.3 * rng.normal(0, 1.0, (B, 12, 16)).cumsum(axis=2) / 4.0     # rows (y) x cols (x)
    x0 = np.stack([np.linspace(-16.0, 27.0, B), rng.normal(0, 0.8, B), rng.normal(0, 0.2, B),
                   rng.uniform(2.0, 4.0, B)], axis=1)
    return B, T, track_s, track_k, track_v, slope, x0


def configure_track(o, T, track_s, track_k, track_v, slope):
    o.horizon = T; o.step = 0.1; o.integrator_type = o.HEUN
    o.max_iterations = 8
    o.u_min = -3.0; o.u_max = 3.0
    for n, v in TRACK_SCALARS.items():
        setattr(o.params, n, v)
    o.params.track_s = track_s
    o.params.track_k = track_k
    o.params.track_v = track_v
    o.params.slope = slope


def zoo_inputs(name, info):
    """Random smooth inputs for a zoo model without a workload generator."""
    B = 8
    rng = np.random.default_rng(11)
    scal = {n: rng.uniform(0.5, 1.5) for n in info["scalar_names"]}
    scal.update({k: v for k, v in dict(ref_step=0.5, s_start=0.0, l=3.0, v_ch=30.0, max_delta=0.6, max_acc=2.0,
                                       min_acc=-3.0, a_offset=0.0, p_phi=50.0, pd=5.0).items() if k in scal})
    arrs = {n: np.cumsum(rng.normal(0.0, 0.02, (B, 200)), axis=1) for n in info["array_names"]}
    for n in arrs:
        if n in ("ref_v", "ref_s_max"):
            arrs[n] = 8.0 + arrs[n]
        if n == "ref_x":
            arrs[n] = np.arange(200)[None, :] * 0.5 + arrs[n]
    x0 = rng.normal(0.0, 0.1, (B, info["X"]))
    if name == "velocity_profile_time":
        x0[:, 1] = 5.0
    return B, scal, arrs, x0


def configure_zoo(o, name, horizon, C, scal, arrs, i=None):
    o.horizon = horizon; o.step = 0.1 if name != "ref_line_smoother_dk" else 0.5
    o.integrator_type = o.EULER
    o.max_iterations = 6; o.min_rel_cost_change = 0.0
    if C:
        o.barrier_weight = 100.0; o.lg_mult_limit = 0.0
    o.u_min = -1.0; o.u_max = 1.0
    for n, v in scal.items():
        setattr(o.params, n, v)
    for n, v in arrs.items():
        setattr(o.params, n, v if i is None else v[i])




def _a(v):
    return np.array(v, dtype=np.float64)


def _final(out, key, o, T):
    out[key + "/x"] = _a(o.x).reshape(T + 1, -1)
    out[key + "/u"] = _a(o.u).reshape(T, -1)
    for n in ("traj_costs", "alpha", "iterations", "termination_condition", "mu_step", "lg_iterations"):
        out[f"{key}/{n}"] = _a(getattr(o, n))


def run_single(make, zoo_info):
    """`make(model)` -> fresh CPU solver object; `zoo_info[name]` -> model_info dict."""
    out = {}
    # ---- ilr: gradient-only variant (optim.c:1010-1089) ---------------------------------------
    pb = sc.smoother(batch=3, horizon=60, max_iterations=6, forced=True, seed0=9)
    for i in range(pb.batch):
        o = sc.apply_to_single(make(pb.model), pb, i)
        o.use_quadratic_terms = False
        o.update()
        _final(out, f"ilr/{i}", o, pb.horizon)
    # ---- shift, dynamics, ct_dynamics (optim.c:1162-1177, 1512-1652) ---------------------------
    pb = sc.mpc_time(batch=4, horizon=30, max_iterations=3, forced=True, seed0=77)
    amounts = [0, 1, 2, 40]
    for i in range(pb.batch):
        o = sc.apply_to_single(make(pb.model), pb, i)
        o.update()
        _final(out, f"shift/{i}/solved", o, pb.horizon)
        out[f"shift/{i}/prev_x"] = _a(o.prev_x)
        out[f"shift/{i}/prev_k"] = _a(o.prev_k)
        out[f"shift/{i}/k"] = _a(o.k)
        out[f"shift/{i}/K"] = _a(o.K)
        o.shift(3)
        out[f"shift/{i}/x3"], out[f"shift/{i}/u3"] = _a(o.x), _a(o.u)
        out[f"shift/{i}/lam3"] = _a(o.lagrange_multiplier)
        o.shift(amounts[i])
        out[f"shift/{i}/xn"], out[f"shift/{i}/un"] = _a(o.x), _a(o.u)
        x = pb.x0[i] + 0.1
        u = np.array([0.3, -0.2])
        for t in (0, 5):
            out[f"dyn/{i}/{t}"] = _a(o.dynamics(x, u, t, 0.01))
            out[f"ctdyn/{i}/{t}"] = _a(o.ct_dynamics(x, u, t, 0.01))
    # ---- negative interpolation argument (optim.c:347-355, SURVEY.md finding 7) -----------------
    pb = sc.mpc(batch=2, horizon=20, max_iterations=1)
    o = sc.apply_to_single(make(pb.model), pb, 0)
    for j, s_r in enumerate(S_R):
        x = pb.x0[0].copy(); x[5] = s_r
        out[f"neg/{j}"] = _a(o.ct_dynamics(x, np.zeros(2), 0, 0.05))
    # ---- sticky state (SURVEY.md finding 8, appendix G10): mu / mu_step are never reset -----------
    # (set by hand: the forced-iteration plateau that produces them in G10 is decided by round-off)
    pb = sc.mpc_time(batch=2, horizon=30, max_iterations=2, forced=True, seed0=88)
    o = sc.apply_to_single(make(pb.model), pb, 0)
    o.mu, o.mu_step = 100.0, 3
    for j, (max_it, max_lg) in enumerate(STICKY):
        o.max_iterations = max_it
        o.max_lg_iterations = max_lg
        o.update()
        out[f"sticky/{j}"] = _a([o.mu_step, o.mu, o.iterations, o.termination_condition, o.traj_costs,
                                 o.lg_iterations])
        out[f"sticky/{j}/u"] = _a(o.u)
    # ---- zoo models without a generator ---------------------------------------------------------
    for name, horizon in ZOO:
        info = zoo_info[name]
        B, scal, arrs, x0 = zoo_inputs(name, info)
        for i in range(B):
            o = make(name)
            configure_zoo(o, name, horizon, info["C"], scal, arrs, i)
            o.x[0] = x0[i]
            o.update()
            _final(out, f"zoo/{name}/{i}", o, horizon)
    # ---- user-defined problem: RK4, augmented Lagrangian, end cost, lookup array ----------------------
    B, T, lanes, x0 = custom_inputs()
    for i in (0, 7, 15):
        o = make(CUSTOM)
        configure_custom(o, T, lanes[i])
        o.x[0] = x0[i]
        o.update()
        _final(out, f"custom/{i}", o, T)
        out[f"custom/{i}/lam"] = _a(o.lagrange_multiplier)
    # ---- user-defined problem with lerp_wrap (dynamics + stage constant) and blerp (2-D map) ----------
    B, T, track_s, track_k, track_v, slope, x0 = track_inputs()
    for i in TRACK_PICK:
        o = make(TRACK)
        configure_track(o, T, track_s, track_k[i], track_v[i], slope[i])
        o.x[0] = x0[i]
        o.update()
        _final(out, f"track/{i}", o, T)
        out[f"track/{i}/ctdyn"] = _a(o.ct_dynamics(x0[i], np.array([0.5, -0.1]), 3, 0.1))
    return out


def run_batched(make, zoo_info, custom_factory):
    """Same keys from the CUDA solver.  `make(model, batch, horizon_max, scenes=None)` -> BatchedOptim;
    `custom_factory[name](batch, horizon_max)` -> BatchedOptim of a user-defined problem."""
    import torch

    def npy(t):
        return t.detach().cpu().numpy().astype(np.float64)

    def final(out, key, q, i, T):
        out[key + "/x"] = npy(q.x[i]).reshape(T + 1, -1)
        out[key + "/u"] = npy(q.u[i]).reshape(T, -1)
        for n in ("traj_costs", "alpha", "iterations", "termination_condition", "mu_step", "lg_iterations"):
            out[f"{key}/{n}"] = npy(getattr(q, n)[i])

    out = {}
    pb = sc.smoother(batch=3, horizon=60, max_iterations=6, forced=True, seed0=9)
    q = sc.apply_to_batched(make(pb.model, pb.batch, pb.horizon), pb)
    q.use_quadratic_terms = False
    q.update()
    for i in range(pb.batch):
        final(out, f"ilr/{i}", q, i, pb.horizon)

    pb = sc.mpc_time(batch=4, horizon=30, max_iterations=3, forced=True, seed0=77)
    q = sc.apply_to_batched(make(pb.model, pb.batch, pb.horizon), pb)
    q.update()
    for i in range(pb.batch):
        final(out, f"shift/{i}/solved", q, i, pb.horizon)
        for n in ("prev_x", "prev_k", "k", "K"):
            out[f"shift/{i}/{n}"] = npy(getattr(q, n)[i])
    q.shift(3)
    for i in range(pb.batch):
        out[f"shift/{i}/x3"], out[f"shift/{i}/u3"] = npy(q.x[i]), npy(q.u[i])
        out[f"shift/{i}/lam3"] = npy(q.lagrange_multiplier[i])
    q.shift(np.array([0, 1, 2, 40]))
    for i in range(pb.batch):
        out[f"shift/{i}/xn"], out[f"shift/{i}/un"] = npy(q.x[i]), npy(q.u[i])
    u = np.tile(np.array([0.3, -0.2]), (pb.batch, 1))
    for t in (0, 5):
        d = npy(q.dynamics(pb.x0 + 0.1, u, t, 0.01))
        c = npy(q.ct_dynamics(pb.x0 + 0.1, u, t, 0.01))
        for i in range(pb.batch):
            out[f"dyn/{i}/{t}"], out[f"ctdyn/{i}/{t}"] = d[i], c[i]

    pb = sc.mpc(batch=2, horizon=20, max_iterations=1)
    q = sc.apply_to_batched(make(pb.model, pb.batch, pb.horizon), pb)
    for j, s_r in enumerate(S_R):
        x = pb.x0[0].copy(); x[5] = s_r
        out[f"neg/{j}"] = npy(q.ct_dynamics(np.tile(x, (2, 1)), np.zeros((2, 2)), 0, 0.05))[0]

    pb = sc.mpc_time(batch=2, horizon=30, max_iterations=2, forced=True, seed0=88)
    q = sc.apply_to_batched(make(pb.model, pb.batch, pb.horizon), pb)
    q.mu, q.mu_step = 100.0, 3
    for j, (max_it, max_lg) in enumerate(STICKY):
        q.max_iterations = max_it
        q.max_lg_iterations = max_lg
        q.update()
        out[f"sticky/{j}"] = _a([int(q.mu_step[0]), float(q.mu[0]), int(q.iterations[0]),
                                 int(q.termination_condition[0]), float(q.traj_costs[0]),
                                 int(q.lg_iterations[0])])
        out[f"sticky/{j}/u"] = npy(q.u[0])

    for name, horizon in ZOO:
        info = zoo_info[name]
        B, scal, arrs, x0 = zoo_inputs(name, info)
        q = make(name, B, horizon)
        configure_zoo(q, name, horizon, info["C"], scal, arrs)
        q.set_initial_state(x0)
        q.update()
        for i in range(B):
            final(out, f"zoo/{name}/{i}", q, i, horizon)

    B, T, lanes, x0 = custom_inputs()
    q = custom_factory[CUSTOM](B, T)
    configure_custom(q, T, lanes)
    q.set_initial_state(x0)
    q.update()
    torch.cuda.synchronize()
    for i in (0, 7, 15):
        final(out, f"custom/{i}", q, i, T)
        out[f"custom/{i}/lam"] = npy(q.lagrange_multiplier[i])

    B, T, track_s, track_k, track_v, slope, x0 = track_inputs()
    q = custom_factory[TRACK](B, T)
    configure_track(q, T, track_s, track_k, track_v, slope)
    q.set_initial_state(x0)
    q.update()
    c = npy(q.ct_dynamics(x0, np.tile(np.array([0.5, -0.1]), (B, 1)), 3, 0.1))
    torch.cuda.synchronize()
    for i in TRACK_PICK:
        final(out, f"track/{i}", q, i, T)
        out[f"track/{i}/ctdyn"] = c[i]
    return out


#: keys compared at a looser, stated tolerance: the sticky-state vector holds a cost
TOL = {}


def compare(got, want, rtol=1e-9):
    """Worst relative error over the keys of `want`; integers (counts, flags) must be equal."""
    worst, bad = 0.0, []
    for k, w in want.items():
        g = np.asarray(got[k], dtype=np.float64).reshape(np.shape(w))
        w = np.asarray(w, dtype=np.float64)
        last = k.rsplit("/", 1)[-1]
        if k.startswith("zoo/") and last in ("mu_step", "alpha"):
            continue        # forced iterations: the last line searches are decided by round-off (finding 9)
        if last in ("iterations", "termination_condition", "mu_step", "lg_iterations"):
            if not np.array_equal(g, w):
                bad.append((k, g.tolist(), w.tolist()))
            continue
        scale = max(np.max(np.abs(w)), 1e-300) if w.size else 1.0
        e = float(np.max(np.abs(g - w)) / scale) if w.size else 0.0
        if k.startswith("neg/") or k.startswith("dyn/") or k.startswith("ctdyn/") or last == "ctdyn":
            lim = 1e-12
        else:
            lim = rtol
        if e > lim:
            bad.append((k, e))
        worst = max(worst, e)
    return worst, bad
