"""Seeded synthetic inputs for the profile-shaping routines (`tpl_b200.prep`): speed-limit
profiles with steps and curvature dips for the velocity ramp, lateral corridors with obstacle
bumps for the evasive-offset ramp.  Used by the tests, the golden generator and bench.py, so all
of them see the same numbers."""

import numpy as np

# planning defaults of the reference's parameter files (data/params/planning/default)
VELOCITY_DEFAULTS = dict(a_min=-2.5, a_max=1.5, j_min=-1.5, j_max=1.0, v_min=1.0, step=1.0)
LATERAL_DEFAULTS = dict(step=0.5, evasion_sharpness=4.0, gap=0.3)


def velocity_case(seed, n=250, with_v0=True, with_a0=True, **over):
    rng = np.random.default_rng(seed)
    p = dict(VELOCITY_DEFAULTS, **over)
    s = np.arange(n) * p["step"]
    lim = np.full(n, rng.uniform(8.0, 16.0))
    for _ in range(rng.integers(1, 4)):                       # speed-limit steps
        a = rng.integers(10, n - 10)
        lim[a:] = rng.uniform(3.0, 16.0)
    kappa = 0.08 * np.sin(s / rng.uniform(12.0, 40.0) + rng.uniform(0, 6.0)) ** 2
    lim = np.minimum(lim, np.sqrt(2.0 / np.maximum(kappa, 1e-4)))   # lateral-acceleration limit
    if rng.uniform() < 0.5:                                   # stop point
        a = rng.integers(n // 2, n - 5)
        lim[a:a + 5] = 0.0
    return dict(p, lim_v=lim, v0=float(rng.uniform(0.0, 14.0)) if with_v0 else None,
                a0=float(rng.uniform(-1.0, 1.0)) if with_a0 else None)


def velocity_cases():
    cases = [velocity_case(100 + i) for i in range(6)]
    cases.append(velocity_case(200, with_v0=False, with_a0=False))
    cases.append(velocity_case(201, with_v0=True, with_a0=False))
    cases.append(velocity_case(202, n=40, step=2.0))
    cases.append(velocity_case(203, n=300, v_min=0.5))
    return cases


def lateral_case(seed, n=200, horizon=None, **over):
    rng = np.random.default_rng(seed)
    p = dict(LATERAL_DEFAULTS, **over)
    horizon = n if horizon is None else horizon
    lower = np.full(n, -1.5) + 0.1 * np.sin(np.arange(n) / 9.0 + rng.uniform(0, 6))
    upper = np.full(n, 1.5) + 0.1 * np.cos(np.arange(n) / 7.0 + rng.uniform(0, 6))
    for _ in range(rng.integers(1, 3)):                       # obstacles pushing the lower bound up
        a = rng.integers(20, max(21, horizon - 30))
        lower[a:a + rng.integers(8, 25)] = rng.uniform(-0.6, 0.5)
    path = np.zeros((n, 6))
    path[:, 5] = np.maximum(rng.uniform(2.0, 12.0) + np.cumsum(rng.normal(0, 0.05, n)), 0.0)
    if rng.uniform() < 0.3:
        path[rng.integers(0, n), 5] = 0.0                     # exercises the max(v, 1e-8) guard
    # the reference passes (lower, -upper_constraint) for one side and the mirrored pair for the other
    return dict(p, horizon=horizon, proj_distance=float(rng.uniform(-0.5, 0.5)), path=path,
                lower=lower, upper=upper)


def lateral_cases():
    cases = [lateral_case(300 + i) for i in range(6)]
    cases.append(lateral_case(400, n=250, horizon=200))
    cases.append(lateral_case(401, n=60, step=1.0))
    cases.append(lateral_case(402, evasion_sharpness=0.5, gap=0.0))
    return cases


def velocity_batch(batch, n=250, seed0=0):
    """Arrays for a whole batch: lim_v (B, n), v0 (B,), a0 (B,), shared scalars."""
    cs = [velocity_case(seed0 + i, n=n) for i in range(batch)]
    return dict(VELOCITY_DEFAULTS, lim_v=np.stack([c["lim_v"] for c in cs]),
                v0=np.array([c["v0"] for c in cs]), a0=np.array([c["a0"] for c in cs]))


def lateral_batch(batch, n=200, seed0=0):
    cs = [lateral_case(seed0 + i, n=n) for i in range(batch)]
    return dict(LATERAL_DEFAULTS, horizon=n, proj_distance=np.array([c["proj_distance"] for c in cs]),
                path_v=np.stack([c["path"][:, 5] for c in cs]), lower=np.stack([c["lower"] for c in cs]),
                upper=np.stack([c["upper"] for c in cs]))


def shift_cases():
    """Trajectory arrays (n, rows) with the arc length travelled since the last cycle: inside the
    grid, beyond its end (extrapolation), backwards, exact multiples of the step, zero."""
    cases = []
    for i, (n, rows, step, arc) in enumerate([(250, 3, 1.0, 0.37), (250, 1, 0.5, 2.5), (100, 2, 0.3, 0.9),
                                              (60, 4, 2.0, -0.8), (40, 2, 0.1, 7.3), (250, 5, 1.0, 0.0),
                                              (30, 1, 0.7, 30.0), (120, 2, 0.25, 1e-9)]):
        rng = np.random.default_rng(800 + i)
        arr = np.cumsum(rng.normal(0.0, 1.0, (n, rows)), axis=0) + rng.uniform(-5, 5)
        cases.append(dict(arr=arr, step=step, arc_len=arc))
    return cases


EGO_DEFAULTS = dict(wheel_base=2.9, v_ch=30.0, max_v=20.0, min_v=0.0, max_steer_angle=0.6)


def ego_case(seed, steps=300, dt=0.01, acc_dead_time=0.18, steer_dead_time=0.12, **over):
    """A driven vehicle: initial state, parameters and the command sequence of a closed-loop run
    (smooth steering and acceleration commands with a few steps; saturating in places)."""
    rng = np.random.default_rng(seed)
    p = dict(EGO_DEFAULTS, acc_dead_time=acc_dead_time, steer_dead_time=steer_dead_time, **over)
    t = np.arange(steps) * dt
    acc = 1.5 * np.sin(t * rng.uniform(0.5, 3.0) + rng.uniform(0, 6)) + np.where(t > rng.uniform(0.5, 2.5), -2.0, 0.5)
    steer = 0.5 * np.sin(t * rng.uniform(0.5, 4.0) + rng.uniform(0, 6)) + rng.uniform(-0.3, 0.3)
    init = dict(x=rng.uniform(-5, 5), y=rng.uniform(-5, 5), yaw=rng.uniform(-3.5, 3.5), v=rng.uniform(0.0, 15.0),
                a=0.0, steer_angle=0.0)
    return dict(params=p, init=init, dt=dt, control_acc=acc, control_steer=steer)


def ego_cases():
    return [ego_case(600), ego_case(601, dt=0.02), ego_case(602, acc_dead_time=0.0, steer_dead_time=0.0),
            ego_case(603, acc_dead_time=0.05, steer_dead_time=0.3, dt=0.01), ego_case(604, steps=150, dt=0.05),
            ego_case(605, v_ch=12.0, max_v=8.0)]


# ---- row f1: reference-path preparation (util.resample_path, util.project) -----------------------
def path_case(seed, n=120, ds=0.7, jitter=0.15, duplicates=False):
    """A planned trajectory as `ModelPredictiveController.update` receives it
    (control/model_predictive_controller.py:116-128): columns x, y, orientation, s, curvature,
    velocity, irregularly spaced along a curvy road (so resampling has real work to do)."""
    rng = np.random.default_rng(seed)
    steps = ds * (1.0 + jitter * rng.uniform(-1.0, 1.0, n))
    s = np.concatenate([[0.0], np.cumsum(steps[:-1])])
    kappa = 0.03 * np.sin(s / 11.0 + rng.uniform(0.0, 6.0))
    heading = rng.uniform(-3.0, 3.0) + np.concatenate([[0.0], np.cumsum(kappa[:-1] * steps[:-1])])
    x = rng.uniform(-5.0, 5.0) + np.concatenate([[0.0], np.cumsum(np.cos(heading[:-1]) * steps[:-1])])
    y = rng.uniform(-5.0, 5.0) + np.concatenate([[0.0], np.cumsum(np.sin(heading[:-1]) * steps[:-1])])
    v = 8.0 + 2.0 * np.sin(s / 20.0 + rng.uniform(0.0, 6.0))
    path = np.stack([x, y, heading, s, kappa, v], axis=1)
    if duplicates:                                   # repeated points are dropped by resample (utils.cpp:431-437)
        path = np.insert(path, [10, 10, 55], path[[10, 10, 55]], axis=0)[:n]
    return path


def path_cases():
    """(path, step_size, steps, start_index, zero_vel_at_end): the MPC call (ref_step, 100 samples, zero
    velocity at the end), the lateral planner's call (opt.step, horizon), a start inside the path, a
    request that runs past the end of the path (extrapolation branch, util.py:164-168)."""
    return [
        (path_case(1), 0.5, 100, 0, True),
        (path_case(2), 0.5, 100, 0, False),
        (path_case(3, n=150), 0.5, 200, 0, False),        # longer than the path: extrapolates
        (path_case(4), 0.35, 120, 17, True),
        (path_case(5, n=80, ds=1.1), 1.0, 60, 3, False),
    ]


def path_batch(batch, n=120, seed0=0):
    """`batch` paths of `n` points, (B, n, 6), plus query positions near them, (B, 2)."""
    paths = np.stack([path_case(seed0 + b, n=n) for b in range(batch)])
    rng = np.random.default_rng(seed0 + 12345)
    k = rng.integers(5, n - 5, batch)
    off = rng.normal(0.0, 0.6, (batch, 2))
    pos = paths[np.arange(batch), k, :2] + off
    return paths, pos
