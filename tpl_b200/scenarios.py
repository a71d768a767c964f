"""Synthetic road / reference-path problem batches of the benchmark shapes.

The reference ships no solver fixtures, so the workloads named in
BASELINE.json are generated here, seeded per problem with
``numpy.random.default_rng(problem_index)`` as SURVEY.md §8(d) specifies, with
solver settings taken from the reference's callers:

* ``mpc_time``  — `ModelPredictiveControllerTime` (control/model_predictive_controller_time.py:72-142)
  and data/params/control/default/state.json,
* ``mpc``       — `ModelPredictiveController` (control/model_predictive_controller.py:69-155),
* ``lateral``   — `PathOptim` (planning/path_vel_decomp/path_optim.py:99-168) with the acc_2024 weights,
* ``velocity``  — `VelocityOptim` (planning/path_vel_decomp/velocity_optim.py:60-66, 153-240),
* ``smoother``  — `PathSmoothing` (planning/path_vel_decomp/path_smoothing.py:24-75).

A :class:`ProblemBatch` is a plain description; ``apply_to_single`` loads problem
``i`` into any object with the reference ``Optim`` interface (the real reference,
the CPU oracle) and ``apply_to_batched`` loads the whole batch into a
:class:`tpl_b200.batched.BatchedOptim`.
"""

from dataclasses import dataclass, field
from typing import Dict

import numpy as np

EULER, HEUN, RK4 = 0, 1, 2


@dataclass
class ProblemBatch:
    model: str
    horizon: int
    step: float
    integrator: int
    max_iterations: int
    max_lg_iterations: int = 1
    min_rel_cost_change: float = 1e-6
    barrier_weight: float = 1.0
    lg_mult_limit: float = np.inf
    scalars: Dict[str, np.ndarray] = field(default_factory=dict)   # name -> (S,)
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)    # name -> (S, L)
    scene_of: np.ndarray = None                                    # (B,) int32
    x0: np.ndarray = None                                          # (B, X)
    u0: np.ndarray = None                                          # (B, T, U)
    u_min: np.ndarray = None                                       # (B, T, U)
    u_max: np.ndarray = None                                       # (B, T, U)

    @property
    def batch(self):
        return self.x0.shape[0]

    @property
    def scenes(self):
        return next(iter(self.scalars.values())).shape[0]

    def subset(self, idx):
        """Problems ``idx`` with their scenes (re-indexed)."""
        idx = np.asarray(idx)
        scenes, inv = np.unique(self.scene_of[idx], return_inverse=True)
        return ProblemBatch(
            self.model, self.horizon, self.step, self.integrator, self.max_iterations,
            self.max_lg_iterations, self.min_rel_cost_change, self.barrier_weight, self.lg_mult_limit,
            {k: v[scenes] for k, v in self.scalars.items()},
            {k: v[scenes] for k, v in self.arrays.items()},
            inv.astype(np.int32), self.x0[idx], self.u0[idx], self.u_min[idx], self.u_max[idx])


# ---------------------------------------------------------------------------------
# loading a batch into solver objects
# ---------------------------------------------------------------------------------

def apply_to_single(opt, pb: ProblemBatch, i: int):
    """Problem ``i`` -> one reference-interface ``Optim`` object."""
    s = int(pb.scene_of[i])
    opt.integrator_type = pb.integrator
    opt.horizon = pb.horizon
    opt.step = pb.step
    opt.max_iterations = pb.max_iterations
    opt.max_lg_iterations = pb.max_lg_iterations
    opt.min_rel_cost_change = pb.min_rel_cost_change
    if opt.barrier_weight.size:
        opt.lg_mult_limit = pb.lg_mult_limit
        opt.barrier_weight[:] = pb.barrier_weight
        opt.lagrange_multiplier[:] = 0.0
    for k, v in pb.scalars.items():
        setattr(opt.params, k, float(v[s]))
    for k, v in pb.arrays.items():
        setattr(opt.params, k, v[s])
    T = pb.horizon
    opt.u_min = pb.u_min[i].reshape(opt.u_min.shape)
    opt.u_max = pb.u_max[i].reshape(opt.u_max.shape)
    opt.x[0] = pb.x0[i]
    opt.u = pb.u0[i].reshape(opt.u.shape)
    return opt


def apply_to_batched(bopt, pb: ProblemBatch):
    """Whole batch -> :class:`tpl_b200.batched.BatchedOptim` (host arrays in,
    one upload per field)."""
    bopt.integrator_type = pb.integrator
    bopt.horizon = pb.horizon
    bopt.step = pb.step
    bopt.max_iterations = pb.max_iterations
    bopt.max_lg_iterations = pb.max_lg_iterations
    bopt.min_rel_cost_change = pb.min_rel_cost_change
    if bopt.C:
        bopt.lg_mult_limit = pb.lg_mult_limit
        bopt.barrier_weight = pb.barrier_weight
        bopt.lagrange_multiplier = 0.0
    bopt.scene_index = pb.scene_of
    for k, v in pb.scalars.items():
        setattr(bopt.params, k, v)
    for k, v in pb.arrays.items():
        setattr(bopt.params, k, v)
    bopt.u_min = pb.u_min
    bopt.u_max = pb.u_max
    bopt.set_initial_state(pb.x0)
    bopt.u = pb.u0
    return bopt


# ---------------------------------------------------------------------------------
# generators
# ---------------------------------------------------------------------------------

def reference_path(rng, n, ds):
    """Curvy road centre line: kappa(s) = 0.02 sin(s/15 + phase), integrated with
    step ``ds``.  Returns x, y, heading, curvature, each of length ``n``."""
    phase = rng.uniform(0.0, 6.0)
    s = np.arange(n) * ds
    kappa = 0.02 * np.sin(s / 15.0 + phase)
    heading = np.cumsum(kappa) * ds
    x = np.cumsum(np.cos(heading)) * ds
    y = np.cumsum(np.sin(heading)) * ds
    return x, y, heading, kappa


def _bounds(B, T, lo, hi):
    lo = np.broadcast_to(np.asarray(lo, dtype=np.float64), (B, T, len(np.atleast_1d(lo)))).copy()
    hi = np.broadcast_to(np.asarray(hi, dtype=np.float64), (B, T, len(np.atleast_1d(hi)))).copy()
    return lo, hi


def _const_scalars(S, **kw):
    return {k: np.full(S, float(v)) for k, v in kw.items()}


def mpc_time(batch, horizon=100, max_iterations=10, forced=True, seed0=0,
             scenes=None, origin=(0.0, 0.0)):
    """Config #2/#3/#5 workload: `trajectory_tracking_mpc_time` (X=6, U=2, C=4).

    ``scenes=None``: every problem has its own reference trajectory.  Otherwise
    ``scenes`` trajectories are shared by ``batch/scenes`` multi-start problems
    each (config #3): starts perturb x0 and the control warm start."""
    S = batch if scenes is None else scenes
    per = batch // S
    assert per * S == batch
    n_ref, ref_dt, v_ref = 80, 0.1, 8.0
    arr = {k: np.zeros((S, n_ref)) for k in ("ref_x", "ref_y", "ref_phi", "ref_v")}
    x0 = np.zeros((batch, 6))
    u0 = np.zeros((batch, horizon, 2))
    u_lo, u_hi = np.array([-3.0, -1.0]), np.array([1.5, 1.0])
    for s in range(S):
        rng = np.random.default_rng(seed0 + s)
        px, py, ph, _ = reference_path(rng, n_ref, v_ref * ref_dt)
        arr["ref_x"][s], arr["ref_y"][s], arr["ref_phi"][s] = px + origin[0], py + origin[1], ph
        arr["ref_v"][s] = v_ref
        base = np.array([0.0, rng.uniform(-.5, .5), rng.uniform(-.05, .05), 0.0, rng.uniform(5.0, 9.0), 0.0])
        base[:2] += origin
        if scenes is None:
            x0[s] = base
        else:
            sl = slice(s * per, (s + 1) * per)
            x0[sl] = base + rng.normal(0.0, 1.0, (per, 6)) * np.array([.3, .3, .03, 0.0, .5, 0.0])
            u0[sl] = np.clip(rng.normal(0.0, .2, (per, horizon, 2)), u_lo, u_hi)
    scal = _const_scalars(
        S, pd=5.0, pv=5.0, pdelta=0.0, min_pdelta_dot=0.1, pdelta_dot=0.1, min_p_phi_dot=0.0,
        p_phi_dot=0.0, p_phi=0.0, pa=2.0, pj=0.5, l=3.165, v_ch=32.0, cog_pos=0.5,
        ref_dt=ref_dt, ref_t_offset=0.0, max_delta=0.7, max_acc=3.0, min_acc=-3.0, a_offset=0.0)
    lo, hi = _bounds(batch, horizon, u_lo, u_hi)
    return ProblemBatch("trajectory_tracking_mpc_time", horizon, 0.05, HEUN, max_iterations,
                        1, 0.0 if forced else 1e-6, 1e4, 0.0, scal, arr,
                        np.repeat(np.arange(S, dtype=np.int32), per), x0, u0, lo, hi)


def mpc(batch, horizon=100, max_iterations=10, forced=True, seed0=0, origin=(0.0, 0.0)):
    """Secondary workload: path-relative `trajectory_tracking_mpc` (X=7, U=2, C=4)."""
    n_ref, ds = 100, 0.5
    arr = {k: np.zeros((batch, n_ref)) for k in ("ref_x", "ref_y", "ref_phi", "ref_k", "ref_v")}
    x0 = np.zeros((batch, 7))
    for b in range(batch):
        rng = np.random.default_rng(seed0 + b)
        px, py, ph, kp = reference_path(rng, n_ref, ds)
        arr["ref_x"][b], arr["ref_y"][b] = px + origin[0], py + origin[1]
        arr["ref_phi"][b], arr["ref_k"][b], arr["ref_v"][b] = ph, kp, 8.0
        x0[b] = [origin[0], origin[1] + rng.uniform(-.5, .5), rng.uniform(-.05, .05), 0.0,
                 rng.uniform(4.0, 9.0), 0.2, 0.0]
    scal = _const_scalars(
        batch, pd=10.0, pv=5.0, pdelta=0.0, min_pdelta_dot=0.1, pdelta_dot=0.0, min_p_phi_dot=0.0,
        p_phi_dot=0.05, p_phi=1000.0, p_phi_ref_dot_diff=0.01, pa=2.0, pj=0.5, l=3.165, v_ch=32.0,
        ref_step=ds, max_delta=0.7, max_acc=2.0, min_acc=-3.0, a_offset=0.0)
    lo, hi = _bounds(batch, horizon, [-3.0, -1.0], [1.5, 1.0])
    return ProblemBatch("trajectory_tracking_mpc", horizon, 0.05, HEUN, max_iterations,
                        1, 0.0 if forced else 1e-6, 1e4, 0.0, scal, arr,
                        np.arange(batch, dtype=np.int32), x0, np.zeros((batch, horizon, 2)), lo, hi)


def lateral(batch, horizon=200, max_iterations=10, forced=True, seed0=0,
            augmented_lagrangian=False, pin_prefix=0):
    """Config #4 workload: `lateral_profile` (X=2, U=1, C=2) with the acc_2024
    weights, a corridor with a 20-stage obstacle bump, pure penalty
    (``lg_mult_limit=0``, path_optim.py:102-104) or the AL variant."""
    arr = {k: np.zeros((batch, horizon)) for k in ("k_ref", "d_offset", "d_lower_constr", "d_upper_constr")}
    x0 = np.zeros((batch, 2))
    s = np.arange(horizon) * 0.5
    for b in range(batch):
        rng = np.random.default_rng(seed0 + b)
        arr["k_ref"][b] = 0.02 * np.sin(s / 12.0 + rng.uniform(0.0, 6.0))
        lower = np.full(horizon, -1.5)
        start = int(rng.integers(30, min(121, max(31, horizon - 20))))
        lower[start:start + 20] = 0.4
        arr["d_lower_constr"][b] = lower
        arr["d_upper_constr"][b] = 1.5
        arr["d_offset"][b] = np.clip(lower + 0.8, -1.0, 1.2)
        x0[b, 0] = rng.uniform(-.5, .5)
    scal = _const_scalars(batch, ref_step=0.5, w_d=0.2, w_v_d=1.0, w_a_d=200.0, w_k=200.0)
    lo, hi = _bounds(batch, horizon, [-2.5], [2.5])
    if pin_prefix:
        lo[:, :pin_prefix] = 0.0
        hi[:, :pin_prefix] = 0.0
    return ProblemBatch("lateral_profile", horizon, 0.5, EULER, max_iterations,
                        3 if augmented_lagrangian else 1, 0.0 if forced else 1e-6, 1000.0,
                        0.1 if augmented_lagrangian else 0.0, scal, arr,
                        np.arange(batch, dtype=np.int32), x0, np.zeros((batch, horizon, 1)), lo, hi)


def velocity(batch, horizon=250, max_iterations=20, forced=False, seed0=0):
    """RSTP speed step: `velocity_profile_space` (X=2, U=1, C=5), settings of
    `VelocityOptim` (velocity_optim.py:60-66, 153-240): speed limit profile with
    a slow zone, curvature from a synthetic path, open time windows."""
    names = ("ref_v", "ref_k", "ref_t_max", "ref_t_min", "ref_t_offset", "ref_v_weight")
    arr = {k: np.zeros((batch, horizon)) for k in names}
    x0 = np.zeros((batch, 2))
    for b in range(batch):
        rng = np.random.default_rng(seed0 + b)
        _, _, _, kp = reference_path(rng, horizon, 0.5)
        v_lim = np.full(horizon, rng.uniform(9.0, 14.0))
        z = int(rng.integers(40, horizon - 60))
        v_lim[z:z + 40] = rng.uniform(4.0, 7.0)
        arr["ref_v"][b], arr["ref_k"][b] = v_lim, kp * 3.0
        arr["ref_t_max"][b], arr["ref_t_offset"][b], arr["ref_v_weight"][b] = 10e10, 1.0, 1.0
        x0[b] = [rng.uniform(5.0, 9.0), 0.0]
    scal = _const_scalars(batch, p_v=0.1, p_a=2.0, max_a_total=5.0, ref_step=0.5)
    lo, hi = _bounds(batch, horizon, [-2.0], [2.0])
    return ProblemBatch("velocity_profile_space", horizon, 0.5, EULER, max_iterations,
                        1, 0.0 if forced else 1e-6, 1000.0, 0.1, scal, arr,
                        np.arange(batch, dtype=np.int32), x0, np.zeros((batch, horizon, 1)), lo, hi)


def smoother(batch, horizon=250, max_iterations=5, forced=False, seed0=0):
    """`ref_line_smoother_k` (X=3, U=1, C=0), settings of `PathSmoothing`
    (path_smoothing.py:24-75): noisy centre line to be smoothed."""
    arr = {k: np.zeros((batch, horizon)) for k in ("ref_x", "ref_y")}
    x0 = np.zeros((batch, 3))
    for b in range(batch):
        rng = np.random.default_rng(seed0 + b)
        px, py, ph, _ = reference_path(rng, horizon, 0.5)
        arr["ref_x"][b] = px + rng.normal(0.0, 0.05, horizon)
        arr["ref_y"][b] = py + rng.normal(0.0, 0.05, horizon)
        x0[b] = [px[0], py[0], ph[0]]
    scal = _const_scalars(batch, w_pos=1.0, w_k=0.1, ref_step=0.5)
    lo, hi = _bounds(batch, horizon, [-1.0], [1.0])
    return ProblemBatch("ref_line_smoother_k", horizon, 0.5, EULER, max_iterations,
                        1, 0.0 if forced else 1e-6, 1.0, np.inf, scal, arr,
                        np.arange(batch, dtype=np.int32), x0, np.zeros((batch, horizon, 1)), lo, hi)


GENERATORS = {
    "trajectory_tracking_mpc_time": mpc_time,
    "trajectory_tracking_mpc": mpc,
    "lateral_profile": lateral,
    "velocity_profile_space": velocity,
    "ref_line_smoother_k": smoother,
}
