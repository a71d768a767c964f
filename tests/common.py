"""Shared test machinery: the parity cases, per-iteration traces and comparisons.

A *case* is a small seeded problem batch plus solver settings.  The same case is
run through
  * the real reference solver (only in the build container, via oracle/_ref) to
    record the golden vectors in tests/golden/ (tests/golden/make_golden.py),
  * the CPU oracle restatement (oracle/ilqr_oracle.c),
  * the CUDA solver through the C ABI (``-m gpu`` tests),
and compared iteration by iteration.

Per-iteration traces use the prefix property of the reference (SURVEY.md
appendix D): a run with ``max_iterations = s`` is an exact prefix of the run
with ``s + 1``, so iteration ``s`` of a solve is observed by solving a copy of
the initial problem with ``max_iterations = s``.
"""

import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tpl_b200 import scenarios as sc   # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

RTOL = 1e-9          # BASELINE.json: per-iteration controls, states and cost within 1e-9 relative (fp64)


# ---------------------------------------------------------------------------------
# cases
# ---------------------------------------------------------------------------------

def _tight_bounds(pb):
    pb.u_min[:] = -0.01
    pb.u_max[:] = 0.01
    return pb


def _rk4(pb):
    pb.integrator = sc.RK4
    return pb


def _euler(pb):
    pb.integrator = sc.EULER
    return pb


def _utm(pb):
    return pb


CASES = {
    # name: (generator, kwargs, post-edit, iterations traced, flavour of the golden reference)
    "mpc_time_forced": (sc.mpc_time, dict(batch=3, horizon=100, max_iterations=10, forced=True), None, 10, "fast"),
    "mpc_time_default": (sc.mpc_time, dict(batch=3, horizon=40, max_iterations=20, forced=False, seed0=10), None, 20, "fast"),
    "mpc_time_multistart": (sc.mpc_time, dict(batch=8, scenes=2, horizon=60, max_iterations=6, forced=True, seed0=20), None, 6, "fast"),
    "mpc_time_rk4": (sc.mpc_time, dict(batch=2, horizon=30, max_iterations=5, forced=True, seed0=30), _rk4, 5, "fast"),
    "mpc_time_euler": (sc.mpc_time, dict(batch=2, horizon=30, max_iterations=5, forced=True, seed0=40), _euler, 5, "fast"),
    "mpc_time_utm": (sc.mpc_time, dict(batch=2, horizon=60, max_iterations=6, forced=True, seed0=50,
                                       origin=(5.7e5, 5.36e6)), None, 6, "fast"),
    "lateral_forced": (sc.lateral, dict(batch=3, horizon=200, max_iterations=10, forced=True), None, 10, "fast"),
    "lateral_default": (sc.lateral, dict(batch=3, horizon=200, max_iterations=10, forced=False, seed0=2), None, 10, "fast"),
    "lateral_al": (sc.lateral, dict(batch=2, horizon=200, max_iterations=5, forced=False, seed0=5,
                                    augmented_lagrangian=True), None, 5, "fast"),
    "lateral_tight": (sc.lateral, dict(batch=2, horizon=200, max_iterations=10, forced=False, seed0=2), _tight_bounds, 10, "fast"),
    "lateral_pinned": (sc.lateral, dict(batch=2, horizon=200, max_iterations=10, forced=False, seed0=2, pin_prefix=20), None, 10, "fast"),
    "velocity_default": (sc.velocity, dict(batch=2, horizon=250, max_iterations=20, forced=False), None, 20, "fast"),
    "smoother_default": (sc.smoother, dict(batch=2, horizon=250, max_iterations=5, forced=False), None, 5, "fast"),
    "mpc_strict": (sc.mpc, dict(batch=2, horizon=60, max_iterations=5, forced=True), None, 5, "strict"),
}

#: cases whose model has an ill-conditioned derivative path (SURVEY.md finding 6):
#: compared with a looser, stated tolerance
LOOSE = {"mpc_strict": 1e-5}


def make_case(name):
    gen, kw, edit, iters, flavour = CASES[name]
    pb = gen(**kw)
    if edit is not None:
        pb = edit(pb)
    return pb, iters, flavour


# ---------------------------------------------------------------------------------
# traces
# ---------------------------------------------------------------------------------

from tpl_b200 import parity   # noqa: E402
from tpl_b200.parity import (SCALARS, analyse, batched_problem_trace, compare_traces, rel_err,   # noqa: E402,F401
                             same_decisions)

_snapshot_single = parity.snapshot_single


def trace_single(factory, pb, i, iters):
    """[snapshot after max_iterations = 0..iters] for problem ``i`` solved by an
    object with the reference ``Optim`` interface (real reference or CPU oracle)."""
    return parity.trace_single(sc.apply_to_single(factory(), pb, i), iters)


def derivatives_single(factory, pb, i):
    """Derivative blocks and gains of the initial trajectory (one iteration)."""
    q = sc.apply_to_single(factory(), pb, i)
    q.max_iterations = 1
    q.update()
    T = q.horizon
    shapes = dict(fx=(T, q_dim(q, "x"), q_dim(q, "x")), fu=(T, q_dim(q, "x"), q_dim(q, "u")),
                  lx=(T, q_dim(q, "x")), lu=(T, q_dim(q, "u")),
                  lxx=(T, q_dim(q, "x"), q_dim(q, "x")), luu=(T, q_dim(q, "u"), q_dim(q, "u")),
                  lux=(T, q_dim(q, "u"), q_dim(q, "x")), k=(T, q_dim(q, "u")),
                  K=(T, q_dim(q, "u"), q_dim(q, "x")))
    return {n: np.array(getattr(q, n), dtype=np.float64).reshape(s) for n, s in shapes.items()}


def q_dim(q, what):
    a = np.asarray(q.x if what == "x" else q.u)
    return 1 if a.ndim == 1 else a.shape[1]


def trace_batched(bopt_factory, pb, iters):
    """Same trace for the whole batch on the CUDA solver: list over s of dicts of
    (B, ...) numpy arrays."""
    return parity.trace_batched(sc.apply_to_batched(bopt_factory(), pb), iters)


def derivatives_batched(bopt_factory, pb):
    q = sc.apply_to_batched(bopt_factory(), pb)
    q.max_iterations = 1
    q.update()
    B, T = q.batch, q.horizon
    out = {}
    for n in ("fx", "fu", "lx", "lu", "lxx", "luu", "lux", "k", "K"):
        out[n] = getattr(q, n).cpu().numpy()
    return out


# ---------------------------------------------------------------------------------
# golden files
# ---------------------------------------------------------------------------------

def golden_path(name):
    return os.path.join(GOLDEN_DIR, name + ".npz")


def save_golden(name, traces, derivs):
    """traces: list over problems of list over s of snapshots."""
    flat = {}
    for i, tr in enumerate(traces):
        for s, snap in enumerate(tr):
            for k, v in snap.items():
                flat[f"p{i}/s{s}/{k}"] = np.asarray(v)
        for k, v in derivs[i].items():
            flat[f"p{i}/d/{k}"] = v
    flat["meta/problems"] = np.asarray(len(traces))
    flat["meta/iters"] = np.asarray(len(traces[0]) - 1)
    np.savez_compressed(golden_path(name), **flat)


def load_golden(name):
    z = np.load(golden_path(name))
    n, iters = int(z["meta/problems"]), int(z["meta/iters"])
    traces, derivs = [], []
    for i in range(n):
        tr = []
        for s in range(iters + 1):
            snap = {k: z[f"p{i}/s{s}/{k}"] for k in ("x", "u")}
            snap.update({k: float(z[f"p{i}/s{s}/{k}"]) for k in SCALARS})
            tr.append(snap)
        traces.append(tr)
        derivs.append({k: z[f"p{i}/d/{k}"] for k in ("fx", "fu", "lx", "lu", "lxx", "luu", "lux", "k", "K")})
    return traces, derivs
