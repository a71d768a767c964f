"""Per-iteration comparison of two solvers with the reference ``Optim`` interface.

The reference has no per-iteration hook, but a run with ``max_iterations = s`` is an exact
prefix of the run with ``s + 1``, so iteration ``s`` of a solve is
observed by solving a copy of the initial problem with ``max_iterations = s``.  Used by the
parity tests (tests/common.py) and by ``bench.py`` to explain a problem whose final solution
differs from the reference's: either a real error, or a *decision flip* on the round-off
plateau of a forced-iteration run — two builds of the reference's own code disagree there
as well (SURVEY.md finding 9).

Nothing here computes: both solvers are handed in by the caller.
"""

import copy

import numpy as np

RTOL = 1e-9          # BASELINE.json: per-iteration controls, states and cost within 1e-9 relative (fp64)
PLATEAU = 1e-9       # a line search that compares costs closer than this (relative) is decided by round-off
AFTER_FLIP = 1e-6    # after such a flip both solvers are on the plateau: costs stay this close

SCALARS = ("traj_costs", "alpha", "mu_step", "iterations", "termination_condition",
           "improved", "trajectory_changed", "lg_iterations")


def snapshot_single(q):
    d = {"x": np.array(q.x, dtype=np.float64).reshape(q.horizon + 1, -1),
         "u": np.array(q.u, dtype=np.float64).reshape(q.horizon, -1)}
    for s in SCALARS:
        d[s] = float(getattr(q, s))
    return d


def trace_single(base, iters):
    """[snapshot after max_iterations = 0..iters] of a configured, not yet solved object."""
    out = []
    for s in range(iters + 1):
        q = copy.deepcopy(base)
        q.max_iterations = s
        q.update()
        out.append(snapshot_single(q))
    return out


def trace_batched(base, iters):
    """Same trace for a whole ``BatchedOptim``: list over s of dicts of (B, ...) numpy arrays."""
    out = []
    for s in range(iters + 1):
        q = copy.deepcopy(base)
        q.max_iterations = s
        q.update()
        T = q.horizon
        d = {"x": q.x.cpu().numpy().reshape(q.batch, T + 1, -1),
             "u": q.u.cpu().numpy().reshape(q.batch, T, -1)}
        for n in SCALARS:
            d[n] = getattr(q, n).cpu().numpy().astype(np.float64)
        out.append(d)
    return out


def batched_problem_trace(tr, i):
    """Slice problem ``i`` out of a batched trace."""
    return [{k: (v[i] if isinstance(v, np.ndarray) else v) for k, v in snap.items()} for snap in tr]


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale) if a.size else 0.0


def same_decisions(a, b):
    return (np.isclose(a["alpha"], b["alpha"], rtol=1e-12, atol=0.0)
            and int(a["mu_step"]) == int(b["mu_step"])
            and int(a["iterations"]) == int(b["iterations"])
            and int(a["termination_condition"]) == int(b["termination_condition"])
            and int(a["improved"]) == int(b["improved"])
            and int(a["trajectory_changed"]) == int(b["trajectory_changed"]))


def _snap_err(a, b):
    return max(rel_err(a["x"], b["x"]), rel_err(a["u"], b["u"]),
               abs(a["traj_costs"] - b["traj_costs"]) / max(abs(b["traj_costs"]), 1e-300))


def analyse(test, ref):
    """Compare two single-problem traces iteration by iteration.

    ``worst``: largest relative error of x, u, cost over the iterations before any decision
    differed.  ``flip``: first iteration whose decisions (step size, regularisation step, flags)
    differ, or None.  ``plateau``: that line search was decided at round-off level — the costs the
    two solvers ended the iteration with agree within ``PLATEAU`` of each other, or the reference's
    own cost moved less than ``PLATEAU`` in it.  ``after``: largest relative COST difference from
    the flip on (the trajectories may legitimately differ there)."""
    worst, flip, plateau, after = 0.0, None, False, 0.0
    for s, (a, b) in enumerate(zip(test, ref)):
        gap = abs(a["traj_costs"] - b["traj_costs"]) / max(abs(b["traj_costs"]), 1e-300)
        if flip is None and not same_decisions(a, b):
            flip = s
            prev = ref[s - 1]["traj_costs"] if s else np.inf
            plateau = bool(abs(prev - b["traj_costs"]) <= PLATEAU * abs(b["traj_costs"]) or gap <= PLATEAU)
        if flip is None:
            worst = max(worst, _snap_err(a, b))
        else:
            after = max(after, gap)
    return {"worst": worst, "flip": flip, "plateau": plateau, "after": after}


def compare_traces(test, ref, rtol=RTOL):
    """(worst relative error before any flip, first flipped iteration or None, flip on the plateau)."""
    r = analyse(test, ref)
    return r["worst"], r["flip"], r["plateau"]
