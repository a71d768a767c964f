"""Host-side mirror of the reference `Optim` interface (tpl_b200/batched.py): shapes,
construction defaults, setters, horizon clamp, deepcopy / __getstate__ — exercised
on CPU buffers.  Compute entry points must refuse to run without CUDA."""

import copy
import os

import numpy as np
import pytest
import torch

from tests import common
from tpl_b200 import _cabi, scenarios as sc
from tpl_b200.batched import BatchedOptim


@pytest.fixture()
def lateral(solver_libs):
    return BatchedOptim(solver_libs["lateral_profile"], batch=4, horizon_max=60, device="cpu")


@pytest.fixture()
def mpc_time(solver_libs):
    return BatchedOptim(solver_libs["trajectory_tracking_mpc_time"], batch=3, scenes=1,
                        horizon_max=50, device="cpu")


def test_construction_defaults(lateral):
    """optim.c:1894-1921, SURVEY.md appendix G11."""
    o = lateral
    assert (o.dt, o.horizon, o.max_iterations, o.max_lg_iterations) == (0.05, 20, 5, 1)
    assert o.min_rel_cost_change == 1e-6 and o.use_quadratic_terms and o.opt_start == 0
    assert bool(torch.isinf(o.lg_mult_limit).all()) and bool((o.barrier_weight == 1).all())
    o.horizon = 40
    assert o.u_max[0, 18:23].tolist() == [float("inf"), float("inf"), 0.0, 0.0, 0.0]
    assert o.u_min[0, 18:23].tolist() == [-float("inf"), -float("inf"), 0.0, 0.0, 0.0]
    assert bool((o.x == 0).all()) and bool((o.lagrange_multiplier == 0).all())


def test_shapes_are_squeezed_like_the_reference(lateral, mpc_time):
    """optim.c:1314-1325: U == 1 gives u of shape (T,) per problem."""
    o = lateral
    o.horizon = 40
    assert o.x.shape == (4, 41, 2) and o.u.shape == (4, 40) and o.K.shape == (4, 40, 2)
    assert o.fx.shape == (4, 40, 2, 2) and o.fu.shape == (4, 40, 2) and o.luu.shape == (4, 40)
    assert o.lux.shape == (4, 40, 2) and o.lagrange_multiplier.shape == (4, 40, 2)
    assert o[1].x.shape == (41, 2) and o[1].u.shape == (40,)
    m = mpc_time
    m.horizon = 30
    assert m.x.shape == (3, 31, 6) and m.u.shape == (3, 30, 2) and m.K.shape == (3, 30, 2, 6)
    assert m.fx.shape == (3, 30, 6, 6) and m.fu.shape == (3, 30, 6, 2) and m.lux.shape == (3, 30, 2, 6)
    assert m.barrier_weight.shape == (3, 4) and m.int_step.shape == (3, 31)


def test_horizon_is_clamped(lateral, solver_libs):
    lateral.horizon = 1000                      # optim.c:1726-1734, capacity 60 here
    assert lateral.horizon == 60 and lateral.T == 60
    lateral.horizon = -5
    assert lateral.horizon == 1
    full = BatchedOptim(solver_libs["ref_line_smoother_k"], batch=1, device="cpu")
    full.horizon = 1000
    assert full.horizon == 299
    assert full.lagrange_multiplier.shape == (1, 299, 0)          # C == 0 model


def test_setters_broadcast_and_write_through(lateral):
    o = lateral
    o.horizon = 30
    o.u_min = -2.5                              # scalar broadcast (path_optim.py:129)
    o.u_max[:, :5] = 0.0                        # in-place write through a view (path_optim.py:164)
    assert bool((o.u_min == -2.5).all()) and bool((o.u_max[:, :5] == 0).all())
    o.u = np.arange(30.0)                       # (T,) broadcast over the batch (velocity_optim.py:170)
    assert torch.equal(o.u[2], torch.arange(30.0, dtype=torch.float64))
    o.x[:, 0] = torch.tensor([[1.0, 2.0]], dtype=torch.float64)
    assert o[3].x[0].tolist() == [1.0, 2.0]
    o[1].u = np.ones(30)
    assert bool((o.u[1] == 1).all()) and bool((o.u[0] == torch.arange(30.0, dtype=torch.float64)).all())
    o.lg_mult_limit = 0.0                       # path_optim.py:102
    o.barrier_weight[:] = 1000.0                # path_optim.py:103
    assert bool((o.lg_mult_limit == 0).all()) and bool((o.barrier_weight == 1000).all())
    o.mu_step = 3
    assert o.mu_step.tolist() == [3, 3, 3, 3]
    with pytest.raises(AttributeError):
        o[0].horizon = 5


def test_params_scalars_and_arrays(mpc_time):
    p = mpc_time.params
    assert p.slots[:3] == ["pd", "pv", "pdelta"] and "ref_x" in p.slots
    p.pd = 5.0
    p.ref_x = np.linspace(0.0, 1.0, 80)         # (L,) shared by all scenes
    assert p.pd.tolist() == [5.0] and p.ref_x.shape == (1, 80)
    with pytest.raises(AttributeError):
        p.nonexistent = 1.0
    with pytest.raises(ValueError):
        p.ref_x = np.zeros((2, 80))             # wrong number of scenes
    with pytest.raises(ValueError):
        mpc_time.scene_index = np.array([0, 1, 0])   # scene 1 does not exist
    state = mpc_time.__getstate__()
    assert sorted(state) == sorted(["u_min", "u_max", "horizon", "opt_start", "barrier_weight",
                                    "lg_mult_limit", "max_iterations", "max_lg_iterations", "step",
                                    "use_quadratic_terms", "params"])     # optim.c:1806-1819
    assert state["params"]["pd"].tolist() == [5.0]


def test_deepcopy_is_independent(lateral):
    pb = sc.lateral(4, horizon=40)
    o = sc.apply_to_batched(lateral, pb)
    c = copy.deepcopy(o)
    assert torch.equal(c.x, o.x) and torch.equal(c.params.k_ref, o.params.k_ref) and c.horizon == o.horizon
    c.x[:, 0] = 7.0
    c.params.w_d = 9.0
    assert not torch.equal(c.x, o.x) and float(o.params.w_d[0]) == 0.2


def test_slots_match_reference(lateral):
    assert lateral.slots[:6] == ["step", "opt_start", "horizon", "int_step", "x", "u"]
    assert "min_rel_cost_change" in lateral.slots and len(lateral.slots) == 42


def test_no_cpu_fallback(lateral, solver_libs):
    for call in (lateral.update, lateral.linearize, lambda: lateral.shift(1), lambda: lateral.shift_interp(0.5),
                 lambda: lateral.dynamics(np.zeros(2), np.zeros(1), 0, 0.1),
                 lambda: lateral.argmin_groups(2)):
        with pytest.raises(_cabi.SolverError):
            call()
    if not torch.cuda.is_available():
        with pytest.raises(_cabi.SolverError):
            BatchedOptim(solver_libs["lateral_profile"], batch=2)
    with pytest.raises(OSError):
        _cabi.load("/nonexistent/libtplb200_x.so")


def test_scenario_batches_are_seeded(solver_libs):
    a, b = sc.mpc_time(8, horizon=20), sc.mpc_time(8, horizon=20)
    assert np.array_equal(a.x0, b.x0) and np.array_equal(a.arrays["ref_x"], b.arrays["ref_x"])
    sub = a.subset([5, 2])
    assert sub.batch == 2 and np.array_equal(sub.x0[0], a.x0[5]) and sub.scenes == 2
    ms = sc.mpc_time(8, scenes=2, horizon=20)
    assert ms.scenes == 2 and ms.scene_of.tolist() == [0, 0, 0, 0, 1, 1, 1, 1]


def test_pipeline_needs_cuda(lateral):
    from tpl_b200.streaming import SolverPipeline
    with pytest.raises(ValueError):
        SolverPipeline(lambda: lateral, depth=0)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            SolverPipeline(lambda: lateral, depth=2)


def test_generated_headers_are_current():
    """The committed model headers are what the code generator emits today (a change to
    codegen.py / derive.py / symext.py must be followed by `python -m tpl_b200.build --regen`)."""
    from tpl_b200 import build, codegen, derive, optimizers
    gen = build._generator_hash()
    for name, make in optimizers.CONFIGS.items():
        cfg = make()
        text = codegen.emit_cuda_model(derive.derive(cfg), name, cfg.definition_hash())
        first, rest = text.split("\n", 1)
        with open(os.path.join(build.GENERATED, name + ".cuh")) as fd:
            assert fd.read() == first + "\n// generator sha1: " + gen + "\n" + rest, name


def test_two_array_and_2d_lookups_host_side():
    """`lerp_wrap` (two arrays of equal length) and `blerp` (a 2-D array), optim.c:410-486: the
    generator records how each array parameter is read, the generated code passes the array
    indices in the reference's argument order, and `params` takes maps as (rows, cols) /
    (S, rows, cols) and describes them to the C ABI as samples per scene + row length."""
    import sympy as sp
    from tests import extra
    from tpl_b200 import codegen, derive, genopt, symext as spx
    cfg = extra.track_definition(genopt, spx)
    d = derive.derive(cfg)
    assert d.array_params == ["track_s", "track_k", "track_v", "slope"]
    assert codegen.array_shapes(d) == ([1, 1, 1, 2], [(0, 1), (0, 2)])
    src = codegen.emit_cuda_model(d, "track", cfg.definition_hash())
    assert "P.lerp_wrap(0, 1, " in src and "P.lerp_wrap(0, 2, " in src and "P.blerp(3, " in src
    assert "static constexpr int ARRAY_NDIM[4] = {1, 1, 1, 2};" in src

    x, u, t, dt = sp.symbols("x u t dt")
    grid = spx.ArraySymbol("grid")
    both = genopt.Config([x], [u], [grid], sp.Matrix([u + spx.lerp(0.0, 1.0, x, grid)]),
                         spx.blerp(0.0, 0.0, 1.0, 1.0, t * dt, 0.0, grid) + u**2)
    with pytest.raises(ValueError, match="both as a 1-D and as a 2-D"):
        codegen.array_shapes(derive.derive(both))

    q = genopt.build(cfg)(batch=6, scenes=3, horizon_max=10, device="cpu")
    q.params.slope = np.arange(12.0).reshape(3, 4)                       # shared by the scenes
    assert tuple(q.params.slope.shape) == (3, 3, 4)
    q.params.slope = np.arange(60.0).reshape(3, 4, 5)
    q.params.track_s = np.arange(7.0)
    desc = q._descriptor()
    assert (desc.array_len[3], desc.array_cols[3]) == (20, 5)
    assert (desc.array_len[0], desc.array_cols[0]) == (7, 0)
    with pytest.raises(ValueError, match="rows, cols"):
        q.params.slope = np.arange(5.0)
    with pytest.raises(ValueError):
        q.params.track_s = np.zeros((2, 3, 4))


def test_every_slot_is_readable():
    """`viz.autogui(opt)` of the reference walks `Optim.__slots__` (optim.c:1736-1782) and reads every
    name: all of them exist here, with the reference's shapes behind a leading batch dimension."""
    from tpl_b200 import build
    lib = build.build_zoo(["trajectory_tracking_mpc_time"])["trajectory_tracking_mpc_time"]
    q = BatchedOptim(lib, batch=3, horizon_max=30, device="cpu")
    q.horizon = 25
    B, T, X, U, C = 3, 25, q.X, q.U, q.C
    want = {"x": (B, T + 1, X), "next_x": (B, T + 1, X), "prev_x": (B, T + 1, X), "u": (B, T, U), "next_u": (B, T, U),
            "prev_u": (B, T, U), "prev_k": (B, T, U), "k": (B, T, U), "g": (B, T, U), "K": (B, T, U, X),
            "int_step": (B, T + 1), "next_int_step": (B, T + 1), "fx": (B, T, X, X), "fu": (B, T, X, U),
            "lx": (B, T, X), "lu": (B, T, U), "lxx": (B, T, X, X), "luu": (B, T, U, U), "lux": (B, T, U, X),
            "lagrange_multiplier": (B, T, C), "barrier_weight": (B, C), "lg_mult_limit": (B, C),
            "u_min": (B, T, U), "u_max": (B, T, U)}
    for name in q.slots:
        value = getattr(q, name)
        if name in want:
            assert tuple(value.shape) == want[name], name
    assert set(want) <= set(q.slots)
    q.prev_u = 1.5                                   # writable like every array attribute
    assert float(q.prev_u.min()) == 1.5
    assert tuple(q[1].next_x.shape) == (T + 1, X) and tuple(q[1].prev_u.shape) == (T, U)


def test_trace_analysis_classifies_flips():
    """tpl_b200.parity.analyse (used by the parity tests and by bench.py's in-run check): agreement is
    measured up to the first iteration whose decisions differ; that flip is legitimate only when the line
    search was decided at round-off level, and the costs must stay together afterwards."""
    from tpl_b200 import parity

    def snap(cost, alpha=1.0, mu_step=0, it=1, x=1.0):
        return {"x": np.full((3, 2), x), "u": np.full((2, 1), x), "traj_costs": cost, "alpha": alpha,
                "mu_step": mu_step, "iterations": it, "termination_condition": 1, "improved": 1,
                "trajectory_changed": 1, "lg_iterations": 1}
    ref = [snap(10.0, it=0), snap(5.0, it=1), snap(5.0 - 1e-12, alpha=0.1, it=2), snap(5.0 - 2e-12, it=3)]
    same = [dict(s) for s in ref]
    r = parity.analyse(same, ref)
    assert r["flip"] is None and r["worst"] == 0.0
    # a different step size at iteration 2, where the reference's cost moved by 2e-13 relative: plateau
    flipped = [dict(s) for s in ref]
    flipped[2] = snap(5.0 - 1.1e-12, alpha=1.0, it=2, x=1.0 + 1e-6)
    flipped[3] = snap(5.0 - 2.1e-12, it=3, x=1.0 + 1e-6)
    r = parity.analyse(flipped, ref)
    assert r["flip"] == 2 and r["plateau"] and r["worst"] == 0.0 and r["after"] < 1e-12
    # the same flip while the cost is still falling: not a plateau
    ref2 = [snap(10.0, it=0), snap(5.0, it=1), snap(4.0, alpha=0.1, it=2)]
    bad = [dict(s) for s in ref2]
    bad[2] = snap(4.5, alpha=1.0, it=2)
    r = parity.analyse(bad, ref2)
    assert r["flip"] == 2 and not r["plateau"]
    # a numerical difference before any flip is reported as `worst`
    off = [dict(s) for s in ref]
    off[1] = snap(5.0 * (1 + 1e-7), it=1)
    assert parity.analyse(off, ref)["worst"] > 1e-8
