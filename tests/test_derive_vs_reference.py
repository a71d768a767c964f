"""The derivation pipeline (tpl_b200/derive.py) against the reference's own generator:
every routine expression must be string-identical.  Needs the reference tree, so it runs
in the build container only (skipped on the GPU box)."""

import os
import sys

import pytest
import sympy as sp

REF = "/root/reference/library"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def _reference_routines(cfg):
    from tpl.optim import genopt as rg, symext as rspx
    if not isinstance(cfg.costs, sp.Matrix):
        cfg.costs = sp.Matrix([cfg.costs])
    if not isinstance(cfg.end_costs, sp.Matrix):
        cfg.end_costs = sp.Matrix([cfg.end_costs])
    costs = rg.augment_costs(cfg.costs, cfg.constraints)                      # genopt.py:544
    routines = rg.gen_dynamics_routines(cfg.states, cfg.actions, cfg.dynamics)
    routines += rg.gen_cost_routines(cfg.states, cfg.actions, costs)
    routines += rg.gen_end_cost_routines(cfg.states, cfg.end_costs)
    if cfg.constraints:
        routines += rg.gen_constraint_routines(cfg.constraints)
    repl = {s: sp.Symbol(f"x[{i}]") for i, s in enumerate(cfg.states)}
    repl.update({s: sp.Symbol(f"u[{i}]") for i, s in enumerate(cfg.actions)})
    return {n: rspx.unfixed(r).xreplace(repl) for n, r in routines}


@pytest.mark.parametrize("name", ["trajectory_tracking_mpc_time", "lateral_profile", "velocity_profile_space",
                                  "ref_line_smoother_k", "ref_line_smoother_dk", "velocity_profile_time",
                                  "trajectory_tracking_mpc"])
def test_expressions_identical_to_reference_generator(name):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from tpl.optim import optimizers as ro
    from tpl_b200 import derive, optimizers as mo
    ref_cfg = getattr(ro, "config_" + name)()
    mine = derive.derive(mo.CONFIGS[name]())
    assert [p.name for p in ref_cfg.params] == mine.param_order
    for rname, expr in _reference_routines(ref_cfg).items():
        if rname == "jacobian":
            continue
        assert str(expr) == str(mine.routines[rname]), f"{name}.{rname}"
