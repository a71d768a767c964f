"""Pins the CPU oracle restatement (oracle/ilqr_oracle.c) against the golden
vectors recorded from the real reference solver (tests/golden/make_golden.py)."""

import numpy as np
import pytest

from tests import common


@pytest.mark.parametrize("case", list(common.CASES))
def test_oracle_matches_reference_golden(case, oracle_libs):
    pb, iters, _ = common.make_case(case)
    golden, golden_derivs = common.load_golden(case)
    tol = common.LOOSE.get(case, common.RTOL)
    factory = lambda: oracle_libs.OracleOptim(pb.model)      # noqa: E731
    for i in range(pb.batch):
        tr = common.trace_single(factory, pb, i, iters)
        worst, flip, plateau = common.compare_traces(tr, golden[i])
        assert worst <= tol, f"{case} problem {i}: relative error {worst:.3e}"
        # decisions may only diverge once the cost has converged to round-off
        assert flip is None or plateau, f"{case} problem {i}: decision flip at iteration {flip} off the plateau"
        d = common.derivatives_single(factory, pb, i)
        for n, g in golden_derivs[i].items():
            err = common.rel_err(d[n], g)
            assert err <= max(tol, 1e-9), f"{case} problem {i}: {n} differs by {err:.3e}"


def test_golden_default_mode_flags():
    """The values SURVEY.md appendix G lists for LAT200/seed2."""
    golden, _ = common.load_golden("lateral_default")
    last = golden[0][-1]
    assert int(last["iterations"]) == 4 and int(last["termination_condition"]) == 2
    assert abs(last["traj_costs"] - 20.637620318244) < 1e-9
    assert abs(golden[0][0]["traj_costs"] - 195.177774308481) < 1e-9


def test_port_matches_reference_extra_cases(oracle_libs, tmp_path):
    """tests/extra.py: `ilr`, `shift`, `dynamics` / `ct_dynamics` (incl. the negative-index wrap),
    sticky state, prev_x / prev_k, the two zoo models without a workload generator and a
    user-defined RK4 + augmented-Lagrangian + end-cost problem, a user-defined problem with
    `lerp_wrap` and `blerp` (2-D array parameter) — the C restatement against the
    vectors recorded from the REAL reference (tests/golden/ref_extra.npz, make_golden_extra.py)."""
    import os

    import numpy as np

    from tests import extra
    from tpl_b200 import _cabi, build, genopt, symext as spx

    want = dict(np.load(os.path.join(common.GOLDEN_DIR, "ref_extra.npz")))
    custom = {extra.CUSTOM: oracle_libs.build_custom(extra.custom_definition(genopt, spx), str(tmp_path)),
              extra.TRACK: oracle_libs.build_custom(extra.track_definition(genopt, spx), str(tmp_path))}

    def make(model):
        if model in custom:
            return oracle_libs.OracleOptim(*custom[model])
        return oracle_libs.OracleOptim(model)

    libs = build.build_zoo([n for n, _ in extra.ZOO])
    zoo_info = {n: _cabi.model_info(_cabi.load(p)) for n, p in libs.items()}
    got = extra.run_single(make, zoo_info)
    assert set(got) == set(want)
    worst, bad = extra.compare(got, want)
    assert not bad, bad[:5]
