"""Problem definitions of the tpl planners and controllers.

Each ``config_*`` function restates, symbol for symbol, one optimal-control
problem of the reference's model zoo
(/root/reference/library/tpl/optim/optimizers.py:12-557) and returns a
:class:`tpl_b200.genopt.Config`.  ``build_optimizers()`` compiles the batched
B200 solvers for all of them and publishes the factories as module globals
under the reference's names (optimizers.py:560-582), e.g.
``optimizers.lateral_profile(batch=4096)``.

States, controls, parameter names and their order are part of the interface
(callers set ``opt.params.<name>``), so they are kept verbatim.
"""

import sympy as sp
from sympy import Symbol as _S

from . import genopt
from . import symext as spx
from .symext import ArraySymbol as _A


def _symbols(names):
    return [_S(n) for n in names.split()]


def _bicycle_penalties(v, delta_dot, phi_dot, a, j, P):
    """The five regularisation terms shared by both MPC models
    (optimizers.py:91-95 and :206-210)."""
    return ((P["min_pdelta_dot"] + P["pdelta_dot"] * v**2) * delta_dot**2
            , (P["min_p_phi_dot"] + P["p_phi_dot"] * v**2) * phi_dot**2
            , P["pa"] * a**2
            , P["pj"] * j**2)


def _box_constraints(delta, a, max_delta, max_acc, min_acc):
    """steer and acceleration limits written as g(x) <= 0
    (optimizers.py:103-107, :217-221)."""
    return [delta - max_delta, -max_delta - delta, a - max_acc, min_acc - a]


def config_trajectory_tracking_mpc():
    """Path-following MPC in path-relative form, X=7, U=2, C=4
    (optimizers.py:12-126).  The arc-length state ``s_r`` indexes the reference
    arrays, so their derivatives become finite differences."""
    x, y, phi, delta, v, s_r, a = _symbols("x y phi delta v s_r a")
    j, delta_dot = _symbols("j delta_dot")

    ref_x, ref_y, ref_phi, ref_k, ref_v = (_A(n) for n in "ref_x ref_y ref_phi ref_k ref_v".split())
    ref_step, l, v_ch, max_delta, max_acc, min_acc, a_offset = _symbols(
        "ref_step l v_ch max_delta max_acc min_acc a_offset")
    pnames = ("pd pv pdelta min_pdelta_dot pdelta_dot min_p_phi_dot p_phi_dot "
              "p_phi p_phi_ref_dot_diff pa pj")
    P = {n: _S(n) for n in pnames.split()}

    def on_path(fn, arr, arg=s_r):
        return fn(0.0, ref_step, arg, arr)

    r_x = on_path(spx.lerp, ref_x)
    r_y = on_path(spx.lerp, ref_y)
    r_phi = on_path(spx.lerp_angle, ref_phi)
    r_k = on_path(spx.lerp, ref_k)
    v_trg = on_path(spx.lerp, ref_v, spx.fixed(s_r))       # not differentiated

    phi_dot = v / (l * (1 + (v / v_ch)**2)) * sp.tan(delta)
    d_r = sp.cos(r_phi) * (y - r_y) - sp.sin(r_phi) * (x - r_x)
    s_dot = v * sp.cos(phi - r_phi) / (1 - d_r * r_k)

    f = sp.Matrix([v * sp.cos(phi), v * sp.sin(phi), phi_dot, delta_dot,
                   a + a_offset, s_dot, j])

    cost = 0.0
    for term in _bicycle_penalties(v, delta_dot, phi_dot, a, j, P):
        cost += term
    cost += P["pv"] * (v - v_trg)**2
    cost += P["pd"] * d_r**2
    cost += P["p_phi"] * (1.0 - sp.cos(phi - r_phi))
    cost += P["p_phi_ref_dot_diff"] * (phi_dot - s_dot * r_k)**2 * v**2

    params = ([P[n] for n in pnames.split()]
              + [l, v_ch, ref_x, ref_y, ref_phi, ref_k, ref_v, ref_step,
                 max_delta, max_acc, min_acc, a_offset])

    return genopt.Config([x, y, phi, delta, v, s_r, a], [j, delta_dot], params,
                         f, cost, end_costs=0.0,
                         constraints=_box_constraints(delta, a, max_delta, max_acc, min_acc))


def config_trajectory_tracking_mpc_time():
    """Time-indexed trajectory-tracking MPC, X=6, U=2, C=4
    (optimizers.py:129-240): kinematic bicycle with centre-of-gravity slip."""
    t, dt = _symbols("t dt")
    x, y, phi, delta, v, a = _symbols("x y phi delta v a")
    j, delta_dot = _symbols("j delta_dot")

    ref_x, ref_y, ref_phi, ref_v = (_A(n) for n in "ref_x ref_y ref_phi ref_v".split())
    ref_dt, ref_t_offset, l, v_ch, max_delta, max_acc, min_acc, a_offset, cog_pos = _symbols(
        "ref_dt ref_t_offset l v_ch max_delta max_acc min_acc a_offset cog_pos")
    pnames = "pd pv pdelta min_pdelta_dot pdelta_dot min_p_phi_dot p_phi_dot p_phi pa pj"
    P = {n: _S(n) for n in pnames.split()}

    rt = ref_t_offset + dt * t

    def at_time(fn, arr):
        return fn(0.0, ref_dt, rt, arr)

    r_x, r_y, v_trg = (at_time(spx.lerp, arr) for arr in (ref_x, ref_y, ref_v))
    r_phi = at_time(spx.lerp_angle, ref_phi)

    beta = sp.atan(sp.tan(delta) * cog_pos)
    phi_dot = v * sp.tan(delta) * sp.cos(beta) / (l * (1 + (v / v_ch)**2))

    f = sp.Matrix([v * sp.cos(phi + beta), v * sp.sin(phi + beta), phi_dot,
                   delta_dot, a + a_offset, j])

    cost = 0.0
    for term in _bicycle_penalties(v, delta_dot, phi_dot, a, j, P):
        cost += term
    cost += P["pv"] * (v - v_trg)**2
    cost += P["pd"] * (x - r_x)**2 + P["pd"] * (y - r_y)**2
    cost += P["p_phi"] * (1.0 - sp.cos(phi - r_phi))

    params = ([P[n] for n in pnames.split()]
              + [l, v_ch, cog_pos, ref_x, ref_y, ref_phi, ref_v, ref_dt, ref_t_offset,
                 max_delta, max_acc, min_acc, a_offset])

    return genopt.Config([x, y, phi, delta, v, a], [j, delta_dot], params,
                         f, cost, end_costs=0.0,
                         constraints=_box_constraints(delta, a, max_delta, max_acc, min_acc))


def config_lateral_profile():
    """Lateral offset profile over arc length (the RSTP path step), X=2, U=1, C=2
    (optimizers.py:243-294)."""
    idx, ds = _S("t"), _S("dt")
    d, v_d, a_d = _symbols("d v_d a_d")
    k_ref, d_offset, d_lower_constr, d_upper_constr = (
        _A(n) for n in "k_ref d_offset d_lower_constr d_upper_constr".split())
    ref_step, w_d, w_v_d, w_a_d, w_k = _symbols("ref_step w_d w_v_d w_a_d w_k")

    s = idx * ds
    k_r, d_o, d_lower, d_upper = (spx.lerp(0.0, ref_step, s, arr)
                                  for arr in (k_ref, d_offset, d_lower_constr, d_upper_constr))

    # curvature of the Cartesian path expressed in Frenet coordinates
    k = (a_d / (v_d**2 + 1) + k_r) * sp.cos(sp.atan(v_d)) / (1 - d * k_r)

    cost = w_d * (d - d_o)**2 + w_v_d * v_d**2 + w_a_d * a_d**2 + w_k * k**2
    end_cost = w_d * (d - d_o)**2 + w_v_d * v_d**2

    return genopt.Config(
        [d, v_d], [a_d],
        [k_ref, d_offset, d_lower_constr, d_upper_constr, ref_step, w_d, w_v_d, w_a_d, w_k],
        sp.Matrix([v_d, a_d]), cost, end_costs=end_cost,
        constraints=[d_lower - d, d - d_upper])


def config_velocity_profile_time():
    """Velocity profile over time with station bounds per step, X=2, U=1, C=4
    (optimizers.py:297-349; not part of ``build_optimizers``)."""
    t = _S("t")
    s, v, a = _symbols("s v a")
    w_v, w_a, ref_step = _symbols("w_v w_a ref_step")
    ref_v, ref_s_max, ref_s_min = (_A(n) for n in "ref_v ref_s_max ref_s_min".split())

    v_max = spx.lerp(0.0, ref_step, s, ref_v)
    s_max = spx.get_array_value(ref_s_max, t)
    s_min = spx.get_array_value(ref_s_min, t)

    constraints = [
        0.0 - v,
        v - v_max,
        sp.Piecewise((s - s_max, s_max > 0), (0.0, True)),
        sp.Piecewise((s_min - s, s_min > 0), (0.0, True)),
    ]
    cost = sp.Matrix([w_v * (1000 - v) + w_a * a**2])

    return genopt.Config([s, v], [a], [w_v, w_a, ref_v, ref_step, ref_s_max, ref_s_min],
                         sp.Matrix([v, a]), cost, end_costs=0.0, constraints=constraints)


def config_velocity_profile_space():
    """Velocity profile over arc length (the RSTP speed step), X=2, U=1, C=5
    (optimizers.py:352-428)."""
    t, dt = _symbols("t dt")
    st, v, a = _symbols("st v a")
    ref_step, p_v, p_a, max_a_total = _symbols("ref_step p_v p_a max_a_total")
    ref_t_offset, ref_v, ref_k, ref_t_max, ref_t_min, ref_v_weight = (
        _A(n) for n in "ref_t_offset ref_v ref_k ref_t_max ref_t_min ref_v_weight".split())

    s = t * dt
    t_offset = spx.box_interp(ref_step, s, ref_t_offset)
    moving = v > 1.0 + 1e-3

    f = sp.Matrix([
        sp.Piecewise((a / v, moving), (a, True)),
        sp.Piecewise((1.0 / v, moving), (t_offset, True)),
    ])

    v_trg = spx.lerp(0.0, ref_step, s, ref_v)
    kk = spx.box_interp(ref_step, s, ref_k)
    t_min = spx.lerp(0.0, ref_step, s, ref_t_min)
    t_max = spx.lerp(0.0, ref_step, s, ref_t_max)
    v_weight = spx.lerp(0.0, ref_step, s, ref_v_weight)

    a_lat = v**2 * kk
    constraints = [
        (a**2 + a_lat**2) - max_a_total**2,                                  # friction circle
        1.0 - v,                                                             # v_min
        v - v_trg,                                                           # v_max
        (st + t_offset) - t_max,                                             # latest arrival
        (t_min - st) * sp.Piecewise((v - 1.0, t_min > 0.0), (1.0, True)),    # earliest arrival
    ]
    cost = sp.Matrix([p_v * (v_trg - v)**2 * v_weight + p_a * a**2])

    return genopt.Config(
        [v, st], [a],
        [p_v, p_a, max_a_total, ref_v, ref_k, ref_step, ref_t_max, ref_t_min,
         ref_t_offset, ref_v_weight],
        f, cost, end_costs=0.0, constraints=constraints)


def _smoother(with_dk):
    """Reference-line smoothers (optimizers.py:431-490 and :493-557)."""
    t, dt = _symbols("t dt")
    x, y, phi, k, dk = _symbols("x y phi k dk")
    w_pos, w_k, w_dk, s_start, ref_step = _symbols("w_pos w_k w_dk s_start ref_step")
    ref_x, ref_y = _A("ref_x"), _A("ref_y")

    s = (s_start + t * dt) if with_dk else t * dt
    x_ref = spx.lerp(0.0, ref_step, s, ref_x)
    y_ref = spx.lerp(0.0, ref_step, s, ref_y)
    cost = w_pos * (x - x_ref)**2 + w_pos * (y - y_ref)**2 + w_k * k**2

    if with_dk:
        cost = cost + w_dk * dk**2
        return genopt.Config([x, y, phi, k], [dk],
                             [w_pos, w_k, w_dk, s_start, ref_x, ref_y, ref_step],
                             sp.Matrix([sp.cos(phi), sp.sin(phi), k, dk]), cost, end_costs=0.0)
    return genopt.Config([x, y, phi], [k], [w_pos, w_k, ref_x, ref_y, ref_step],
                         sp.Matrix([sp.cos(phi), sp.sin(phi), k]), cost, end_costs=0.0)


def config_ref_line_smoother_k():
    """Curvature-controlled reference-line smoother, X=3, U=1, C=0."""
    return _smoother(with_dk=False)


def config_ref_line_smoother_dk():
    """Curvature-rate-controlled reference-line smoother, X=4, U=1, C=0."""
    return _smoother(with_dk=True)


#: every problem the reference ships, by the name its factory is published under
CONFIGS = {
    "trajectory_tracking_mpc": config_trajectory_tracking_mpc,
    "trajectory_tracking_mpc_time": config_trajectory_tracking_mpc_time,
    "lateral_profile": config_lateral_profile,
    "velocity_profile_space": config_velocity_profile_space,
    "ref_line_smoother_k": config_ref_line_smoother_k,
    "ref_line_smoother_dk": config_ref_line_smoother_dk,
    "velocity_profile_time": config_velocity_profile_time,
}

#: the six that ``build_optimizers`` builds (optimizers.py:562-569)
DEFAULT_BUILD = tuple(n for n in CONFIGS if n != "velocity_profile_time")


def build_optimizers(force_rebuild=False, names=DEFAULT_BUILD):
    """Compile the batched solvers and publish ``optimizers.<name>`` factories
    (optimizers.py:560-582)."""
    todo = [n for n in names if force_rebuild or n not in globals()]
    if not todo:
        return
    factories = genopt.build_parallel([CONFIGS[n]() for n in todo], names=todo,
                                      force=force_rebuild)
    globals().update(dict(zip(todo, factories)))
