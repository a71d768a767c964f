"""Symbolic derivation of everything the iLQR solver needs from a problem.

This is the B200 build's restatement of the reference's derivation pipeline
(/root/reference/library/tpl/optim/genopt.py:65-179 and :544-569):

* stage cost augmented with the augmented-Lagrangian / gated-penalty terms
  (genopt.py:73-90),
* Jacobians of the *Euler* step ``x + dt f(x,u)`` — whatever integrator rolls
  the trajectory out (genopt.py:93-110),
* gradient / Hessian blocks of the augmented stage cost and of the end cost
  (genopt.py:113-172), with every ``Derivative(opaque_fn, ...)`` replaced by a
  central finite difference of step 1e-4 (genopt.py:65-70),
* the raw constraint vector (genopt.py:175-179).

The result is a dictionary of sympy matrices over a fixed set of leaf symbols
(``x[i]``, ``u[i]``, ``lg_mult[i]``, ``lg_weight[i]``, ``t``, ``dt`` and the
problem's parameters), ready for the printers in ``codegen``.
"""

from dataclasses import dataclass, field
from typing import Dict, List

import sympy as sp

from . import symext as spx

# the reference writes the step and the gate threshold as ``10e-5``
FD_STEP = 10e-5          # genopt.py:65
MULT_GATE = 10e-5        # genopt.py:81


def _as_matrix(e):
    return e if isinstance(e, sp.MatrixBase) else sp.Matrix([e])


def finite_difference_derivatives(expr, step=FD_STEP):
    """Replace each ``Derivative`` atom by sympy's central finite difference
    (genopt.py:65-70)."""
    for d in expr.atoms(sp.Derivative):
        expr = expr.subs(d, d.as_finite_difference(step))
    return expr


def augment_stage_cost(cost, constraints):
    """l + sum_c [ lambda_c g_c + gate(g_c, lambda_c) w_c g_c^2 ]  (genopt.py:73-90).

    The quadratic term is switched off only while the constraint is inactive
    (g < 0) *and* its multiplier is (numerically) zero."""
    cost = _as_matrix(cost)
    for ci, g in enumerate(constraints):
        lam = sp.Symbol(f"lg_mult[{ci}]")
        w = sp.Symbol(f"lg_weight[{ci}]")
        cost = cost + sp.Matrix([g * lam])
        cost = cost + sp.Matrix([
            sp.Piecewise((0.0, (g < 0) & (sp.Abs(lam) < MULT_GATE)),
                         (w * g**2, True))])
    return cost


@dataclass
class Derivation:
    """Everything derived from one problem definition."""
    state_names: List[str]
    action_names: List[str]
    n_constraints: int
    scalar_params: List[str]
    array_params: List[str]
    param_order: List[str]                 # declaration order (scalars and arrays mixed)
    routines: Dict[str, sp.MatrixBase] = field(default_factory=dict)

    @property
    def X(self):
        return len(self.state_names)

    @property
    def U(self):
        return len(self.action_names)

    @property
    def C(self):
        return self.n_constraints


# routine name -> (rows, cols) description, in the reference's order (genopt.py:105-110,141-148,168-172,177-179)
ROUTINE_ORDER = (
    "ctDynamics", "stateJacobian", "actionJacobian",
    "costs", "stateGradient", "actionGradient",
    "stateStateHessian", "actionActionHessian", "actionStateHessian",
    "endCosts", "endGradient", "endHessian",
    "constraints",
)


def _vectorise(states, actions, m):
    """x_i -> ``x[i]``, u_i -> ``u[i]`` (genopt.py:51-62), simultaneous."""
    repl = {s: sp.Symbol(f"x[{i}]") for i, s in enumerate(states)}
    repl.update({a: sp.Symbol(f"u[{i}]") for i, a in enumerate(actions)})
    return m.xreplace(repl)


def derive(config) -> Derivation:
    """Run the whole derivation for a :class:`tpl_b200.genopt.Config`."""
    states = list(config.states)
    actions = list(config.actions)
    params = list(config.params.keys()) if isinstance(config.params, dict) else list(config.params)
    constraints = list(config.constraints)
    X = len(states)

    f = _as_matrix(config.dynamics)
    stage_cost = augment_stage_cost(_as_matrix(config.costs), constraints)
    end_cost = _as_matrix(config.end_costs)      # NOT augmented (genopt.py:552)

    r = {}

    # -- dynamics (genopt.py:93-110)
    dt = sp.Symbol("dt")
    euler = sp.Matrix([states[i] + dt * f[i] for i in range(X)])
    jac = finite_difference_derivatives(euler.jacobian(states + actions))
    r["ctDynamics"] = f
    r["stateJacobian"] = jac[:, :X]
    r["actionJacobian"] = jac[:, X:]

    # -- stage cost (genopt.py:113-148)
    var = states + actions
    grad = finite_difference_derivatives(stage_cost.jacobian(var)).T
    hess = sp.hessian(stage_cost, var)
    if hess != hess.T:
        raise RuntimeError("Detected non-symmetric Hessian!")
    hess = finite_difference_derivatives(finite_difference_derivatives(hess))
    r["costs"] = stage_cost
    r["stateGradient"] = grad[:X, :]
    r["actionGradient"] = grad[X:, :]
    r["stateStateHessian"] = hess[:X, :X]
    r["actionActionHessian"] = hess[X:, X:]
    r["actionStateHessian"] = hess[X:, :X]

    # -- end cost (genopt.py:151-172)
    egrad = finite_difference_derivatives(end_cost.jacobian(states)).T
    ehess = sp.hessian(end_cost, states)
    if ehess != ehess.T:
        raise RuntimeError("detected non-symmetric hessian")
    ehess = finite_difference_derivatives(finite_difference_derivatives(ehess))
    r["endCosts"] = end_cost
    r["endGradient"] = egrad
    r["endHessian"] = ehess

    # -- constraints (genopt.py:175-179)
    r["constraints"] = sp.Matrix(constraints) if constraints else sp.zeros(0, 1)

    out = {}
    for name in ROUTINE_ORDER:
        m = r[name]
        if m.shape[0] * m.shape[1] > 0:
            m = spx.unfixed(m)                       # genopt.py:557-558
            m = _vectorise(states, actions, m)       # genopt.py:568-569
        out[name] = m

    return Derivation(
        state_names=[s.name for s in states],
        action_names=[a.name for a in actions],
        n_constraints=len(constraints),
        scalar_params=[p.name for p in params if not isinstance(p, spx.ArraySymbol)],
        array_params=[p.name for p in params if isinstance(p, spx.ArraySymbol)],
        param_order=[p.name for p in params],
        routines=out,
    )
