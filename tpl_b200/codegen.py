"""sympy -> CUDA device code (and plain C for the test oracle).

The reference prints one C function per derivative routine with sympy's
``codegen(cse=True)`` and a patched ``C99CodePrinter``
(/root/reference/library/tpl/optim/genopt.py:182-318, :573-582).  The B200 build
prints *fused* groups instead: all seven linearisation outputs of a stage
(``fx, fu, lx, lu, lxx, luu, lux``) share one common-subexpression pass, so the
trigonometric and interpolation work the reference repeats in every routine
(SURVEY.md §7 "hard parts") is evaluated once per stage.  Every expression is
still the reference's expression; only sharing differs.

Two dialects:

* ``cuda``  — ``template <typename R, typename PV> __device__`` members of a
  ``struct Model`` (``R`` = real type, ``PV`` = parameter view giving scalars and
  interpolation lookups), consumed by ``csrc/solver.cuh``.
* ``c``     — ``static`` C99 functions over ``double``, one per reference
  routine (no fusion), consumed by the CPU oracle in ``oracle/``.
"""

import re
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import sympy as sp
from sympy.printing.c import C99CodePrinter
from sympy.printing.precedence import precedence

from . import symext as spx
from .derive import Derivation

_OPAQUE_NAMES = {
    spx.lerp: "lerp",
    spx.lerp_angle: "lerp_angle",
    spx.box_interp: "box_interp",
    spx.get_array_value: "array_value",
    spx.blerp: "blerp",
    spx.lerp_wrap: "lerp_wrap",
}


@dataclass
class Dialect:
    name: str
    literal: str            # format of a floating literal, e.g. "R(%s)"
    opaque_call: str        # format: fn, array index, args
    scalar_load: str        # format: index
    real: str


CUDA = Dialect("cuda", "R(%s)", "P.%s(%d%s)", "P.scalar(%d)", "R")
PLAIN_C = Dialect("c", "%s", "tplo_%s(P, %d%s)", "P->scalar[%d]", "double")


class ModelPrinter(C99CodePrinter):
    """C99 printer with the reference's ``Pow`` shortcuts (genopt.py:188-215)
    made single-evaluation (``sq(a)`` instead of ``(a*a)``), typed literals, and
    parameter / interpolation access through the dialect."""

    def __init__(self, deriv: Derivation, dialect: Dialect):
        super().__init__({"allow_unknown_functions": True})
        self.d = deriv
        self.dialect = dialect
        self.scalar_index = {n: i for i, n in enumerate(deriv.scalar_params)}
        self.array_index = {n: i for i, n in enumerate(deriv.array_params)}
        self.used_scalars = set()

    # -- leaves --------------------------------------------------------------
    def _print_Symbol(self, s):
        if s.name in self.scalar_index:
            self.used_scalars.add(s.name)
            return "p_" + s.name
        if s.name in self.array_index:
            raise ValueError(f"array parameter {s.name} used outside an interpolation function")
        return s.name

    def _print_ArraySymbol(self, s):
        return self._print_Symbol(s)

    def _lit(self, text):
        return self.dialect.literal % text

    def _print_Float(self, f):
        return self._lit(super()._print_Float(f))

    def _print_Rational(self, r):
        return self._lit("%d.0/%d.0" % (r.p, r.q))

    # -- powers ----------------------------------------------------------------
    def _print_Pow(self, e):
        b, x = e.base, e.exp
        one = self._lit("1.0")
        pb = self._print(b)
        if x == -1:
            return "%s/%s" % (one, self.parenthesize(b, precedence(e)))
        if x == 2:
            return "sq(%s)" % pb
        if x == -2:
            return "%s/sq(%s)" % (one, pb)
        if x == sp.Rational(1, 2) or x == 0.5:
            return "sqrt(%s)" % pb
        if x == sp.Rational(-1, 2) or x == -0.5:
            return "%s/sqrt(%s)" % (one, pb)
        if x == sp.Rational(3, 2) or x == 1.5:
            return "pow3h(%s)" % pb
        if x == sp.Rational(-3, 2) or x == -1.5:
            return "%s/pow3h(%s)" % (one, pb)
        if x.is_Integer and 2 < int(x) <= 8:
            return "ipow%d(%s)" % (int(x), pb)
        if x.is_Integer and -8 <= int(x) < -2:
            return "%s/ipow%d(%s)" % (one, -int(x), pb)
        return "pow(%s, %s)" % (pb, self._print(x))

    # -- opaque lookups ---------------------------------------------------------
    def _print_Function(self, e):
        fn = _OPAQUE_NAMES.get(type(e))
        if fn is None:
            return super()._print_Function(e)
        arrays = [a for a in e.args if isinstance(a, spx.ArraySymbol)]
        others = [a for a in e.args if not isinstance(a, spx.ArraySymbol)]
        if len(arrays) != 1:
            raise NotImplementedError(f"{fn} with {len(arrays)} array arguments is not supported")
        idx = self.array_index[arrays[0].name]
        args = "".join(", " + self._print(a) for a in others)
        return self.dialect.opaque_call % (fn, idx, args)


def _flatten(m):
    return [m[i, j] for i in range(m.shape[0]) for j in range(m.shape[1])]


def entry_kinds(m) -> List[int]:
    """Structure of a matrix: 0 = identically zero, 1 = identically one,
    2 = anything else.  Lets the solver skip work on known entries."""
    kinds = []
    for e in _flatten(m):
        if e == 0:
            kinds.append(0)
        elif e == 1:
            kinds.append(1)
        else:
            kinds.append(2)
    return kinds


def _is_boolean(expr):
    """True for relational / logical sub-expressions hoisted by the CSE pass."""
    from sympy.logic.boolalg import Boolean
    if isinstance(expr, Boolean) or getattr(expr, "is_Relational", False):
        return True
    if isinstance(expr, sp.Piecewise):
        return all(_is_boolean(e) for e, _ in expr.args)
    return False


def _one_line(text):
    """sympy prints ternaries over several lines; fold them."""
    return re.sub(r"\s*\n\s*", " ", text)


def print_body(printer: ModelPrinter, outputs: Sequence[Tuple[str, sp.MatrixBase]],
               indent="    ") -> str:
    """One CSE pass over all ``outputs`` and the statements that evaluate them."""
    real = printer.dialect.real
    flat, spans = [], []
    for name, m in outputs:
        es = _flatten(m)
        spans.append((name, len(flat), len(es)))
        flat.extend(es)
    printer.used_scalars = set()
    if flat:
        common, reduced = sp.cse(flat, symbols=sp.numbered_symbols("v"))
    else:
        common, reduced = [], []
    lines = []
    for sym, expr in common:
        ctype = "bool" if _is_boolean(expr) else real
        lines.append(f"{indent}const {ctype} {sym} = {_one_line(printer.doprint(expr))};")
    for name, start, n in spans:
        for k in range(n):
            lines.append(f"{indent}{name}[{k}] = {_one_line(printer.doprint(reduced[start + k]))};")
    loads = [f"{indent}const {real} p_{n} = {printer.dialect.scalar_load % printer.scalar_index[n]};"
             for n in printer.d.scalar_params if n in printer.used_scalars]
    return "\n".join(loads + lines)


# ---------------------------------------------------------------------------------
# CUDA model header
# ---------------------------------------------------------------------------------

_STAGE_ARGS = "const R* x, const R* u, const R* lg_mult, const R* lg_weight, const R t, const R dt"
_DYN_ARGS = "const R* x, const R* u, const R t, const R dt"
_END_ARGS = "const R* x, const R t, const R dt"


def _cuda_fn(name, args, outs, body):
    outs_decl = "".join(f", R* {o}" for o in outs)
    return (f"    template <typename R, typename PV>\n"
            f"    __device__ __forceinline__ static void {name}(const PV& P, {args}{outs_decl}) {{\n"
            f"{body}\n"
            f"    }}\n")


def _c_array(ctype, name, values, fmt="%d"):
    vals = ", ".join(fmt % v for v in values) if values else "0"
    n = max(1, len(values))
    return f"    static constexpr {ctype} {name}[{n}] = {{{vals}}};\n"


def emit_cuda_model(d: Derivation, name: str, definition_hash: str) -> str:
    """The ``struct Model`` consumed by csrc/solver.cuh."""
    r = d.routines
    pr = ModelPrinter(d, CUDA)
    ind = "        "
    parts = []
    parts.append(_cuda_fn("ct_dynamics", _DYN_ARGS, ["f"],
                          print_body(pr, [("f", r["ctDynamics"])], ind)))
    parts.append(_cuda_fn("dynamics_jacobians", _DYN_ARGS, ["fx", "fu"],
                          print_body(pr, [("fx", r["stateJacobian"]), ("fu", r["actionJacobian"])], ind)))
    parts.append(_cuda_fn("stage_cost", _STAGE_ARGS, ["c"],
                          print_body(pr, [("c", r["costs"])], ind)))
    parts.append(_cuda_fn("cost_derivatives", _STAGE_ARGS, ["lx", "lu", "lxx", "luu", "lux"],
                          print_body(pr, [("lx", r["stateGradient"]), ("lu", r["actionGradient"]),
                                          ("lxx", r["stateStateHessian"]),
                                          ("luu", r["actionActionHessian"]),
                                          ("lux", r["actionStateHessian"])], ind)))
    parts.append(_cuda_fn("cost_gradients", _STAGE_ARGS, ["lx", "lu"],
                          print_body(pr, [("lx", r["stateGradient"]), ("lu", r["actionGradient"])], ind)))
    parts.append(_cuda_fn("linearize", _STAGE_ARGS, ["fx", "fu", "lx", "lu", "lxx", "luu", "lux"],
                          print_body(pr, [("fx", r["stateJacobian"]), ("fu", r["actionJacobian"]),
                                          ("lx", r["stateGradient"]), ("lu", r["actionGradient"]),
                                          ("lxx", r["stateStateHessian"]),
                                          ("luu", r["actionActionHessian"]),
                                          ("lux", r["actionStateHessian"])], ind)))
    parts.append(_cuda_fn("end_cost", _END_ARGS, ["c"],
                          print_body(pr, [("c", r["endCosts"])], ind)))
    parts.append(_cuda_fn("end_derivatives", _END_ARGS, ["vx", "vxx"],
                          print_body(pr, [("vx", r["endGradient"]), ("vxx", r["endHessian"])], ind)))
    parts.append(_cuda_fn("constraints", _STAGE_ARGS, ["g"],
                          print_body(pr, [("g", r["constraints"])], ind)))

    def names(tag, items):
        body = ", ".join(f'"{s}"' for s in items) if items else '""'
        return f"    static constexpr const char* {tag}[{max(1, len(items))}] = {{{body}}};\n"

    head = (
        f"// AUTO-GENERATED by tpl_b200.codegen from the problem definition '{name}'.\n"
        f"// definition sha1: {definition_hash}\n"
        f"// Do not edit; regenerate with `python -m tpl_b200.build --regen`.\n"
        f"#pragma once\n\n"
        f"struct Model {{\n"
        f"    static constexpr int X = {d.X};\n"
        f"    static constexpr int U = {d.U};\n"
        f"    static constexpr int C = {d.C};\n"
        f"    static constexpr int NUM_SCALARS = {len(d.scalar_params)};\n"
        f"    static constexpr int NUM_ARRAYS = {len(d.array_params)};\n"
        f'    static constexpr const char* NAME = "{name}";\n'
        f'    static constexpr const char* DEFINITION_SHA1 = "{definition_hash}";\n'
        + names("STATE_NAMES", d.state_names)
        + names("ACTION_NAMES", d.action_names)
        + names("SCALAR_NAMES", d.scalar_params)
        + names("ARRAY_NAMES", d.array_params)
        + names("PARAM_ORDER", d.param_order)
        + f"    static constexpr int NUM_PARAMS = {len(d.param_order)};\n"
        + "    // structure of the derivative blocks: 0 = zero, 1 = one, 2 = general\n"
        + _c_array("signed char", "FX_KIND", entry_kinds(r["stateJacobian"]))
        + _c_array("signed char", "FU_KIND", entry_kinds(r["actionJacobian"]))
        + _c_array("signed char", "LXX_KIND", entry_kinds(r["stateStateHessian"]))
        + _c_array("signed char", "LUU_KIND", entry_kinds(r["actionActionHessian"]))
        + _c_array("signed char", "LUX_KIND", entry_kinds(r["actionStateHessian"]))
        + _c_array("signed char", "VXX_END_KIND", entry_kinds(r["endHessian"]))
        + "\n"
    )
    return head + "\n".join(parts) + "};\n"


# ---------------------------------------------------------------------------------
# plain-C routines for the CPU oracle (one function per reference routine)
# ---------------------------------------------------------------------------------

_C_SIG = {
    "dyn": "const double* x, const double* u, const double t, const double dt",
    "stage": ("const double* x, const double* u, const double* lg_mult, "
              "const double* lg_weight, const double t, const double dt"),
    "end": "const double* x, const double t, const double dt",
}
_C_KIND = {
    "ctDynamics": "dyn", "stateJacobian": "dyn", "actionJacobian": "dyn",
    "costs": "stage", "stateGradient": "stage", "actionGradient": "stage",
    "stateStateHessian": "stage", "actionActionHessian": "stage", "actionStateHessian": "stage",
    "endCosts": "end", "endGradient": "end", "endHessian": "end",
    "constraints": "stage",
}


def emit_c_model(d: Derivation, name: str, definition_hash: str) -> str:
    """Per-routine C99 functions with the reference's routine names
    (genopt.py:105-110, 141-148, 168-172, 177-179), each with its own CSE pass
    like the reference's ``codegen(cse=True)``."""
    pr = ModelPrinter(d, PLAIN_C)
    out = [
        f"/* AUTO-GENERATED by oracle/gen_models.py for '{name}' (definition sha1 {definition_hash}). */",
        f"#define TPLO_X {d.X}",
        f"#define TPLO_U {d.U}",
        f"#define TPLO_C {d.C}",
        f"#define TPLO_NUM_SCALARS {len(d.scalar_params)}",
        f"#define TPLO_NUM_ARRAYS {len(d.array_params)}",
        f'#define TPLO_MODEL_NAME "{name}"',
        "",
    ]
    for rname, m in d.routines.items():
        body = print_body(pr, [("out", m)], "    ")
        out.append(f"static void {rname}(const tplo_params* P, {_C_SIG[_C_KIND[rname]]}, double* out) {{")
        out.append("    (void)P; (void)x; (void)t; (void)dt;")
        out.append(body)
        out.append("}\n")
    return "\n".join(out)
