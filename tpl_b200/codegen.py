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


CUDA = Dialect("cuda", "R(%s)", "P.%s(%s%s)", "P.scalar(%d)", "R")
PLAIN_C = Dialect("c", "%s", "tplo_%s(P, %s%s)", "P->scalar[%d]", "double")


class ModelPrinter(C99CodePrinter):
    """C99 printer with the reference's ``Pow`` shortcuts (genopt.py:188-215)
    made single-evaluation (``sq(a)`` instead of ``(a*a)``), typed literals, and
    parameter / interpolation access through the dialect."""

    # CUDA dialect: straight-line elementary functions of csrc/fast_math.cuh (m_* resolve to
    # them, or to libdevice when a library is built with -DTPLB_LIBDEVICE_MATH)
    CUDA_FUNCTIONS = {"sin": "m_sin", "cos": "m_cos", "tan": "m_tan"}

    def __init__(self, deriv: Derivation, dialect: Dialect):
        settings = {"allow_unknown_functions": True}
        if dialect.name == "cuda":
            settings["user_functions"] = dict(self.CUDA_FUNCTIONS)
        super().__init__(settings)
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
    def _params_only(self, e):
        return bool(e.free_symbols) and all(f.name in self.scalar_index for f in e.free_symbols)

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
            return "%s(%s)" % ("m_sqrt" if self.dialect.name == "cuda" else "sqrt", pb)
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
        if len(arrays) != (2 if fn == "lerp_wrap" else 1):
            raise NotImplementedError(f"{fn} with {len(arrays)} array arguments")
        idx = ", ".join(str(self.array_index[a.name]) for a in arrays)      # lerp_wrap: xs, arr (optim.c:450)
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


_RSQRT = sp.Function("m_rsqrt")
_INV = sp.Function("m_inv")


def _strength_reduce(exprs):
    """CUDA dialect only, applied to the already CSE'd expressions right before printing:

    * ``b**(-1/2)`` -> ``m_rsqrt(b)`` (one MUFU + Newton steps instead of sqrt + divide), and
      ``b**(-3/2)`` -> ``m_rsqrt(b)**3``;
    * ``b**(-n)``   -> ``m_inv(b)**n``: every division becomes a multiplication by a straight-line
      reciprocal, and all negative powers of one base share ONE reciprocal (the bicycle models
      use 1/b, 1/b^2, 1/b^3, 1/b^4 of the same two sums) — ``m_rsqrt(b)**(2n)`` where the
      reciprocal root of ``b`` is needed anyway.  Reciprocals of parameter-only expressions are
      loop invariant, so the compiler hoists them out of the stage loops entirely.
      (sympy prints negative powers inside a product as a division, so this is a tree rewrite,
      not a printer hook.)

    Every rewritten factor differs from the division it replaces by a few ulp at most."""
    half, three_half = sp.Rational(-1, 2), sp.Rational(-3, 2)
    pows = set()
    for e in exprs:
        pows |= e.atoms(sp.Pow)
    rs_bases = {p.base for p in pows if p.exp in (half, three_half)}

    def rewrite(p):
        b, x = p.base, p.exp
        if x == half:
            return _RSQRT(b)
        if x == three_half:
            return sp.Pow(_RSQRT(b), 3)
        if x.is_Integer:
            n = -int(x)
            return sp.Pow(_RSQRT(b), 2 * n) if b in rs_bases else sp.Pow(_INV(b), n)
        return _INV(sp.Pow(b, -x))

    return [e.replace(lambda p: isinstance(p, sp.Pow) and p.exp.is_number and p.exp.is_negative, rewrite)
            for e in exprs]


def _pair_sincos(exprs):
    """sin(a) and cos(a) of the same argument -> one ``sincos`` call.  Returns the
    rewritten expressions and the list of paired arguments (already rewritten)."""
    sins, coss = {}, {}
    for e in exprs:
        for f in e.atoms(sp.sin):
            sins[f.args[0]] = f
        for f in e.atoms(sp.cos):
            coss[f.args[0]] = f
    args = sorted((a for a in sins if a in coss), key=str)
    repl = {}
    for k, a in enumerate(args):
        repl[sins[a]] = sp.Symbol(f"sn{k}")
        repl[coss[a]] = sp.Symbol(f"cs{k}")
    return [e.xreplace(repl) for e in exprs], [a.xreplace(repl) for a in args]


def _expand_trig_of_known_angles(exprs):
    """CUDA dialect: ``sin / cos(A +- B)`` by the angle-addition formulas when the sine or cosine of A
    and of B are evaluated in the same routine anyway.  The path-relative bicycle model takes
    ``cos(ref_phi(s) - phi)`` next to ``sincos(ref_phi(s))`` and ``sincos(phi)``: evaluated literally
    the third call waits for the lookup AND runs a full argument reduction + polynomial inside the
    rollout's dependent chain; as ``cos A cos B + sin A sin B`` it is two FMAs on values that exist
    already.  Mathematically identical, absolute rounding error of the same size."""
    known = set()
    for e in exprs:
        for f in e.atoms(sp.sin, sp.cos):
            known.add(f.args[0])

    def signed(term):
        if isinstance(term, sp.Mul) and term.args[0] == -1:
            return -1, sp.Mul(*term.args[1:])
        return 1, term

    def rewrite(f):
        arg = f.args[0]
        if not isinstance(arg, sp.Add) or len(arg.args) != 2:
            return f
        (sa, a), (sb, b) = signed(arg.args[0]), signed(arg.args[1])
        if a not in known or b not in known or a.is_number or b.is_number:
            return f
        sin_a, sin_b = sa * sp.sin(a), sb * sp.sin(b)
        if isinstance(f, sp.cos):
            return sp.cos(a) * sp.cos(b) - sin_a * sin_b
        return sin_a * sp.cos(b) + sp.cos(a) * sin_b

    return [e.replace(lambda x: isinstance(x, (sp.sin, sp.cos)), rewrite) for e in exprs]


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
    pair_args = []
    if printer.dialect.name == "cuda":
        flat = _expand_trig_of_known_angles(flat)
        flat, pair_args = _pair_sincos(flat)
    n_out = len(flat)
    if flat:
        common, reduced = sp.cse(flat + pair_args, symbols=sp.numbered_symbols("v"))
    else:
        common, reduced = [], []

    if printer.dialect.name == "cuda":
        rewritten = _strength_reduce([e for _, e in common] + list(reduced))
        common = [(sym, e) for (sym, _), e in zip(common, rewritten)]
        reduced = rewritten[len(common):]

    # definitions in dependency order: CSE temporaries and sincos pairs
    defs = {}
    for sym, expr in common:
        ctype = "bool" if _is_boolean(expr) else real
        defs[sym.name] = ([sym.name], {f.name for f in expr.free_symbols},
                          f"{indent}const {ctype} {sym} = {_one_line(printer.doprint(expr))};")
    for k in range(len(pair_args)):
        arg = reduced[n_out + k]
        node = ([f"sn{k}", f"cs{k}"], {f.name for f in arg.free_symbols},
                f"{indent}{real} sn{k}, cs{k}; m_sincos({_one_line(printer.doprint(arg))}, &sn{k}, &cs{k});")
        defs[f"sn{k}"] = defs[f"cs{k}"] = node
    lines, done = [], set()

    def need(name):
        if name in done or name not in defs:
            return
        produced, deps, text = defs[name]
        done.update(produced)
        for dep in sorted(deps):
            need(dep)
        lines.append(text)

    for sym, _ in common:
        need(sym.name)
    for e in reduced[:n_out]:
        for f in sorted(e.free_symbols, key=str):
            need(f.name)
    for name, start, n in spans:
        for k in range(n):
            lines.append(f"{indent}{name}[{k}] = {_one_line(printer.doprint(reduced[start + k]))};")
    loads = [f"{indent}const {real} p_{n} = {printer.dialect.scalar_load % printer.scalar_index[n]};"
             for n in printer.d.scalar_params if n in printer.used_scalars]
    return "\n".join(loads + lines)


# ---------------------------------------------------------------------------------
# CUDA model header
# ---------------------------------------------------------------------------------

_STAGE_ARGS = ("const R* x, const R* u, const R* lg_mult, const R* lg_weight, const R* sc, "
               "const R t, const R dt")
_DYN_ARGS = "const R* x, const R* u, const R* sc, const R t, const R dt"
_END_ARGS = "const R* x, const R* sc, const R t, const R dt"

# order of the blocks inside one stage's derivative record
DERIV_BLOCKS = ("stateJacobian", "actionJacobian", "stateGradient", "actionGradient",
                "stateStateHessian", "actionActionHessian", "actionStateHessian")


def _cuda_fn(name, args, outs, body):
    outs_decl = "".join(f", R* {o}" for o in outs)
    return (f"    template <typename R, typename PV>\n"
            f"    __device__ __forceinline__ static void {name}(const PV& P, {args}{outs_decl}) {{\n"
            f"{body}\n"
            f"    }}\n")


def _lookup_fn(name, values, ctype="int"):
    """``constexpr`` table lookup usable from host and device code."""
    vals = ", ".join(str(v) for v in values) if values else "0"
    n = max(1, len(values))
    return (f"    __host__ __device__ static constexpr {ctype} {name}(int e) {{\n"
            f"        constexpr {ctype} table[{n}] = {{{vals}}};\n"
            f"        return table[e];\n"
            f"    }}\n")


def expand_trig_of_atan(expr):
    """sin / cos of ``a + atan(z)`` by the angle-addition formulas:

        cos(a + atan z) = (cos a - z sin a) / sqrt(1 + z^2)
        sin(a + atan z) = (sin a + z cos a) / sqrt(1 + z^2)

    The kinematic-bicycle models take the sine and cosine of heading + slip angle with
    slip = atan(c tan(delta)).  Evaluated literally this is a chain of three dependent
    transcendental calls (tan -> atan -> sincos) in the innermost recursion of the
    rollout; rewritten, the atan disappears and sincos(heading) no longer waits for
    tan(delta).  Mathematically identical, rounding differs by a few ulp."""
    def rewrite(f):
        arg = f.args[0]
        if not isinstance(arg, sp.Add):
            return f
        atans = [t for t in arg.args if isinstance(t, sp.atan)]
        neg = [t for t in arg.args if isinstance(t, sp.Mul) and t.args[0] == -1 and len(t.args) == 2
               and isinstance(t.args[1], sp.atan)]
        if len(atans) + len(neg) != 1:
            return f
        if atans:
            z, rest = atans[0].args[0], arg - atans[0]
        else:
            z, rest = -neg[0].args[1].args[0], arg - neg[0]
        r = 1 / sp.sqrt(1 + z**2)
        if isinstance(f, sp.cos):
            return (sp.cos(rest) - z * sp.sin(rest)) * r
        return (sp.sin(rest) + z * sp.cos(rest)) * r

    return expr.replace(lambda e: isinstance(e, (sp.sin, sp.cos)), rewrite)


def hoist_stage_constants(d: Derivation):
    """Interpolation lookups whose arguments depend only on the stage index, the
    step and the parameters are constant per (scene, stage) for a whole solve
    (SURVEY.md appendix C).  Returns the routines with every such call replaced
    by ``sc[i]`` and the list of hoisted calls; the values are produced by the
    same printed expression, so nothing changes numerically."""
    allowed = {"t", "dt"} | set(d.scalar_params) | set(d.array_params)
    calls = set()
    for m in d.routines.values():
        for fn in spx.OPAQUE_FUNCTIONS:
            for call in m.atoms(fn):
                if all(s.name in allowed for s in call.free_symbols):
                    calls.add(call)
    calls = sorted(calls, key=str)
    repl = {c: sp.Symbol(f"sc[{i}]") for i, c in enumerate(calls)}
    routines = {n: (expand_trig_of_atan(m.xreplace(repl)) if m.shape[0] * m.shape[1] else m)
                for n, m in d.routines.items()}
    return routines, calls


def derivative_layout(r):
    """Compact storage of one stage's derivative record.

    Dense order: fx | fu | lx | lu | lxx | luu | lux, row-major each.  Entries that
    are identically 0 or 1 are not stored (slot -1 / -2); the two halves of the
    symmetric Hessian blocks share one slot.  Returns (slot per dense entry,
    owner flag per dense entry, number of slots, dense offsets)."""
    slots, owner, offsets = [], [], {}
    n = 0
    for name in DERIV_BLOCKS:
        m = r[name]
        offsets[name] = len(slots)
        rows, cols = m.shape
        local = {}
        for i in range(rows):
            for j in range(cols):
                e = m[i, j]
                if e == 0:
                    slots.append(-1); owner.append(0)
                elif e == 1:
                    slots.append(-2); owner.append(0)
                elif name in ("stateStateHessian", "actionActionHessian") and j < i and m[j, i] == e:
                    slots.append(local[(j, i)]); owner.append(0)
                else:
                    local[(i, j)] = n
                    slots.append(n); owner.append(1)
                    n += 1
    return slots, owner, n, offsets


def array_shapes(d: Derivation):
    """(ndim per array parameter, [(xs, arr) index pairs of lerp_wrap]) as the lookups use them:
    blerp reads a 2-D array (optim.c:483-486), everything else 1-D; lerp_wrap's two arrays must
    have the same length (optim.c:450-455).  The host checks bound arrays against this, where
    the reference's `check()` raises after the solve."""
    index = {n: i for i, n in enumerate(d.array_params)}
    ndim = {}
    pairs = set()
    for m in d.routines.values():
        for call in m.atoms(*spx.OPAQUE_FUNCTIONS):
            arrays = [a for a in call.args if isinstance(a, spx.ArraySymbol)]
            for a in arrays:
                want = 2 if isinstance(call, spx.blerp) else 1
                if ndim.setdefault(a.name, want) != want:
                    raise ValueError(f"array parameter {a.name} is used both as a 1-D and as a 2-D array")
            if isinstance(call, spx.lerp_wrap) and len(arrays) == 2:
                pairs.add((index[arrays[0].name], index[arrays[1].name]))
    return [ndim.get(n, 1) for n in d.array_params], sorted(pairs)


def _count_lookups(m):
    names = {"lerp", "lerp_angle", "lerp_wrap", "box_interp", "blerp", "get_array_value"}
    return len({f for e in _flatten(m) for f in sp.sympify(e).atoms(sp.Function) if type(f).__name__ in names})


def emit_cuda_model(d: Derivation, name: str, definition_hash: str) -> str:
    """The ``struct Model`` consumed by csrc/solver.cuh."""
    r, hoisted = hoist_stage_constants(d)
    pr = ModelPrinter(d, CUDA)
    ind = "        "
    parts = []
    parts.append(_cuda_fn("stage_constants", "const R t, const R dt", ["sc"],
                          print_body(pr, [("sc", sp.Matrix(hoisted) if hoisted else sp.zeros(0, 1))], ind)))
    parts.append(_cuda_fn("ct_dynamics", _DYN_ARGS, ["f"],
                          print_body(pr, [("f", r["ctDynamics"])], ind)))
    parts.append(_cuda_fn("dynamics_jacobians", _DYN_ARGS, ["fx", "fu"],
                          print_body(pr, [("fx", r["stateJacobian"]), ("fu", r["actionJacobian"])], ind)))
    parts.append(_cuda_fn("stage_cost", _STAGE_ARGS, ["c"],
                          print_body(pr, [("c", r["costs"])], ind)))
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

    def ints(tag, items):
        body = ", ".join(str(v) for v in items) if items else "0"
        return f"    static constexpr int {tag}[{max(1, len(items))}] = {{{body}}};\n"

    ndim, wrap_pairs = array_shapes(d)
    slots, owner, nslots, offsets = derivative_layout(r)
    vxx_kind = entry_kinds(r["endHessian"])
    hoisted_doc = "".join(f"    //   sc[{i}] = {c}\n" for i, c in enumerate(hoisted))

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
        f"    static constexpr int NUM_PARAMS = {len(d.param_order)};\n"
        f'    static constexpr const char* NAME = "{name}";\n'
        f'    static constexpr const char* DEFINITION_SHA1 = "{definition_hash}";\n'
        + names("STATE_NAMES", d.state_names)
        + names("ACTION_NAMES", d.action_names)
        + names("SCALAR_NAMES", d.scalar_params)
        + names("ARRAY_NAMES", d.array_params)
        + names("PARAM_ORDER", d.param_order)
        + "    // dimensions of every array parameter (2: read by blerp) and the (xs, arr) pairs of lerp_wrap\n"
        + ints("ARRAY_NDIM", ndim)
        + f"    static constexpr int NUM_WRAP_PAIRS = {len(wrap_pairs)};\n"
        + ints("WRAP_PAIRS", [i for p in wrap_pairs for i in p])
        + "\n    // per-(scene, stage) constants: lookups that depend on the stage index only\n"
        + hoisted_doc
        + f"    static constexpr int NUM_STAGE_CONSTS = {len(hoisted)};\n"
        + "    // array lookups left in the dynamics (their argument depends on the state): such a rollout branches\n"
        + f"    static constexpr int DYNAMICS_LOOKUPS = {_count_lookups(r['ctDynamics'])};\n"
        + "\n    // one stage's derivative record: dense order fx|fu|lx|lu|lxx|luu|lux (row-major);\n"
        + "    // deriv_slot(e) >= 0: index in the compact record, -1: identically 0, -2: identically 1\n"
        + f"    static constexpr int DERIV_DENSE = {len(slots)};\n"
        + f"    static constexpr int DERIV_COMPACT = {nslots};\n"
        + _lookup_fn("deriv_slot", slots)
        + _lookup_fn("deriv_owner", owner)
        + "    // end-cost Hessian: 0 = zero, 1 = one, 2 = general\n"
        + _lookup_fn("vxx_end_kind", vxx_kind)
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
