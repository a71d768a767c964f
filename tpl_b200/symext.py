"""Symbolic vocabulary for optimal-control problem definitions.

Mirrors the interface of the reference's ``tpl.optim.symext``
(/root/reference/library/tpl/optim/symext.py:13-157): opaque array parameters,
opaque interpolation functions the code generator knows how to print, and the
``fixed``/``unfixed`` pair that hides a symbol from differentiation.

The interpolation functions are plain undefined sympy functions: sympy
differentiates them into ``Derivative(...)`` nodes, which the derivation
pipeline (``genopt.derive``) replaces by central finite differences exactly as
the reference does (genopt.py:65-70).
"""

import copy

import sympy as sp
from sympy.core import cache as _sp_cache

_FIXED_PREFIX = "fixed_"


class ArraySymbol(sp.Symbol):
    """Opaque 1-D (or 2-D for ``blerp``) array of doubles of run-time length.

    Only ever appears as the last argument of an interpolation function;
    reference: symext.py:13-22.
    """

    def __new__(cls, name, **assumptions):
        return super().__new__(cls, name, **assumptions)


class _Opaque(sp.Function):
    """Base of the functions that have a meaning only to the code printers."""


class get_array_value(_Opaque):
    """``arr[(size_t) i]`` — reference symext.py:25-35, optim.c:330."""


class box_interp(_Opaque):
    """Nearest-lower-sample lookup ``(dx, x, arr)`` — optim.c:357-372."""
    nargs = (3,)


class lerp(_Opaque):
    """Clamped linear interpolation ``(x0, dx, x, arr)`` — optim.c:374-390."""
    nargs = (4,)


class lerp_angle(_Opaque):
    """Shortest-arc angular interpolation ``(x0, dx, x, arr)`` — optim.c:392-408."""
    nargs = (4,)


class blerp(_Opaque):
    """Bilinear interpolation ``(x0, y0, dx, dy, x, y, arr)`` — optim.c:457-486."""
    nargs = (7,)


class lerp_wrap(_Opaque):
    """Periodic linear interpolation ``(len, dx, x, xs, arr)`` — optim.c:410-455."""
    nargs = (5,)


OPAQUE_FUNCTIONS = (get_array_value, box_interp, lerp, lerp_angle, blerp, lerp_wrap)


def clear_cache():
    """sympy's cache keeps symbols alive by name; reference symext.py:102-108."""
    _sp_cache.clear_cache()


def clone(expr):
    """Deep copy of symbols that survives sympy's cache (symext.py:111-122)."""
    clear_cache()
    return copy.deepcopy(expr)


def fixed(expr):
    """Rename every free symbol ``s`` to ``fixed_s`` so that differentiation
    with respect to ``s`` does not see it (symext.py:125-139)."""
    renames = {s: sp.Symbol(_FIXED_PREFIX + s.name)
               for s in expr.free_symbols
               if not s.name.startswith("fixed")}
    for old, new in renames.items():
        expr = expr.subs(old, new)
    return expr


def unfixed(expr):
    """Inverse of :func:`fixed` (symext.py:142-157): drop the first
    ``_``-separated token of every symbol whose name starts with ``fixed``."""
    for s in list(expr.free_symbols):
        if s.name.startswith("fixed"):
            expr = expr.subs(s, sp.Symbol("_".join(s.name.split("_")[1:])))
    return expr
