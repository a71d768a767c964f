"""Problem definition -> batched B200 iLQR solver.

Keeps the reference's entry points (/root/reference/library/tpl/optim/genopt.py):
``Config`` (:428-449), ``build(config)`` (:650-655) and ``build_parallel(configs)``
(:658-665) — but instead of a CPython extension wrapping a scalar C solver it
emits CUDA model routines (``codegen``), compiles them together with the
hand-written sm_100a solver kernels (``csrc/``) into one C-ABI shared library
per problem (``include/tplb200.h``) and returns a factory for
:class:`tpl_b200.batched.BatchedOptim`.

Libraries are cached by the sha1 of the problem definition and of the solver
sources, like the reference caches by ``code_hash`` (genopt.py:493-505).
"""

import hashlib
import os
from concurrent.futures import ThreadPoolExecutor

import sympy as sp

from . import symext as spx


class Config:
    """Problem definition; same fields as the reference's ``genopt.Config``
    (genopt.py:428-449).  ``params`` is a list of symbols or a dict
    ``symbol -> default value``."""

    def __init__(self, states, actions, params, dynamics, costs,
                 end_costs=0.0, constraints=(), use_cache=True,
                 output_dir=None):
        self.states = list(states)
        self.actions = list(actions)
        self.params = params
        self.dynamics = dynamics
        self.costs = costs
        self.end_costs = end_costs
        self.constraints = list(constraints)
        self.use_cache = use_cache
        self.output_dir = output_dir

    # -- helpers -----------------------------------------------------------
    @property
    def param_symbols(self):
        return list(self.params.keys()) if isinstance(self.params, dict) else list(self.params)

    def definition_hash(self):
        """sha1 over the printed definition (genopt.py:493-505)."""
        text = "|".join(str(v) for v in (
            self.states, self.actions, self.param_symbols,
            sp.Matrix([self.dynamics]) if not isinstance(self.dynamics, sp.MatrixBase) else self.dynamics,
            self.costs, self.end_costs, self.constraints))
        return hashlib.sha1(text.encode("utf8")).hexdigest()


def _lib_dir(config):
    from .build import default_lib_dir
    return os.path.expanduser(config.output_dir) if config.output_dir else default_lib_dir()


def build_module(config, name=None, force=False):
    """Generate + compile the solver library for ``config``; returns its path.
    Counterpart of genopt.py:464-619."""
    from . import build as _build
    return _build.build_model_library(config, name=name, lib_dir=_lib_dir(config),
                                      force=force or not config.use_cache)


def get_opt_builder(lib_path, config):
    """Factory applying the default parameter values (genopt.py:622-647)."""
    from .batched import BatchedOptim

    defaults = {}
    if isinstance(config.params, dict):
        defaults = {s.name: v for s, v in config.params.items() if v is not None}

    def init_opt(batch=1, scenes=None, horizon_max=None, **kw):
        o = BatchedOptim(lib_path, batch=batch, scenes=scenes, horizon_max=horizon_max, **kw)
        for pname, val in defaults.items():
            setattr(o.params, pname, val)
        return o

    init_opt.lib_path = lib_path
    return init_opt


def build(config, name=None, force=False):
    """``Opt = genopt.build(config); opt = Opt(batch=B)`` (genopt.py:650-655)."""
    return get_opt_builder(build_module(config, name=name, force=force), config)


def build_parallel(configs, names=None, force=False):
    """Build several problems concurrently (genopt.py:658-665).  sympy work runs
    in this process; the nvcc invocations, which dominate, run in parallel."""
    names = names or [None] * len(configs)
    from . import build as _build
    prepared = [_build.prepare_model_sources(c, name=n, lib_dir=_lib_dir(c)) for c, n in zip(configs, names)]
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(prepared)))) as pool:
        paths = list(pool.map(lambda p: _build.compile_prepared(p, force=force), prepared))
    spx.clear_cache()
    return [get_opt_builder(p, c) for p, c in zip(paths, configs)]
