"""TEST INFRASTRUCTURE — Python face of the CPU oracle (oracle/ilqr_oracle.c).

``OracleOptim(name)`` behaves like one reference ``Optim`` object
(optim.c:1654-1892): same attribute names, numpy arrays of the current horizon,
``update() / shift() / dynamics() / ct_dynamics()``.  Only tests/, smoke() and
bench.py's cpu_baseline leg may import this module.
"""

import copy
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_ARRAYS, MAX_SCALARS = 16, 64
H_MAX = 300                      # optim.c:49


class _Params(C.Structure):
    _fields_ = [("scalar", C.c_double * MAX_SCALARS),
                ("array", C.c_void_p * MAX_ARRAYS),
                ("length", C.c_int64 * MAX_ARRAYS),
                ("cols", C.c_int64 * MAX_ARRAYS)]


_PTRS = ("x u next_x next_u prev_x prev_k fx fu lx lu lxx luu lux g k K "
         "lam barrier_weight lg_mult_limit u_min u_max").split()


class _Problem(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in
                 "T opt_start max_iterations max_lg_iterations integrator use_quadratic_terms".split()]
                + [("dt", C.c_double), ("min_rel_cost_change", C.c_double),
                   ("traj_costs", C.c_double), ("alpha", C.c_double), ("mu", C.c_double)]
                + [(n, C.c_int32) for n in
                   "iterations lg_iterations mu_step trajectory_changed improved termination_condition".split()]
                + [(n, C.c_void_p) for n in _PTRS]
                + [("params", _Params)])


def build_libs():
    """Compile every oracle library (idempotent; ``make`` decides)."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def _model_meta(name, cfg=None):
    """scalar / array parameter names of a problem definition (the package's definitions
    are the single source of truth for the names)."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from tpl_b200 import optimizers, symext
    cfg = cfg or optimizers.CONFIGS[name]()
    ps = cfg.param_symbols
    scal = [p.name for p in ps if not isinstance(p, symext.ArraySymbol)]
    arrs = [p.name for p in ps if isinstance(p, symext.ArraySymbol)]
    return scal, arrs


_META_CACHE = {}


def build_custom(cfg, out_dir):
    """Oracle library for a user-defined problem: generate its C routines and compile
    ilqr_oracle.c against them.  Returns (name, library path)."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from tpl_b200 import codegen, derive
    name = "custom_" + cfg.definition_hash()[:12]
    os.makedirs(out_dir, exist_ok=True)
    header = os.path.join(out_dir, name + ".h")
    with open(header, "w") as fd:
        fd.write(codegen.emit_c_model(derive.derive(cfg), name, cfg.definition_hash()))
    lib = os.path.join(out_dir, f"liboracle_{name}.so")
    subprocess.run(["gcc", "-O2", "-fno-fast-math", "-ffp-contract=off", "-fPIC", "-shared", "-w",
                    f'-DTPLO_MODEL_HEADER="{header}"', os.path.join(HERE, "ilqr_oracle.c"),
                    "-o", lib, "-lm"], check=True)
    _META_CACHE[name] = _model_meta(name, cfg)
    return name, lib


class OracleParams:
    """Named parameters; arrays are copied in on assignment (optim.c:1361-1391)."""

    def __init__(self, owner, scalars, arrays):
        object.__setattr__(self, "_o", owner)
        object.__setattr__(self, "_scalars", scalars)
        object.__setattr__(self, "_arrays", arrays)
        object.__setattr__(self, "_store", {a: np.zeros(0) for a in arrays})

    def names(self):
        return list(self._o._param_order)

    def __getattr__(self, n):
        if n in self._scalars:
            return float(self._o._c.params.scalar[self._scalars.index(n)])
        if n in self._arrays:
            return self._store[n]
        raise AttributeError(n)

    def __setattr__(self, n, v):
        if n in self._scalars:
            self._o._c.params.scalar[self._scalars.index(n)] = float(v)
        elif n in self._arrays:
            a = np.ascontiguousarray(np.asarray(v, dtype=np.float64)).copy()
            self._store[n] = a
            i = self._arrays.index(n)
            self._o._c.params.array[i] = a.ctypes.data
            self._o._c.params.length[i] = a.shape[0] if a.ndim else 0
            self._o._c.params.cols[i] = a.shape[1] if a.ndim == 2 else 0
        else:
            raise AttributeError(n)


class OracleOptim:
    EULER, HEUN, RK4 = 0, 1, 2

    def __init__(self, name, lib_path=None):
        path = lib_path or os.path.join(HERE, "lib", f"liboracle_{name}.so")
        if not os.path.exists(path):
            build_libs()
        self._name = name
        self._lib_path = lib_path
        self._lib = C.CDLL(path)
        dims = (C.c_int * 5)()
        size = self._lib.tplo_dims(dims)
        assert size == C.sizeof(_Problem), (size, C.sizeof(_Problem))
        self.X, self.U, self.C = dims[0], dims[1], dims[2]
        if name not in _META_CACHE:
            _META_CACHE[name] = _model_meta(name)
        scal, arrs = _META_CACHE[name]
        assert len(scal) == dims[3] and len(arrs) == dims[4]
        self._param_order = scal + arrs
        self._c = _Problem()
        X, U, Cn = self.X, self.U, self.C
        shapes = dict(x=(H_MAX, X), u=(H_MAX, U), next_x=(H_MAX, X), next_u=(H_MAX, U),
                      prev_x=(H_MAX, X), prev_k=(H_MAX, U),
                      fx=(H_MAX, X, X), fu=(H_MAX, X, U), lx=(H_MAX, X), lu=(H_MAX, U),
                      lxx=(H_MAX, X, X), luu=(H_MAX, U, U), lux=(H_MAX, U, X),
                      g=(H_MAX, U), k=(H_MAX, U), K=(H_MAX, U, X),
                      lam=(H_MAX, Cn), barrier_weight=(Cn,), lg_mult_limit=(Cn,),
                      u_min=(H_MAX, U), u_max=(H_MAX, U))
        self._buf = {n: np.zeros(s) for n, s in shapes.items()}
        for n in _PTRS:
            setattr(self._c, n, self._buf[n].ctypes.data)
        # construction defaults (optim.c:1894-1921)
        c = self._c
        c.dt, c.T, c.min_rel_cost_change = 0.05, 20, 1e-6
        c.max_iterations, c.max_lg_iterations, c.use_quadratic_terms = 5, 1, 1
        self._buf["lg_mult_limit"][:] = np.inf
        self._buf["barrier_weight"][:] = 1.0
        self._buf["u_max"][:20] = np.inf
        self._buf["u_min"][:20] = -np.inf
        self.params = OracleParams(self, scal, arrs)

    # -- arrays, squeezed like optim.c:1314-1347 --------------------------------
    _LEN = dict(x=1, next_x=1, prev_x=1)

    def _view(self, n):
        b = self._buf[n]
        if n in ("barrier_weight", "lg_mult_limit"):
            return b
        T = self._c.T + self._LEN.get(n, 0)
        v = b[:T]
        if v.ndim > 1 and 1 in v.shape[1:]:
            v = v.reshape([T] + [d for d in v.shape[1:] if d != 1])
        return v

    _ALIASES = dict(lagrange_multiplier="lam")
    _SCALARS = {"traj_costs", "alpha", "mu", "iterations", "lg_iterations", "mu_step",
                "trajectory_changed", "improved", "termination_condition", "max_iterations",
                "max_lg_iterations", "min_rel_cost_change", "opt_start", "use_quadratic_terms", "dt"}

    def __getattr__(self, n):
        if n.startswith("_") or n in ("params", "X", "U", "C"):
            raise AttributeError(n)
        n = self._ALIASES.get(n, n)
        if n in self._buf:
            return self._view(n)
        if n in self._SCALARS:
            return getattr(self._c, n)
        if n == "step":
            return self._c.dt
        if n in ("horizon", "T"):
            return self._c.T
        if n == "integrator_type":
            return self._c.integrator
        raise AttributeError(n)

    def __setattr__(self, n, v):
        if n.startswith("_") or n in ("params", "X", "U", "C"):
            return object.__setattr__(self, n, v)
        n = self._ALIASES.get(n, n)
        if n in self._buf:
            self._view(n)[...] = v
        elif n in self._SCALARS:
            setattr(self._c, n, type(getattr(self._c, n))(v))
        elif n == "step":
            self._c.dt = float(v)
        elif n in ("horizon", "T"):
            self._c.T = min(H_MAX - 1, max(1, int(v)))          # optim.c:1726-1734
        elif n == "integrator_type":
            self._c.integrator = int(v)
        else:
            raise AttributeError(n)

    # -- methods ------------------------------------------------------------------
    def update(self):
        self._lib.tplo_update(C.byref(self._c))

    def linearize(self):
        self._lib.tplo_linearize(C.byref(self._c))

    def shift(self, amount):
        self._lib.tplo_shift(C.byref(self._c), int(amount))

    def _point(self, fn, x, u, t, dt):
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.ascontiguousarray(np.atleast_1d(u), dtype=np.float64)
        if x.shape != (self.X,):
            raise ValueError(f'Expected "x_arr" with shape ({self.X}), but found {x.shape}')
        if u.shape != (self.U,):
            raise ValueError(f'Expected "u_arr" with shape ({self.U}), but found {u.shape}')
        out = np.zeros(self.X)
        fn(C.byref(self._c), x.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p),
           C.c_int(int(t)), C.c_double(float(dt)), out.ctypes.data_as(C.c_void_p))
        return out

    def dynamics(self, x, u, t, dt):
        return self._point(self._lib.tplo_dynamics, x, u, t, dt)

    def ct_dynamics(self, x, u, t, dt):
        return self._point(self._lib.tplo_ct_dynamics, x, u, t, dt)

    def __deepcopy__(self, memo):
        o = OracleOptim(self._name, self._lib_path)
        for n, b in self._buf.items():
            o._buf[n][...] = b
        for f, _ in _Problem._fields_:
            if f not in _PTRS and f != "params":
                setattr(o._c, f, getattr(self._c, f))
        for i in range(MAX_SCALARS):
            o._c.params.scalar[i] = self._c.params.scalar[i]
        for a in self.params._arrays:
            setattr(o.params, a, copy.deepcopy(self.params._store[a]))
        return o
