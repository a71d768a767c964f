"""`BatchedOptim` — the reference's ``Optim`` object for a batch of problems.

Host-side mirror of the generated CPython type ``genopt<sha1>.Optim``
(/root/reference/library/tpl/optim/templates/optim.c:1485-1921, "optim.c"): the
same attribute and method names, with a leading batch dimension.  All state
lives in torch CUDA tensors that are nothing but device buffers; every
operation is one call into the C ABI of ``include/tplb200.h``.

    opt = optimizers.trajectory_tracking_mpc_time(batch=4096, horizon_max=100)
    opt.horizon = 100; opt.step = 0.05; opt.integrator_type = opt.HEUN
    opt.params.ref_x = ref_x            # (S, L) or (L,) host/device array
    opt.x[:, 0] = x0                    # zero-copy device view, (B, T+1, X)
    opt.update()                        # optim.c:1091-1160 for all B problems
    opt.traj_costs, opt.iterations, opt.termination_condition    # (B,) tensors
    opt[i].x                            # per-problem view with the reference's shapes

Array attributes are *views* of the solver's structure-of-arrays buffers
(problem index fastest in memory), squeezed like the reference getters
(optim.c:1314-1325): ``u`` has shape ``(B, T)`` when U == 1.  Assignment
broadcasts and copies in (optim.c:1327-1347).

There is no CPU fallback: compute methods raise unless the buffers are on a CUDA
device and the sm_100a library loaded.
"""

import copy as _copy
import ctypes as C

import numpy as np
import torch

from . import _cabi

EULER, HEUN, RK4 = 0, 1, 2


def _as_tensor(value, device, dtype=torch.float64):
    if isinstance(value, torch.Tensor):
        return value.to(device=device, dtype=dtype, non_blocking=True)
    return torch.as_tensor(np.asarray(value, dtype=np.float64 if dtype == torch.float64 else None),
                           dtype=dtype).to(device)


def _source(value, dtype=torch.float64):
    """``value`` as a tensor of ``dtype`` on whatever device it lives on; the copy into the
    solver's buffer (``dst.copy_(src, non_blocking=True)``) then moves it in one step —
    asynchronously when ``value`` is a pinned host tensor."""
    if isinstance(value, torch.Tensor):
        return value if value.dtype == dtype else value.to(dtype)
    return torch.as_tensor(np.asarray(value, dtype=np.float64 if dtype == torch.float64 else None), dtype=dtype)


def _fit(value, view):
    """Squeeze surplus unit dimensions of ``value`` (the reference setter squeezes
    before copying, optim.c:1334-1336), then broadcast to ``view``."""
    while value.ndim > view.ndim and 1 in value.shape:
        dims = [i for i, d in enumerate(value.shape) if d == 1]
        value = value.squeeze(dims[-1])
    return value.expand_as(view)


class BatchedParams:
    """``opt.params`` (optim.c:1405-1476, genopt.py:321-417): named scalars and
    arrays.  A scalar is one value per scene (assign a float to broadcast); an
    array is ``(S, L)`` (assign ``(L,)`` to share it between all scenes); an array the problem
    reads with ``blerp`` is ``(S, rows, cols)`` / ``(rows, cols)``."""

    def __init__(self, owner):
        object.__setattr__(self, "_o", owner)

    @property
    def slots(self):
        return list(self._o._info["param_order"])

    @property
    def scalar_names(self):
        o = self._o
        return sorted(o._scalar_index, key=o._scalar_index.get)

    @property
    def array_names(self):
        o = self._o
        return sorted(o._array_index, key=o._array_index.get)

    def set_scalars(self, values):
        """All scalar parameters in one copy: ``values`` is (len(scalar_names), S) or
        (len(scalar_names),), rows in ``scalar_names`` order (batched extension: one
        host->device transfer instead of one per name)."""
        o = self._o
        t = _source(values)
        if t.ndim == 1:
            t = t.unsqueeze(1)
        if t.ndim != 2 or t.shape[0] != o._scalars.shape[0] or t.shape[1] not in (1, o.scenes):
            raise ValueError(f"Expected scalars with shape ({o._scalars.shape[0]}, {o.scenes}), "
                             f"but found {tuple(t.shape)}")
        o._scalars.copy_(t.expand_as(o._scalars), non_blocking=True)

    def __dir__(self):
        return self.slots + ["slots"]

    def __getattr__(self, n):
        o = self._o
        if n == "__slots__":
            return self.slots
        if n in o._scalar_index:
            return o._scalars[o._scalar_index[n]]
        if n in o._array_index:
            return o._arrays[o._array_index[n]]
        raise AttributeError(n)

    def __setattr__(self, n, v):
        o = self._o
        if n in o._scalar_index:
            o._scalars[o._scalar_index[n]].copy_(_source(v).expand(o.scenes), non_blocking=True)
        elif n in o._array_index:
            t = _source(v)
            i = o._array_index[n]
            nd = o._array_ndim[i]                      # 2: a map read by blerp (rows, cols), else 1
            if t.ndim == 0:
                raise ValueError(f"parameter {n} is an array")
            if t.ndim == nd:
                t = t.unsqueeze(0).expand(o.scenes, *t.shape)
            if t.ndim != nd + 1 or t.shape[0] != o.scenes:
                dims = "L" if nd == 1 else "rows, cols"
                raise ValueError(f'Expected "{n}" with shape ({o.scenes}, {dims}) or ({dims}), but found {tuple(t.shape)}')
            if o._arrays[i].shape == t.shape:          # same shape: refill the buffer in place
                o._arrays[i].copy_(t, non_blocking=True)
            else:
                o._arrays[i] = t.to(o.device).contiguous().clone()
                o._dirty = True
        else:
            raise AttributeError(f"no parameter named {n!r}")

    def __getstate__(self):
        o = self._o
        d = {n: o._scalars[i].cpu().numpy() for n, i in o._scalar_index.items()}
        d.update({n: o._arrays[i].cpu().numpy() for n, i in o._array_index.items()})
        return d


class HostMirror:
    """Pinned host staging buffers of one ``BatchedOptim`` — the reference keeps its arrays in
    host memory (numpy views of the ``Optim`` struct, optim.c:1302-1347); a batched device solve
    needs one host->device and one device->host transfer per batch instead.

    The buffers have the solver's own layout (problem index fastest), so ``opt.upload`` /
    ``opt.download`` are plain DMA transfers without a repacking kernel; the attributes are
    zero-copy *views* of them in the reference's shapes.  Inputs: ``x0`` (B, X), ``u0`` (B, T, U)
    squeezed like the device views, ``scalars`` (num_scalars, S) in ``params.scalar_names`` order,
    ``arrays[name]`` (S, L).  Results: ``x`` (B, T+1, X), ``u`` (B, T, U), ``traj_costs`` /
    ``iterations`` / ``termination_condition`` (B,).
    """

    def __init__(self, opt):
        B, T, X, U = opt.batch, opt.horizon, opt.X, opt.U
        pin = dict(pin_memory=True)
        self.horizon = T
        # inputs: initial state and control warm start
        self._x0 = torch.zeros((X, B), dtype=torch.float64, **pin)
        self._u0 = torch.zeros((T, U, B), dtype=torch.float64, **pin)
        self.x0 = self._x0.t()
        self.u0 = opt._traj_view(self._u0, T, (U,))
        # results
        self._x = torch.zeros((T + 1, X, B), dtype=torch.float64, **pin)
        self._u = torch.zeros((T, U, B), dtype=torch.float64, **pin)
        self.x = opt._traj_view(self._x, T + 1, (X,))
        self.u = opt._traj_view(self._u, T, (U,))
        self.traj_costs = torch.zeros(B, dtype=torch.float64, **pin)
        self._flags = torch.zeros((2, B), dtype=torch.int32, **pin)
        self.iterations, self.termination_condition = self._flags[0], self._flags[1]
        self.scalars = torch.zeros(tuple(opt._scalars.shape), dtype=torch.float64, **pin)
        self.arrays = {n: torch.zeros(tuple(opt._arrays[i].shape), dtype=torch.float64, **pin)
                       for n, i in opt._array_index.items()}

    def upload_bytes(self):
        return (self._x0.numel() + self._u0.numel() + self.scalars.numel()
                + sum(a.numel() for a in self.arrays.values())) * 8

    def download_bytes(self):
        return (self._x.numel() + self._u.numel() + self.traj_costs.numel()) * 8 + self._flags.numel() * 4


class ProblemView:
    """``opt[i]``: one problem with exactly the reference's shapes."""

    def __init__(self, owner, index):
        object.__setattr__(self, "_o", owner)
        object.__setattr__(self, "_i", index)

    def __getattr__(self, n):
        o = self._o
        if n in o._FIELDS or n in o._STATUS:
            return getattr(o, n)[self._i]
        return getattr(o, n)

    def __setattr__(self, n, v):
        o = self._o
        if n in o._FIELDS or n in o._STATUS:
            getattr(o, n)[self._i] = _as_tensor(v, o.device, getattr(o, n).dtype)
        else:
            raise AttributeError(f"{n} is shared by the whole batch; set it on the BatchedOptim")


class BatchedOptim:
    EULER, HEUN, RK4 = EULER, HEUN, RK4

    # name -> (buffer, rows-extra, component shape fn)
    _FIELDS = ("x", "u", "prev_x", "prev_u", "prev_k", "k", "K", "g", "lagrange_multiplier",
               "barrier_weight", "lg_mult_limit", "u_min", "u_max",
               "fx", "fu", "lx", "lu", "lxx", "luu", "lux", "int_step", "next_x", "next_u", "next_int_step")
    _STATUS = ("traj_costs", "alpha", "mu", "iterations", "lg_iterations", "mu_step",
               "trajectory_changed", "improved", "termination_condition")
    _SETTINGS = ("dt", "max_iterations", "max_lg_iterations", "min_rel_cost_change",
                 "opt_start", "use_quadratic_terms", "integrator_type", "keep_previous", "keep_records", "precision",
                 "line_search_rounds", "single_launch")

    def __init__(self, lib_path, batch=1, scenes=None, horizon_max=None, device=None):
        self._lib_path = lib_path
        self._lib = _cabi.load(lib_path)
        self._info = info = _cabi.model_info(self._lib)
        self.X, self.U, self.C = info["X"], info["U"], info["C"]
        self.batch = int(batch)
        self.scenes = int(scenes) if scenes is not None else self.batch
        self.t_max = int(horizon_max) if horizon_max is not None else _cabi.HORIZON_MAX
        if not (1 <= self.t_max <= _cabi.HORIZON_MAX):
            raise ValueError(f"horizon_max must be in 1..{_cabi.HORIZON_MAX}")
        if device is None:
            if not torch.cuda.is_available():
                raise _cabi.SolverError("tpl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)

        B, S, Tm, X, U, Cn = self.batch, self.scenes, self.t_max, self.X, self.U, self.C
        f64 = dict(dtype=torch.float64, device=self.device)
        i32 = dict(dtype=torch.int32, device=self.device)
        z = torch.zeros
        # trajectories, problem index fastest (include/tplb200.h)
        self._x = z(Tm + 1, X, B, **f64)
        self._u = z(Tm, U, B, **f64)
        self._prev_x = z(Tm + 1, X, B, **f64)
        self._prev_k = z(Tm, U, B, **f64)
        self._prev_u = None                # the reference never writes prev_u (optim.c:844-848): zeros on demand
        self._next = None                  # next_x / next_u, gathered on demand
        self._k = z(Tm, U, B, **f64)
        self._K = z(Tm, U * X, B, **f64)
        self._g = z(Tm, U, B, **f64)
        self._lam = z(Tm, Cn, B, **f64)
        self._bw = torch.ones(Cn, B, **f64)                       # optim.c:1908-1910
        self._lim = torch.full((Cn, B), float("inf"), **f64)      # optim.c:1905-1907
        self._u_min = z(Tm, U, B, **f64)
        self._u_max = z(Tm, U, B, **f64)
        self._u_min[:min(20, Tm)] = -float("inf")                  # only t < 20 (optim.c:1911-1918)
        self._u_max[:min(20, Tm)] = float("inf")
        self._status = {n: z(B, **(f64 if n in ("traj_costs", "alpha", "mu") else i32)) for n in self._STATUS}
        self._scene_index = (torch.arange(B, **i32) if S == B else z(B, **i32))
        self._scalar_index = {n: i for i, n in enumerate(info["scalar_names"])}
        self._array_index = {n: i for i, n in enumerate(info["array_names"])}
        self._scalars = z(max(1, len(self._scalar_index)), S, **f64)
        self._array_ndim = list(info["array_ndim"])
        self._arrays = [z(*((S,) + (0,) * nd), **f64) for nd in self._array_ndim]
        self._horizons = None
        self._workspace = None
        self._deriv_dense = None
        self._deriv_stale = True
        self._dirty = True
        self._events = None

        # settings (optim.c:1899-1904)
        self.dt = 0.05
        self._T = min(20, Tm)
        self.min_rel_cost_change = 1e-6
        self.max_iterations = 5
        self.max_lg_iterations = 1
        self.use_quadratic_terms = True
        self.opt_start = 0
        self.integrator_type = EULER
        self.keep_previous = True          # maintain prev_x / prev_k on accepted steps (optim.c:844-845)
        self.keep_records = True           # fx..lux readable after update() (always true for small batches)
        self.precision = "fp64"            # "fp32": kernels compute in single precision
        self.line_search_rounds = 0        # 0 auto, 1 all step sizes at once, 2 in two rounds (tplb200.h)
        self.single_launch = 0             # 0 auto, 1 whole update() in one launch (CTA per problem), -1 never
        self.params = BatchedParams(self)

    # -- settings -----------------------------------------------------------------
    @property
    def horizon(self):
        return self._T

    @horizon.setter
    def horizon(self, v):
        self._T = min(self.t_max, max(1, int(v)))                 # optim.c:1726-1734
        if getattr(self, "_horizons", None) is not None and int(self._horizons.max()) > self._T:
            self._horizons = torch.clamp(self._horizons, max=self._T)

    T = horizon

    @property
    def horizons(self):
        """Optional per-problem horizons ``T_b`` (B,), each in 1..``horizon`` — every reference
        ``Optim`` object owns its T (``PathOptim`` sets it from the length of the local map every
        cycle, path_optim.py:126).  ``None``: all problems use ``horizon``.  Array attributes keep
        the shape of the largest horizon; rows past a problem's own horizon are not touched."""
        return self._horizons

    @horizons.setter
    def horizons(self, v):
        if v is None:
            self._horizons = None
            return
        h = torch.as_tensor(np.asarray(v.cpu() if isinstance(v, torch.Tensor) else v), dtype=torch.int32)
        if h.shape != (self.batch,):
            raise ValueError(f'Expected "horizons" with shape ({self.batch}), but found {tuple(h.shape)}')
        if int(h.min()) < 1 or int(h.max()) > self._T:
            raise ValueError(f"horizons must lie in 1..horizon (= {self._T}); set `horizon` to the largest one first")
        self._horizons = h.to(self.device)

    @property
    def step(self):
        return self.dt

    @step.setter
    def step(self, v):
        self.dt = float(v)

    @property
    def scene_index(self):
        return self._scene_index

    @scene_index.setter
    def scene_index(self, v):
        idx = torch.as_tensor(np.asarray(v), dtype=torch.int32).to(self.device)
        if idx.shape != (self.batch,):
            raise ValueError(f'Expected "scene_index" with shape ({self.batch}), but found {tuple(idx.shape)}')
        if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= self.scenes):
            raise ValueError("scene_index out of range")
        self._scene_index.copy_(idx)

    @property
    def slots(self):
        """Names the reference lists in ``Optim.__slots__`` (optim.c:1736-1782)."""
        return ["step", "opt_start", "horizon", "int_step", "x", "u", "next_int_step", "next_x", "next_u",
                "prev_x", "prev_u", "prev_k", "fx", "fu", "lx", "lu", "lxx", "luu", "lux", "g", "k", "K",
                "lagrange_multiplier", "barrier_weight", "lg_mult_limit", "u_max", "u_min", "params",
                "traj_costs", "iterations", "lg_iterations", "runtime", "alpha", "mu", "mu_step",
                "trajectory_changed", "improved", "termination_condition", "use_quadratic_terms",
                "max_iterations", "max_lg_iterations", "min_rel_cost_change"]

    # -- array views ----------------------------------------------------------------
    def _traj_view(self, buf, rows, comp_shape):
        v = buf[:rows].permute(2, 0, 1)                           # (B, rows, comps), zero-copy
        comp_shape = [d for d in comp_shape if d != 1]           # squeeze (optim.c:1319-1325)
        if len(comp_shape) == 0:
            return v[:, :, 0] if v.shape[2] == 1 else v.reshape(v.shape[0], v.shape[1])
        if len(comp_shape) == 2:
            return v.unflatten(2, comp_shape)
        return v

    def _deriv_view(self, name, comp_shape):
        """fx..lux: the solver keeps only the entries that are not identically 0/1;
        the dense records the reference exposes are expanded on demand."""
        if self._deriv_dense is None:
            self._deriv_dense = torch.zeros(self.t_max, self._info["deriv_stride"], self.batch,
                                            dtype=torch.float64, device=self.device)
            self._deriv_stale = True
        if self._deriv_stale and self.device.type == "cuda":
            with torch.cuda.device(self.device):
                q = self._descriptor()
                _cabi.check(self._lib, self._lib.tplb_expand_derivatives(C.byref(q), self._stream()),
                            "tplb_expand_derivatives")
            self._deriv_stale = False
        off = self._info["offsets"][name]
        n = int(np.prod(comp_shape))
        blk = self._deriv_dense[:self._T, off:off + n, :]
        return self._traj_view(blk, self._T, comp_shape)

    def __getattr__(self, n):
        # only reached for names without a real attribute / property
        if n.startswith("_"):
            raise AttributeError(n)
        T, X, U, Cn = self._T, self.X, self.U, self.C
        if n == "x":
            return self._traj_view(self._x, T + 1, (X,))
        if n == "u":
            return self._traj_view(self._u, T, (U,))
        if n == "prev_x":
            return self._traj_view(self._prev_x, T + 1, (X,))
        if n == "prev_k":
            return self._traj_view(self._prev_k, T, (U,))
        if n == "prev_u":
            if self._prev_u is None:
                self._prev_u = torch.zeros(self.t_max, U, self.batch, dtype=torch.float64, device=self.device)
            return self._traj_view(self._prev_u, T, (U,))
        if n in ("next_x", "next_u"):                             # optim.c:1657-1659
            nx, nu = self._next_trajectory()
            return self._traj_view(nx, T + 1, (X,)) if n == "next_x" else self._traj_view(nu, T, (U,))
        if n == "k":
            return self._traj_view(self._k, T, (U,))
        if n == "g":
            return self._traj_view(self._g, T, (U,))
        if n == "K":
            return self._traj_view(self._K, T, (U, X))
        if n == "lagrange_multiplier":
            return self._traj_view(self._lam, T, (Cn,))
        if n == "u_min":
            return self._traj_view(self._u_min, T, (U,))
        if n == "u_max":
            return self._traj_view(self._u_max, T, (U,))
        if n == "barrier_weight":
            return self._bw.t()
        if n == "lg_mult_limit":
            return self._lim.t()
        if n in ("fx", "lxx"):
            return self._deriv_view(n, (X, X))
        if n == "fu":
            return self._deriv_view(n, (X, U))
        if n == "lx":
            return self._deriv_view(n, (X,))
        if n == "lu":
            return self._deriv_view(n, (U,))
        if n == "luu":
            return self._deriv_view(n, (U, U))
        if n == "lux":
            return self._deriv_view(n, (U, X))
        if n in ("int_step", "next_int_step"):                    # calcIntStep is always dt (optim.c:636-651)
            return torch.full((self.batch, T + 1), self.dt, dtype=torch.float64, device=self.device)
        if n in self._STATUS:
            return self._status[n]
        if n in ("runtime", "elapsed_update_time"):
            return self._runtime_ms()
        if n == "__slots__":
            return self.slots
        raise AttributeError(n)

    def __setattr__(self, n, v):
        if n in ("next_x", "next_u", "int_step", "next_int_step"):
            raise AttributeError(f"{n} is derived from the last line search and cannot be assigned")
        if n in self._FIELDS:
            view = getattr(self, n)
            if isinstance(v, (int, float)):
                view.fill_(float(v))
            else:
                view.copy_(_fit(_source(v), view), non_blocking=True)
        elif n in self._STATUS:
            t = self._status[n]
            if isinstance(v, (int, float)):                       # no host tensor: also legal inside a CUDA graph
                t.fill_(v)
            else:
                t.copy_(_as_tensor(v, self.device, t.dtype).expand_as(t))
        else:
            object.__setattr__(self, n, v)

    def __getitem__(self, i):
        if not (-self.batch <= i < self.batch):
            raise IndexError(i)
        return ProblemView(self, i % self.batch)

    def __len__(self):
        return self.batch

    def _next_trajectory(self):
        """The candidates the last line searches ended on, gathered into [t][i][B] buffers."""
        if self._next is None:
            self._next = (torch.zeros(self.t_max + 1, self.X, self.batch, dtype=torch.float64, device=self.device),
                          torch.zeros(self.t_max, self.U, self.batch, dtype=torch.float64, device=self.device))
        if self.device.type == "cuda" and self._workspace is not None:
            with torch.cuda.device(self.device):
                q = self._descriptor()
                _cabi.check(self._lib, self._lib.tplb_next_trajectory(
                    C.byref(q), self._next[0].data_ptr(), self._next[1].data_ptr(), self._stream()),
                    "tplb_next_trajectory")
        return self._next

    def set_initial_state(self, x0):
        """``opt.x[:, 0] = x0`` for host or device ``x0`` of shape (B, X)."""
        self._x[0].copy_(_source(x0).expand(self.batch, self.X).t(), non_blocking=True)

    # -- host staging (HostMirror) -------------------------------------------------------------
    def host_mirror(self):
        """Pinned host buffers for this solver's inputs and results (see ``HostMirror``)."""
        return HostMirror(self)

    def upload(self, m, params=True):
        """Initial state, control warm start and (optionally) all parameters from the pinned
        mirror ``m``: contiguous asynchronous copies on the current stream."""
        self._require_cuda("upload()")
        T = self._T
        self._x[0].copy_(m._x0, non_blocking=True)
        self._u[:T].copy_(m._u0, non_blocking=True)
        if params:
            self._scalars.copy_(m.scalars, non_blocking=True)
            for n, i in self._array_index.items():
                self._arrays[i].copy_(m.arrays[n], non_blocking=True)

    def download(self, m):
        """Trajectories, costs, iteration counts and termination flags into the pinned mirror
        ``m`` (asynchronous; synchronise the stream before reading them on the host)."""
        self._require_cuda("download()")
        T = self._T
        m._x.copy_(self._x[:T + 1], non_blocking=True)
        m._u.copy_(self._u[:T], non_blocking=True)
        m.traj_costs.copy_(self._status["traj_costs"], non_blocking=True)
        m._flags[0].copy_(self._status["iterations"], non_blocking=True)
        m._flags[1].copy_(self._status["termination_condition"], non_blocking=True)

    # -- C ABI plumbing ---------------------------------------------------------------
    def _ensure_workspace(self):
        if self._workspace is None:
            nbytes = self._lib.tplb_workspace_bytes(self.batch, self.scenes, self.t_max)
            self._workspace = torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=self.device)
            self._workspace_bytes = nbytes

    def _descriptor(self):
        self._ensure_workspace()
        q = _cabi.Batch()
        q.struct_bytes = C.sizeof(_cabi.Batch)
        q.batch, q.scenes, q.horizon, q.t_max = self.batch, self.scenes, self._T, self.t_max
        q.opt_start = int(self.opt_start)
        q.max_iterations = int(self.max_iterations)
        q.max_lg_iterations = int(self.max_lg_iterations)
        q.integrator_type = int(self.integrator_type)
        q.use_quadratic_terms = int(bool(self.use_quadratic_terms))
        q.keep_previous = int(bool(self.keep_previous))
        q.keep_records = int(bool(self.keep_records))
        q.precision = {"fp64": 0, "fp32": 1}[self.precision]
        q.line_search_rounds = int(self.line_search_rounds)
        q.single_launch = int(self.single_launch)
        q.dt = float(self.dt)
        q.min_rel_cost_change = float(self.min_rel_cost_change)
        for name, t in (("x", self._x), ("u", self._u), ("prev_x", self._prev_x), ("prev_k", self._prev_k),
                        ("k", self._k), ("K", self._K), ("g", self._g), ("lagrange_multiplier", self._lam),
                        ("barrier_weight", self._bw), ("lg_mult_limit", self._lim),
                        ("u_min", self._u_min), ("u_max", self._u_max),
                        ("scene_index", self._scene_index), ("scalars", self._scalars)):
            setattr(q, name, t.data_ptr())
        for name in self._STATUS:
            setattr(q, name, self._status[name].data_ptr())
        for i, a in enumerate(self._arrays):
            q.arrays[i] = a.data_ptr()
            q.array_len[i] = a.numel() // a.shape[0]
            q.array_cols[i] = a.shape[2] if a.ndim == 3 else 0
        q.workspace = self._workspace.data_ptr()
        q.workspace_bytes = self._workspace_bytes
        q.deriv_dense = self._deriv_dense.data_ptr() if self._deriv_dense is not None else None
        q.horizons = self._horizons.data_ptr() if self._horizons is not None else None
        return q

    def _require_cuda(self, what):
        if self.device.type != "cuda":
            raise _cabi.SolverError(f"{what} needs the CUDA solver; buffers are on {self.device} (no CPU fallback)")

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- methods (optim.c:1884-1892) ------------------------------------------------------
    def update(self):
        """One reference ``update()`` for every problem; asynchronous on the
        current CUDA stream."""
        self._require_cuda("update()")
        with torch.cuda.device(self.device):
            q = self._descriptor()
            if torch.cuda.is_current_stream_capturing():   # inside a CUDA graph: no timing events
                _cabi.check(self._lib, self._lib.tplb_update(C.byref(q), self._stream()), "tplb_update")
                self._events = None
            else:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _cabi.check(self._lib, self._lib.tplb_update(C.byref(q), self._stream()), "tplb_update")
                e1.record()
                self._events = (e0, e1)
            self._deriv_stale = True

    def update_profiled(self):
        """``update()`` with a CUDA-event pair around every kernel launch: returns
        ``{kernel class: (milliseconds, launches)}``.  Measurement aid (synchronises
        after each launch); results are identical to ``update()``."""
        self._require_cuda("update_profiled()")
        n = len(_cabi.KERNEL_CLASSES)
        ms, cnt = (C.c_float * n)(), (C.c_int32 * n)()
        with torch.cuda.device(self.device):
            q = self._descriptor()
            _cabi.check(self._lib, self._lib.tplb_update_profiled(C.byref(q), self._stream(), ms, cnt),
                        "tplb_update_profiled")
            self._deriv_stale = True
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_cabi.KERNEL_CLASSES)}

    def work_counters(self):
        """(linearisations, backward sweeps, sequential-equivalent rollouts) of the
        last ``update()``, each a (B,) int32 device tensor."""
        self._require_cuda("work_counters()")
        self._ensure_workspace()
        base = self._lib.tplb_workspace_counters(self._workspace.data_ptr(), self.batch, self.scenes, self.t_max)
        off = (base - self._workspace.data_ptr()) // 4
        c = self._workspace.view(torch.int32)[off:off + 3 * self.batch].view(3, self.batch)
        return c[0], c[1], c[2]

    def linearize(self):
        """Fill ``fx, fu, lx, lu, lxx, luu, lux`` for the current trajectory."""
        self._require_cuda("linearize()")
        if self._deriv_dense is None:
            self._deriv_dense = torch.zeros(self.t_max, self._info["deriv_stride"], self.batch,
                                            dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            q = self._descriptor()
            _cabi.check(self._lib, self._lib.tplb_linearize(C.byref(q), self._stream()), "tplb_linearize")
        self._deriv_stale = False

    def shift(self, amount):
        """Warm-start shift; ``amount`` is an int or a (B,) integer array."""
        self._require_cuda("shift()")
        with torch.cuda.device(self.device):
            q = self._descriptor()
            if isinstance(amount, (int, np.integer)):
                code = self._lib.tplb_shift(C.byref(q), int(amount), None, self._stream())
            else:
                a = torch.as_tensor(np.asarray(amount), dtype=torch.int32).to(self.device)
                if a.shape != (self.batch,):
                    raise ValueError(f'Expected "amount" with shape ({self.batch}), but found {tuple(a.shape)}')
                code = self._lib.tplb_shift(C.byref(q), 0, a.data_ptr(), self._stream())
            _cabi.check(self._lib, code, "tplb_shift")

    def _points(self, x, u, t, dt, continuous):
        self._require_cuda("dynamics()")
        x = _as_tensor(x, self.device)
        u = _as_tensor(u, self.device)
        single = x.ndim == 1
        if single:
            x, u = x.unsqueeze(0), u.reshape(1, -1)
        if u.ndim == 1:
            u = u.unsqueeze(1)
        if x.ndim != 2 or x.shape[1] != self.X:
            raise ValueError(f'Expected "x_arr" with shape ({self.X}), but found {tuple(x.shape[1:] if not single else x.shape[1:])}')
        if u.ndim != 2 or u.shape[1] != self.U or u.shape[0] != x.shape[0]:
            raise ValueError(f'Expected "u_arr" with shape ({self.U}), but found {tuple(u.shape[1:])}')
        n = x.shape[0]
        xin, uin = x.t().contiguous(), u.t().contiguous()
        out = torch.empty_like(xin)
        with torch.cuda.device(self.device):
            q = self._descriptor()
            _cabi.check(self._lib, self._lib.tplb_dynamics(
                C.byref(q), xin.data_ptr(), uin.data_ptr(), None, n, int(t), float(dt),
                int(continuous), out.data_ptr(), self._stream()), "tplb_dynamics")
        res = out.t()
        return res[0] if single else res

    def dynamics_soa(self, x, u, t, dt, out=None, continuous=False):
        """``dynamics`` on device tensors already in the solver's layout: ``x`` (X, n), ``u`` (U, n)
        -> (X, n), no transposes or host traffic (the MPC's dead-time roll-forward calls it
        ~18 times per control cycle, model_predictive_controller_time.py:159-171).  ``out`` may
        alias ``x``."""
        self._require_cuda("dynamics_soa()")
        if x.shape[0] != self.X or u.shape != (self.U, x.shape[1]) or not (x.is_contiguous() and u.is_contiguous()):
            raise ValueError(f'Expected contiguous x ({self.X}, n) and u ({self.U}, n), found {tuple(x.shape)}, {tuple(u.shape)}')
        out = torch.empty_like(x) if out is None else out
        with torch.cuda.device(self.device):
            q = self._descriptor()
            _cabi.check(self._lib, self._lib.tplb_dynamics(
                C.byref(q), x.data_ptr(), u.data_ptr(), None, x.shape[1], int(t), float(dt),
                int(continuous), out.data_ptr(), self._stream()), "tplb_dynamics")
        return out

    def dynamics(self, x, u, t, dt):
        """Discrete dynamics with the current integrator: ``(X,)`` or ``(n, X)``
        states (point i uses the scene of problem i when n == batch)."""
        return self._points(x, u, t, dt, 0)

    def ct_dynamics(self, x, u, t, dt):
        return self._points(x, u, t, dt, 1)

    def shift_interp(self, arc_len):
        """Warm start between planning cycles (velocity_optim.py:163-168): ``x[:-1]`` and the
        multipliers are resampled linearly, ``u`` with a zero-order hold, at ``ss + arc_len`` on the
        grid ``ss = i * step`` — per problem ``arc_len`` of shape (B,).  On the device, in place."""
        from . import prep
        if self.device.type != "cuda":
            raise _cabi.SolverError("shift_interp runs on a CUDA device only; there is no CPU fallback")
        T = self._T
        with torch.cuda.device(self.device):
            self._x[:T].copy_(prep.shift_interp_soa(self._x[:T], self.dt, arc_len, "linear"))
            self._u[:T].copy_(prep.shift_interp_soa(self._u[:T], self.dt, arc_len, "zero"))
            if self.C:
                lam = self._lam[:T]
                lam.copy_(prep.shift_interp_soa(lam, self.dt, arc_len, "linear"))

    def argmin_groups(self, per_group):
        """Multi-start reduction: best finite ``traj_costs`` of each contiguous
        group of ``per_group`` problems -> (min_cost, arg_min) device tensors."""
        self._require_cuda("argmin_groups()")
        if self.batch % per_group:
            raise ValueError("batch is not a multiple of per_group")
        groups = self.batch // per_group
        mn = torch.empty(groups, dtype=torch.float64, device=self.device)
        am = torch.empty(groups, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib, self._lib.tplb_argmin_groups(
                self._status["traj_costs"].data_ptr(), groups, per_group, mn.data_ptr(), am.data_ptr(),
                self._stream()), "tplb_argmin_groups")
        return mn, am

    def take(self, indices):
        """Trajectories of the problems ``indices`` (a device index tensor, e.g. the ``arg_min`` of
        ``argmin_groups``): ``(x (n, T+1, X), u (n, T, U))`` views of freshly gathered buffers in
        the solver's layout.  A multi-start caller needs only the winners on the host; downloading
        them instead of the whole batch cuts the device->host traffic by the group size."""
        self._require_cuda("take()")
        T = self._T
        idx = indices.to(device=self.device, dtype=torch.int64)
        return (self._traj_view(self._x[:T + 1].index_select(2, idx), T + 1, (self.X,)),
                self._traj_view(self._u[:T].index_select(2, idx), T, (self.U,)))

    def measure_fp64_tflops(self, repeats=5):
        self._require_cuda("measure_fp64_tflops()")
        with torch.cuda.device(self.device):
            return float(self._lib.tplb_measure_fp64_tflops(repeats, self._stream()))

    def _runtime_ms(self):
        if self._events is None:
            return 0.0
        self._events[1].synchronize()
        return float(self._events[0].elapsed_time(self._events[1]))

    # -- copies / state (optim.c:1784-1820) ------------------------------------------------
    def __deepcopy__(self, memo):
        o = BatchedOptim(self._lib_path, self.batch, self.scenes, self.t_max, self.device)
        for n in ("_x", "_u", "_prev_x", "_prev_k", "_k", "_K", "_g", "_lam", "_bw", "_lim",
                  "_u_min", "_u_max", "_scene_index", "_scalars"):
            getattr(o, n).copy_(getattr(self, n))
        for n in self._STATUS:
            o._status[n].copy_(self._status[n])
        o._arrays = [a.clone() for a in self._arrays]
        for n in self._SETTINGS:
            object.__setattr__(o, n, getattr(self, n))
        o._T = self._T
        o._horizons = None if self._horizons is None else self._horizons.clone()
        if self._workspace is not None:
            o._ensure_workspace()
            o._workspace.copy_(self._workspace)
        if self._deriv_dense is not None:
            o._deriv_dense = self._deriv_dense.clone()
            o._deriv_stale = self._deriv_stale
        return o

    def __getstate__(self):
        return {
            "u_min": self.u_min.cpu().numpy(), "u_max": self.u_max.cpu().numpy(),
            "horizon": self._T, "opt_start": self.opt_start,
            "barrier_weight": self.barrier_weight.cpu().numpy(),
            "lg_mult_limit": self.lg_mult_limit.cpu().numpy(),
            "max_iterations": self.max_iterations, "max_lg_iterations": self.max_lg_iterations,
            "step": self.dt, "use_quadratic_terms": bool(self.use_quadratic_terms),
            "params": self.params.__getstate__(),
        }
