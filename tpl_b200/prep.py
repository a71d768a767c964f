"""Batched profile shaping in front of the lateral / velocity solves (SURVEY.md section 8,
row f2): the reference's two `rampify_profile` functions with a leading batch dimension,
on the GPU through the C ABI of ``include/tplb200_prep.h``.

    rampify_velocity_profile(v0, a0, lim_v, a_min, a_max, j_min, j_max, v_min, step)
        reference: library/tpl/planning/utils.py:5-65, arguments in the same order;
        lim_v (B, N), v0 / a0 (B,) or None  ->  profile (B, N, 2)
    rampify_lateral_profile(step, horizon, evasion_sharpness, proj_distance, path, gap, lower, upper)
        reference: library/tpl/planning/path_vel_decomp/path_optim.py:11-55;
        path (B, N, >=6) (column 5 is used) or (B, N) = that column, lower / upper (B, N),
        proj_distance (B,)  ->  d_offset (B, N)

    shift_interp(arr, step, arc_len, kind="linear")
        reference: VelocityOptim.shift_interp (planning/path_vel_decomp/velocity_optim.py:86-104), i.e.
        scipy interp1d(ss, arr, kind, axis=0, fill_value="extrapolate")(ss + arc_len) with ss = i*step;
        arr (B, n) or (B, n, rows), arc_len (B,)  ->  same shape as arr
    shift_interp_soa(buf, step, arc_len, kind)
        the same on a solver buffer [n][rows][B] (no transposes; used by BatchedOptim.shift_interp)

There is no CPU fallback: the calls raise without the CUDA library or a CUDA device."""

import ctypes as C
import os

import numpy as np
import torch

_D = C.c_void_p
EXPORTS = ("tplb_prep_abi_version", "tplb_prep_last_error", "tplb_rampify_velocity", "tplb_rampify_lateral",
           "tplb_shift_interp", "tplb_update_ego", "tplb_project", "tplb_resample_scratch_doubles",
           "tplb_resample_path", "tplb_frenet_to_cartesian")
PROJECTION_FIELDS = ("distance", "arc_len", "alpha", "index", "start", "end", "point_x", "point_y",
                     "tangent_x", "tangent_y", "angle", "in_bounds")
KINDS = {"linear": 0, "zero": 1}
ABI_VERSION = 2


class PrepError(RuntimeError):
    pass


def library_path():
    from . import build
    return os.path.join(build.default_lib_dir(), "libtplb200_prep.so")


_LIB = None


def load(path=None):
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or library_path()
    if not os.path.exists(path):
        raise PrepError(f"{path} is missing: build it with `python -m tpl_b200.build` (needs nvcc)")
    lib = C.CDLL(path)
    lib.tplb_prep_abi_version.restype = C.c_int32
    lib.tplb_prep_last_error.restype = C.c_char_p
    lib.tplb_rampify_velocity.argtypes = [C.c_int32, C.c_int32, _D, _D, _D] + [C.c_double] * 6 + [_D, C.c_void_p]
    lib.tplb_rampify_velocity.restype = C.c_int32
    lib.tplb_rampify_lateral.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, _D, _D,
                                         C.c_double, _D, _D, _D, C.c_void_p]
    lib.tplb_rampify_lateral.restype = C.c_int32
    lib.tplb_shift_interp.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_double, _D, C.c_int32, _D, _D, C.c_void_p]
    lib.tplb_shift_interp.restype = C.c_int32
    lib.tplb_project.argtypes = [C.c_int32, C.c_int32, C.c_int32, _D, _D, C.c_int32, _D, C.c_void_p]
    lib.tplb_project.restype = C.c_int32
    lib.tplb_resample_scratch_doubles.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.tplb_resample_scratch_doubles.restype = C.c_size_t
    lib.tplb_resample_path.argtypes = [C.c_int32, C.c_int32, _D, C.c_double, C.c_int32, _D, C.c_int32, C.c_int32,
                                       _D, _D, _D, C.c_void_p]
    lib.tplb_resample_path.restype = C.c_int32
    lib.tplb_frenet_to_cartesian.argtypes = [C.c_int32, C.c_int32, C.c_int32, _D, _D, C.c_void_p]
    lib.tplb_frenet_to_cartesian.restype = C.c_int32
    if lib.tplb_prep_abi_version() != ABI_VERSION:
        raise PrepError(f"{path}: ABI version {lib.tplb_prep_abi_version()} != {ABI_VERSION}")
    _LIB = lib
    return lib


def _device(device):
    if not torch.cuda.is_available():
        raise PrepError("profile shaping runs on a CUDA device only; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _rows(value, device, name, batch=None):
    """(B, N) host or device array -> [N][B] device tensor (problem index fastest)."""
    t = value if isinstance(value, torch.Tensor) else torch.as_tensor(np.asarray(value, dtype=np.float64))
    t = t.to(device=device, dtype=torch.float64, non_blocking=True)
    if t.ndim == 1:
        t = t.unsqueeze(0)
    if t.ndim != 2 or (batch is not None and t.shape[0] not in (1, batch)):
        raise ValueError(f'Expected "{name}" with shape (B, N), but found {tuple(t.shape)}')
    if batch is not None and t.shape[0] == 1:
        t = t.expand(batch, -1)
    return t.t().contiguous()


def _per_problem(value, device, batch, name):
    t = value if isinstance(value, torch.Tensor) else torch.as_tensor(np.asarray(value, dtype=np.float64))
    t = t.to(device=device, dtype=torch.float64, non_blocking=True).reshape(-1)
    if t.numel() not in (1, batch):
        raise ValueError(f'Expected "{name}" with shape ({batch},), but found {tuple(t.shape)}')
    return t.expand(batch).contiguous()


def _check(lib, rc, what):
    if rc != 0:
        raise PrepError(f"{what}: {lib.tplb_prep_last_error().decode()}")


def rampify_velocity_profile(v0, a0, lim_v, a_min, a_max, j_min, j_max, v_min, step, device=None):
    lib = load()
    dev = _device(device)
    with torch.cuda.device(dev):
        lim = _rows(lim_v, dev, "lim_v")
        n, batch = lim.shape
        v0_t = None if v0 is None else _per_problem(v0, dev, batch, "v0")
        a0_t = None if a0 is None else _per_problem(a0, dev, batch, "a0")
        out = torch.empty((n, 2, batch), dtype=torch.float64, device=dev)
        rc = lib.tplb_rampify_velocity(batch, n, None if v0_t is None else v0_t.data_ptr(),
                                       None if a0_t is None else a0_t.data_ptr(), lim.data_ptr(),
                                       float(a_min), float(a_max), float(j_min), float(j_max), float(v_min),
                                       float(step), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _check(lib, rc, "tplb_rampify_velocity")
    return out.permute(2, 0, 1)


def rampify_lateral_profile(step, horizon, evasion_sharpness, proj_distance, path, gap, lower, upper, device=None):
    lib = load()
    dev = _device(device)
    with torch.cuda.device(dev):
        lo = _rows(lower, dev, "lower")
        n, batch = lo.shape
        up = _rows(upper, dev, "upper", batch)
        p = path if isinstance(path, torch.Tensor) else torch.as_tensor(np.asarray(path, dtype=np.float64))
        if p.ndim == 3 or (p.ndim == 2 and p.shape[-1] != n):
            p = p[..., 5]                                    # the reference reads path[i, 5] only
        pv = _rows(p, dev, "path", batch)
        if up.shape != lo.shape or pv.shape != lo.shape:
            raise ValueError("path, lower and upper must have the same number of samples")
        proj = _per_problem(proj_distance, dev, batch, "proj_distance")
        out = torch.empty((n, batch), dtype=torch.float64, device=dev)
        rc = lib.tplb_rampify_lateral(batch, n, int(horizon), float(step), float(evasion_sharpness),
                                      proj.data_ptr(), pv.data_ptr(), float(gap), lo.data_ptr(), up.data_ptr(),
                                      out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _check(lib, rc, "tplb_rampify_lateral")
    return out.t()


def shift_interp_soa(buf, step, arc_len, kind="linear"):
    """``buf``: contiguous CUDA tensor [n][rows][B] -> new tensor of the same shape."""
    lib = load()
    if not (isinstance(buf, torch.Tensor) and buf.is_cuda and buf.is_contiguous() and buf.ndim == 3
            and buf.dtype == torch.float64):
        raise PrepError("shift_interp_soa needs a contiguous fp64 CUDA tensor [n][rows][B]")
    n, rows, batch = buf.shape
    with torch.cuda.device(buf.device):
        off = _per_problem(arc_len, buf.device, batch, "arc_len")
        out = torch.empty_like(buf)
        rc = lib.tplb_shift_interp(batch, n, rows, float(step), off.data_ptr(), KINDS[kind], buf.data_ptr(),
                                   out.data_ptr(), torch.cuda.current_stream(buf.device).cuda_stream)
        _check(lib, rc, "tplb_shift_interp")
    return out


def shift_interp(arr, step, arc_len, kind="linear", device=None):
    dev = _device(device)
    t = arr if isinstance(arr, torch.Tensor) else torch.as_tensor(np.asarray(arr, dtype=np.float64))
    t = t.to(device=dev, dtype=torch.float64, non_blocking=True)
    squeeze = t.ndim == 2
    if squeeze:
        t = t.unsqueeze(-1)
    if t.ndim != 3:
        raise ValueError(f'Expected "arr" with shape (B, n) or (B, n, rows), but found {tuple(t.shape)}')
    out = shift_interp_soa(t.permute(1, 2, 0).contiguous(), step, arc_len, kind).permute(2, 0, 1)
    return out.squeeze(-1) if squeeze else out


# ---- row f1: reference-path preparation -------------------------------------------------------------
def _paths(paths, dev, min_cols):
    t = paths if isinstance(paths, torch.Tensor) else torch.as_tensor(np.asarray(paths, dtype=np.float64))
    t = t.to(device=dev, dtype=torch.float64, non_blocking=True)
    if t.ndim == 2:
        t = t.unsqueeze(0)
    if t.ndim != 3 or t.shape[2] < min_cols:
        raise ValueError(f'Expected "path" with shape (B, P, >={min_cols}), but found {tuple(t.shape)}')
    return t.contiguous()


def project(paths, position, closed=False, device=None):
    """``util.project`` (library/src/utils.cpp:257-408) for B (path, position) pairs: paths (B, P, >=2),
    position (B, 2) -> dict of (B,) tensors named like the reference's ``Projection`` members
    (``distance, arc_len, alpha, index, start, end, point_x, ...``); the MPC takes ``x0[5] = arc_len``
    from it (control/model_predictive_controller.py:188-189)."""
    lib = load()
    dev = _device(device)
    with torch.cuda.device(dev):
        p = _paths(paths, dev, 2)
        batch, points, stride = p.shape
        pos = (position if isinstance(position, torch.Tensor)
               else torch.as_tensor(np.asarray(position, dtype=np.float64))).to(dev, torch.float64).reshape(-1, 2).contiguous()
        if pos.shape[0] != batch:
            raise ValueError(f'Expected "position" with shape ({batch}, 2), but found {tuple(pos.shape)}')
        out = torch.empty((len(PROJECTION_FIELDS), batch), dtype=torch.float64, device=dev)
        rc = lib.tplb_project(batch, points, stride, p.data_ptr(), pos.data_ptr(), int(bool(closed)), out.data_ptr(),
                              torch.cuda.current_stream(dev).cuda_stream)
        _check(lib, rc, "tplb_project")
    return dict(zip(PROJECTION_FIELDS, out))


def resample_path(paths, step_size, steps, start_index=0, zero_vel_at_end=False, closed=False, device=None):
    """``util.resample_path`` (library/tpl/util.py:134-191 over tplcpp.resample, library/src/utils.cpp:410-560)
    for B paths (B, P, 6) of x, y, orientation, s, curvature, velocity -> ``(rs, ok)``:
    ``rs`` (6, B, steps) — ``rs[0]`` ... ``rs[5]`` are (B, steps) arrays that can be assigned to solver
    parameters directly (``opt.params.ref_x = rs[0]``); ``rs.permute(1, 2, 0)`` is the reference's
    (steps, 6) array per problem — and ``ok`` (B,) bool, False where the reference returns None."""
    lib = load()
    dev = _device(device)
    with torch.cuda.device(dev):
        p = _paths(paths, dev, 6)
        if p.shape[2] != 6:
            p = p[:, :, :6].contiguous()
        batch, points, _ = p.shape
        start = None
        if not isinstance(start_index, (int, np.integer)) or int(start_index) != 0:
            start = torch.as_tensor(np.broadcast_to(np.asarray(start_index), (batch,)).copy(), dtype=torch.int32).to(dev)
        rs = torch.empty((6, batch, int(steps)), dtype=torch.float64, device=dev)
        ok = torch.empty(batch, dtype=torch.int32, device=dev)
        scratch = torch.empty(lib.tplb_resample_scratch_doubles(batch, points, int(steps)), dtype=torch.float64, device=dev)
        rc = lib.tplb_resample_path(batch, points, p.data_ptr(), float(step_size), int(steps),
                                    None if start is None else start.data_ptr(), int(bool(zero_vel_at_end)),
                                    int(bool(closed)), rs.data_ptr(), ok.data_ptr(), scratch.data_ptr(),
                                    torch.cuda.current_stream(dev).cuda_stream)
        _check(lib, rc, "tplb_resample_path")
    return rs, ok.bool()


def frenet_to_cartesian(paths, opt):
    """planning/path_vel_decomp/path_optim.py:303-305 for the whole batch, in place on ``paths`` (B, n, 6)
    (a contiguous CUDA tensor, n = opt.horizon): the lateral offsets ``opt.x[:, :-1, 0]`` and slopes
    ``opt.x[:, :-1, 1]`` of the solved lateral problems move the reference line to the planned path."""
    lib = load()
    if not (isinstance(paths, torch.Tensor) and paths.is_cuda and paths.is_contiguous() and paths.ndim == 3
            and paths.shape[2] == 6 and paths.dtype == torch.float64):
        raise PrepError("frenet_to_cartesian needs a contiguous fp64 CUDA tensor (B, n, 6)")
    batch, n, _ = paths.shape
    if batch != opt.batch or n > opt.horizon:
        raise ValueError(f"paths {tuple(paths.shape)} do not match the solver (batch {opt.batch}, horizon {opt.horizon})")
    with torch.cuda.device(paths.device):
        rc = lib.tplb_frenet_to_cartesian(batch, n, opt.X, paths.data_ptr(), opt._x.data_ptr(),
                                          torch.cuda.current_stream(paths.device).cuda_stream)
        _check(lib, rc, "tplb_frenet_to_cartesian")
    return paths
