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

There is no CPU fallback: the calls raise without the CUDA library or a CUDA device."""

import ctypes as C
import os

import numpy as np
import torch

_D = C.c_void_p
EXPORTS = ("tplb_prep_abi_version", "tplb_prep_last_error", "tplb_rampify_velocity", "tplb_rampify_lateral")
ABI_VERSION = 1


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
