"""TEST INFRASTRUCTURE — ctypes front end of oracle/prep_oracle.c (one problem per call,
numpy in / numpy out, the reference functions' argument order).  Not imported by tpl_b200."""

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_D = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _lib():
    path = os.path.join(HERE, "lib", "liboracle_prep.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", HERE, "lib/liboracle_prep.so"], check=True)
    lib = C.CDLL(path)
    lib.tplo_rampify_velocity.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, _D] + \
        [C.c_double] * 6 + [_D, _D]
    lib.tplo_rampify_velocity.restype = None
    lib.tplo_rampify_lateral.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _D, C.c_double,
                                         _D, _D, _D, _D, _D]
    lib.tplo_rampify_lateral.restype = None
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _lib()
    return _LIB


def rampify_velocity(v0, a0, lim_v, a_min, a_max, j_min, j_max, v_min, step):
    """planning/utils.py:5-65 -> (n, 2)."""
    lim_v = np.ascontiguousarray(lim_v, dtype=np.float64)
    n = len(lim_v)
    scratch, out = np.empty(n), np.empty((n, 2))
    lib().tplo_rampify_velocity(n, v0 is not None, 0.0 if v0 is None else float(v0),
                                a0 is not None, 0.0 if a0 is None else float(a0), lim_v,
                                a_min, a_max, j_min, j_max, v_min, step, scratch, out)
    return out


def rampify_lateral(step, horizon, evasion_sharpness, proj_distance, path, gap, lower, upper):
    """planning/path_vel_decomp/path_optim.py:11-55 -> (len(path),)."""
    path = np.asarray(path, dtype=np.float64)
    path_v = np.ascontiguousarray(path[:, 5])
    n = len(path_v)
    lower = np.ascontiguousarray(lower, dtype=np.float64)
    upper = np.ascontiguousarray(upper, dtype=np.float64)
    fwd, bwd, out = np.empty(n), np.empty(n), np.empty(n)
    lib().tplo_rampify_lateral(n, int(horizon), step, evasion_sharpness, float(proj_distance), path_v, gap,
                               lower, upper, fwd, bwd, out)
    return out


def shift_interp(arr, step, arc_len, kind="linear"):
    """VelocityOptim.shift_interp (planning/path_vel_decomp/velocity_optim.py:86-104) for one
    problem: scipy.interpolate.interp1d(ss, arr, kind, axis=0, fill_value="extrapolate")(ss + arc_len)
    with ss = arange(0, n*step, step).  scipy (unpinned in library/setup.py:108, 1.18.1 here) is not
    imported; this restates its published algorithm: `linear` = numpy.searchsorted (side left),
    indices clipped to 1..n-1, slope*(x_new - x_lo) + y_lo; `zero` = the previous sample, held
    constant beyond both ends."""
    arr = np.asarray(arr, dtype=np.float64)
    n = arr.shape[0]
    ss = np.arange(n) * float(step)
    xq = ss + float(arc_len)
    if kind == "zero":
        idx = np.clip(np.searchsorted(ss, xq, side="right") - 1, 0, n - 1)
        return arr[idx]
    hi = np.clip(np.searchsorted(ss, xq), 1, n - 1)
    lo = hi - 1
    shape = (-1,) + (1,) * (arr.ndim - 1)
    slope = (arr[hi] - arr[lo]) / (ss[hi] - ss[lo]).reshape(shape)
    return slope * (xq - ss[lo]).reshape(shape) + arr[lo]


class _Ego(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("x", "y", "yaw", "v", "a", "steer_angle", "control_acc", "control_steer",
                                          "acc_dead_time", "steer_dead_time", "wheel_base", "v_ch", "max_v",
                                          "min_v", "max_steer_angle")]


class _History(C.Structure):
    _fields_ = [("len", C.c_int), ("t", C.c_double * 64), ("value", C.c_double * 64)]


class EgoOracle:
    """SimCore.update_ego (simulation/core.py:91-134) for one vehicle; attribute names of the
    reference's `ego` object."""

    def __init__(self, **kw):
        object.__setattr__(self, "_e", _Ego())
        object.__setattr__(self, "_acc", _History())
        object.__setattr__(self, "_steer", _History())
        for k, v in kw.items():
            setattr(self, k, v)

    def __getattr__(self, n):
        return getattr(self._e, n)

    def __setattr__(self, n, v):
        setattr(self._e, n, float(v))

    def update(self, t, dt):
        fn = lib().tplo_update_ego
        fn.argtypes = [C.POINTER(_Ego), C.POINTER(_History), C.POINTER(_History), C.c_double, C.c_double]
        fn.restype = None
        fn(C.byref(self._e), C.byref(self._acc), C.byref(self._steer), float(t), float(dt))


# ---- row f1: project / resample / interp_resampled_path -----------------------------------------
PROJECTION_FIELDS = ("distance", "arc_len", "alpha", "index", "start", "end", "point_x", "point_y",
                     "tangent_x", "tangent_y", "angle", "in_bounds")


def project(points, position, closed=False):
    """library/src/utils.cpp:257-408 -> dict of the Projection fields."""
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64)[:, :2])
    out = np.empty(12)
    fn = lib().tplo_project
    fn.argtypes = [_D, C.c_int, C.c_double, C.c_double, C.c_int, _D]
    fn.restype = None
    fn(pts, len(pts), float(position[0]), float(position[1]), int(bool(closed)), out)
    return dict(zip(PROJECTION_FIELDS, out))


def resample(points, sampling_dist, steps, start_index=0, closed=False):
    """library/src/utils.cpp:410-560 -> (rows, 5) array x, y, alpha, prev, next, or None where the
    reference raises RuntimeError."""
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64)[:, :2])
    rsi = np.zeros((max(int(steps), 1), 5))
    scratch = np.empty(2 * max(len(pts), 1))
    fn = lib().tplo_resample
    fn.argtypes = [_D, C.c_int, C.c_double, C.c_int, C.c_long, C.c_int, _D, _D]
    fn.restype = C.c_int
    n = fn(pts, len(pts), float(sampling_dist), int(steps), int(start_index), int(bool(closed)), scratch, rsi)
    if n < 0:
        return None
    return rsi[:n]


def interp_resampled_path(path, rsi, step_size, steps, zero_vel_at_end=False, closed=False):
    """library/tpl/util.py:155-191 -> (steps, 6)."""
    path = np.ascontiguousarray(path, dtype=np.float64)
    rsi = np.ascontiguousarray(rsi, dtype=np.float64)
    rs = np.zeros((int(steps), 6))
    fn = lib().tplo_interp_resampled_path
    fn.argtypes = [_D, C.c_int, _D, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, _D]
    fn.restype = None
    fn(path, len(path), rsi, len(rsi), float(step_size), int(steps), int(bool(zero_vel_at_end)), int(bool(closed)), rs)
    return rs


def resample_path(path, step_size, steps, start_index=0, zero_vel_at_end=False, closed=False):
    """library/tpl/util.py:134-152."""
    path = np.asarray(path, dtype=np.float64)
    rsi = resample(path[:, :2], step_size, steps, start_index, closed)
    if rsi is None:
        return None
    return interp_resampled_path(path, rsi, step_size, steps, zero_vel_at_end, closed)
