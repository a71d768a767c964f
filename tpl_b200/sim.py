"""Batched closed-loop simulator step: the reference's
``SimCore.update_ego(ego, t, dt)`` (library/tpl/simulation/core.py:91-134) for B vehicles at
once on the GPU — actuator dead time, kinematic bicycle with characteristic velocity, clamps.

``BatchedEgo`` carries the reference's ``ego`` attribute names as (B,) CUDA tensors plus the two
command histories that ``SimCore`` keeps as Python lists.  With it a batch of perturbed
closed-loop scenarios (controller -> vehicle -> controller ...) never leaves the device."""

import ctypes as C
import math

import numpy as np
import torch

from . import prep

_D = C.c_void_p


class _Ego(C.Structure):
    _fields_ = [("struct_bytes", C.c_int32), ("batch", C.c_int32), ("capacity", C.c_int32), ("reserved0", C.c_int32),
                ("x", _D), ("y", _D), ("yaw", _D), ("v", _D), ("a", _D), ("steer_angle", _D),
                ("control_acc", _D), ("control_steer", _D)] + \
               [(n, C.c_double) for n in ("acc_dead_time", "steer_dead_time", "wheel_base", "v_ch", "max_v", "min_v",
                                          "max_steer_angle")] + \
               [("acc_t", _D), ("acc_value", _D), ("acc_len", _D), ("steer_t", _D), ("steer_value", _D),
                ("steer_len", _D)]


class BatchedEgo:
    STATE = ("x", "y", "yaw", "v", "a", "steer_angle", "control_acc", "control_steer")
    PARAMS = ("acc_dead_time", "steer_dead_time", "wheel_base", "v_ch", "max_v", "min_v", "max_steer_angle")

    def __init__(self, batch, capacity=32, device=None, **params):
        self._lib = prep.load()
        lib = self._lib
        lib.tplb_update_ego.argtypes = [C.POINTER(_Ego), C.c_double, C.c_double, C.c_void_p]
        lib.tplb_update_ego.restype = C.c_int32
        self.device = prep._device(device)
        self.batch, self.capacity = int(batch), int(capacity)
        f64 = dict(dtype=torch.float64, device=self.device)
        self._state = {n: torch.zeros(batch, **f64) for n in self.STATE}
        # defaults of the reference's vehicle model (tpl/environment: wheel base 2.9 m, v_ch 30 m/s)
        self.acc_dead_time = self.steer_dead_time = 0.0
        self.wheel_base, self.v_ch = 2.9, 30.0
        self.max_v, self.min_v, self.max_steer_angle = 60.0, 0.0, 0.6
        for k, v in params.items():
            if k not in self.PARAMS:
                raise AttributeError(k)
            setattr(self, k, float(v))
        self._hist = {n: torch.zeros((capacity, batch), **f64) for n in ("acc_t", "acc_value", "steer_t", "steer_value")}
        self._len = {n: torch.zeros(batch, dtype=torch.int32, device=self.device) for n in ("acc_len", "steer_len")}

    def __getattr__(self, n):
        st = self.__dict__.get("_state", {})
        if n in st:
            return st[n]
        raise AttributeError(n)

    def __setattr__(self, n, v):
        st = self.__dict__.get("_state", {})
        if n in st:
            t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, dtype=np.float64))
            st[n].copy_(t.to(dtype=torch.float64).expand(self.batch), non_blocking=True)
        else:
            object.__setattr__(self, n, v)

    def reset_histories(self):
        for t in self._len.values():
            t.zero_()

    def update(self, t, dt):
        """``SimCore.update_ego(ego, t, dt)`` for every vehicle; asynchronous on the current stream."""
        q = _Ego()
        q.struct_bytes, q.batch, q.capacity = C.sizeof(_Ego), self.batch, self.capacity
        for n in self.STATE:
            setattr(q, n, self._state[n].data_ptr())
        for n in self.PARAMS:
            setattr(q, n, float(getattr(self, n)))
        for n, tns in self._hist.items():
            setattr(q, n, tns.data_ptr())
        for n, tns in self._len.items():
            setattr(q, n, tns.data_ptr())
        if dt > 0 and self.capacity < max(self.acc_dead_time // dt, self.steer_dead_time // dt) + 2:
            raise prep.PrepError("history capacity too small for dead_time / dt")
        with torch.cuda.device(self.device):
            rc = self._lib.tplb_update_ego(C.byref(q), float(t), float(dt),
                                           torch.cuda.current_stream(self.device).cuda_stream)
        prep._check(self._lib, rc, "tplb_update_ego")
