"""ctypes mirror of include/tplb200.h — the only place Python touches the C ABI."""

import ctypes as C

ABI_VERSION = 3
MAX_ARRAYS = 16
LINE_SEARCH_STEPS = 8
HORIZON_MAX = 299

_D = C.c_void_p   # device pointer


class ModelInfo(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("X", C.c_int32), ("U", C.c_int32), ("C", C.c_int32),
        ("num_scalars", C.c_int32), ("num_arrays", C.c_int32), ("num_params", C.c_int32),
        ("name", C.c_char_p), ("definition_sha1", C.c_char_p),
        ("state_names", C.POINTER(C.c_char_p)), ("action_names", C.POINTER(C.c_char_p)),
        ("scalar_names", C.POINTER(C.c_char_p)), ("array_names", C.POINTER(C.c_char_p)),
        ("param_order", C.POINTER(C.c_char_p)),
        ("deriv_stride", C.c_int32),
        ("off_fx", C.c_int32), ("off_fu", C.c_int32), ("off_lx", C.c_int32), ("off_lu", C.c_int32),
        ("off_lxx", C.c_int32), ("off_luu", C.c_int32), ("off_lux", C.c_int32),
        ("deriv_compact", C.c_int32), ("num_stage_consts", C.c_int32),
        ("array_ndim", C.POINTER(C.c_int32)), ("num_wrap_pairs", C.c_int32), ("wrap_pairs", C.POINTER(C.c_int32)),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32), ("batch", C.c_int32), ("scenes", C.c_int32),
        ("horizon", C.c_int32), ("t_max", C.c_int32),
        ("opt_start", C.c_int32), ("max_iterations", C.c_int32), ("max_lg_iterations", C.c_int32),
        ("integrator_type", C.c_int32), ("use_quadratic_terms", C.c_int32),
        ("keep_previous", C.c_int32), ("precision", C.c_int32),
        ("line_search_rounds", C.c_int32), ("keep_records", C.c_int32),
        ("single_launch", C.c_int32), ("reserved0", C.c_int32),
        ("dt", C.c_double), ("min_rel_cost_change", C.c_double),
        ("x", _D), ("u", _D), ("prev_x", _D), ("prev_k", _D), ("k", _D), ("K", _D), ("g", _D),
        ("lagrange_multiplier", _D), ("barrier_weight", _D), ("lg_mult_limit", _D),
        ("u_min", _D), ("u_max", _D),
        ("traj_costs", _D), ("alpha", _D), ("mu", _D),
        ("iterations", _D), ("lg_iterations", _D), ("mu_step", _D),
        ("trajectory_changed", _D), ("improved", _D), ("termination_condition", _D),
        ("scene_index", _D), ("scalars", _D),
        ("arrays", _D * MAX_ARRAYS), ("array_len", C.c_int32 * MAX_ARRAYS),
        ("array_cols", C.c_int32 * MAX_ARRAYS),
        ("workspace", _D), ("workspace_bytes", C.c_size_t),
        ("deriv_dense", _D),
        ("horizons", _D),
    ]


#: every symbol include/tplb200.h declares (tests check the library exports all of them)
EXPORTS = (
    "tplb_abi_version", "tplb_model", "tplb_last_error", "tplb_workspace_bytes",
    "tplb_workspace_counters",
    "tplb_update", "tplb_update_profiled", "tplb_linearize", "tplb_expand_derivatives",
    "tplb_next_trajectory", "tplb_shift", "tplb_dynamics", "tplb_argmin_groups", "tplb_selftest_math",
    "tplb_measure_fp64_tflops",
)


class SolverError(RuntimeError):
    pass


def load(path):
    """dlopen a solver library and type its entry points.  Raises if the file is
    missing — there is no CPU fallback."""
    lib = C.CDLL(path)
    lib.tplb_abi_version.restype = C.c_int32
    lib.tplb_model.restype = C.POINTER(ModelInfo)
    lib.tplb_last_error.restype = C.c_char_p
    lib.tplb_workspace_bytes.restype = C.c_size_t
    lib.tplb_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.tplb_workspace_counters.restype = C.c_void_p
    lib.tplb_workspace_counters.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.tplb_update_profiled.restype = C.c_int32
    lib.tplb_update_profiled.argtypes = [C.POINTER(Batch), C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    for fn in ("tplb_update", "tplb_linearize", "tplb_expand_derivatives"):
        getattr(lib, fn).restype = C.c_int32
        getattr(lib, fn).argtypes = [C.POINTER(Batch), C.c_void_p]
    lib.tplb_next_trajectory.restype = C.c_int32
    lib.tplb_next_trajectory.argtypes = [C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tplb_shift.restype = C.c_int32
    lib.tplb_shift.argtypes = [C.POINTER(Batch), C.c_int32, C.c_void_p, C.c_void_p]
    lib.tplb_dynamics.restype = C.c_int32
    lib.tplb_dynamics.argtypes = [C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]
    lib.tplb_argmin_groups.restype = C.c_int32
    lib.tplb_argmin_groups.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tplb_selftest_math.restype = C.c_int32
    lib.tplb_selftest_math.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.tplb_measure_fp64_tflops.restype = C.c_double
    lib.tplb_measure_fp64_tflops.argtypes = [C.c_int32, C.c_void_p]
    if lib.tplb_abi_version() != ABI_VERSION:
        raise SolverError(f"{path}: ABI version {lib.tplb_abi_version()} != {ABI_VERSION}")
    return lib


def model_info(lib):
    """Plain-Python copy of the library's tplb_model_info."""
    m = lib.tplb_model().contents

    def strs(p, n):
        return [p[i].decode() for i in range(n)]

    return dict(
        X=m.X, U=m.U, C=m.C, name=m.name.decode(), definition_sha1=m.definition_sha1.decode(),
        state_names=strs(m.state_names, m.X), action_names=strs(m.action_names, m.U),
        scalar_names=strs(m.scalar_names, m.num_scalars), array_names=strs(m.array_names, m.num_arrays),
        param_order=strs(m.param_order, m.num_params),
        deriv_stride=m.deriv_stride, deriv_compact=m.deriv_compact, num_stage_consts=m.num_stage_consts,
        array_ndim=[int(m.array_ndim[i]) for i in range(m.num_arrays)],
        wrap_pairs=[(int(m.wrap_pairs[2 * i]), int(m.wrap_pairs[2 * i + 1])) for i in range(m.num_wrap_pairs)],
        offsets=dict(fx=m.off_fx, fu=m.off_fu, lx=m.off_lx, lu=m.off_lu,
                     lxx=m.off_lxx, luu=m.off_luu, lux=m.off_lux),
    )


KERNEL_CLASSES = ("stage_consts", "rollout_init", "multiplier", "linearize", "backward",
                  "rollout", "stage_cost", "select", "accept", "finalize", "solo")


def check(lib, code, what):
    if code != 0:
        raise SolverError(f"{what} failed ({code}): {lib.tplb_last_error().decode()}")
