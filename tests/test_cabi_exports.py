"""The drop-in boundary: every symbol include/tplb200.h declares is exported by every
solver library, the ctypes mirror matches the header, and nothing computes on CPU."""

import ctypes
import os
import re

import pytest

from tests import common
from tpl_b200 import _cabi

HEADER = os.path.join(common.ROOT, "include", "tplb200.h")


def declared_functions():
    with open(HEADER) as fd:
        text = fd.read()
    return sorted(set(re.findall(r"TPLB_API\s+[\w\s\*]+?\b(tplb_\w+)\s*\(", text)))


def test_header_and_python_mirror_agree():
    assert declared_functions() == sorted(_cabi.EXPORTS)
    with open(HEADER) as fd:
        text = fd.read()
    assert int(re.search(r"#define TPLB_ABI_VERSION (\d+)", text).group(1)) == _cabi.ABI_VERSION
    assert int(re.search(r"#define TPLB_MAX_ARRAYS (\d+)", text).group(1)) == _cabi.MAX_ARRAYS
    assert int(re.search(r"#define TPLB_HORIZON_MAX (\d+)", text).group(1)) == _cabi.HORIZON_MAX
    n_classes = int(re.search(r"TPLB_NUM_KERNEL_CLASSES = (\d+)", text).group(1))
    assert n_classes == len(_cabi.KERNEL_CLASSES)


def test_every_library_exports_every_symbol(solver_libs):
    from tpl_b200 import optimizers
    assert set(solver_libs) == set(optimizers.CONFIGS)
    for name, path in solver_libs.items():
        lib = ctypes.CDLL(path)
        for sym in declared_functions():
            assert hasattr(lib, sym), f"{name}: {sym} not exported"
        lib = _cabi.load(path)
        info = _cabi.model_info(lib)
        assert info["name"] == name
        cfg = optimizers.CONFIGS[name]()
        assert info["X"] == len(cfg.states) and info["U"] == len(cfg.actions)
        assert info["C"] == len(cfg.constraints)
        assert info["param_order"] == [p.name for p in cfg.param_symbols]
        assert info["definition_sha1"] == cfg.definition_hash()
        X, U = info["X"], info["U"]
        assert info["deriv_stride"] == 2 * X * X + 2 * X * U + X + U + U * U
        assert 0 < info["deriv_compact"] <= info["deriv_stride"]
        # host-only queries work without a GPU
        assert lib.tplb_workspace_bytes(64, 64, 50) > 0


def test_model_dimensions_match_reference_table(solver_libs):
    """SURVEY.md appendix C."""
    want = {"trajectory_tracking_mpc": (7, 2, 4), "trajectory_tracking_mpc_time": (6, 2, 4),
            "lateral_profile": (2, 1, 2), "velocity_profile_space": (2, 1, 5),
            "ref_line_smoother_k": (3, 1, 0), "ref_line_smoother_dk": (4, 1, 0),
            "velocity_profile_time": (2, 1, 4)}
    for name, dims in want.items():
        info = _cabi.model_info(_cabi.load(solver_libs[name]))
        assert (info["X"], info["U"], info["C"]) == dims


def test_argument_validation_without_gpu(solver_libs):
    """Rejected descriptors never reach a launch, so this runs on a CPU-only box."""
    lib = _cabi.load(solver_libs["lateral_profile"])
    q = _cabi.Batch()
    assert lib.tplb_update(ctypes.byref(q), None) == -4          # TPLB_E_ABI: struct_bytes unset
    q.struct_bytes = ctypes.sizeof(_cabi.Batch)
    assert lib.tplb_update(ctypes.byref(q), None) == -1          # TPLB_E_ARG
    q.batch, q.scenes, q.horizon, q.t_max = 4, 4, 400, 400
    assert lib.tplb_update(ctypes.byref(q), None) == -2          # TPLB_E_HORIZON
    assert b"horizon" in lib.tplb_last_error()
    q.horizon, q.t_max, q.opt_start = 20, 20, 3
    assert lib.tplb_update(ctypes.byref(q), None) == -3          # TPLB_E_UNSUPPORTED


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under tpl_b200/ may reference it."""
    pkg = os.path.join(common.ROOT, "tpl_b200")
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(root, fn)) as fd:
                    text = fd.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{fn} imports oracle"
                assert "liboracle" not in text and "ilqr_oracle" not in text, fn


def _c_layout(tmp_path, header, struct, fields):
    """sizeof and offsetof of a header struct as a plain C compiler sees them."""
    import subprocess
    src = tmp_path / f"{struct}.c"
    lines = [f'#include <stdio.h>\n#include <stddef.h>\n#include "{header}"\nint main(void) {{',
             f'  printf("%zu\\n", sizeof({struct}));']
    lines += [f'  printf("%zu\\n", offsetof({struct}, {f}));' for f in fields]
    lines.append("  return 0;\n}")
    src.write_text("\n".join(lines))
    exe = tmp_path / f"{struct}.out"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    return int(out[0]), [int(v) for v in out[1:]]


def test_struct_layouts_match_a_c_compiler(tmp_path):
    """The headers are plain C99, and the ctypes mirrors have the compiler's size and offsets."""
    inc = os.path.join(common.ROOT, "include")
    fields = [n for n, _ in _cabi.Batch._fields_]
    size, offs = _c_layout(tmp_path, os.path.join(inc, "tplb200.h"), "tplb_batch", fields)
    assert size == ctypes.sizeof(_cabi.Batch)
    assert offs == [getattr(_cabi.Batch, n).offset for n in fields]
    fields = [n for n, _ in _cabi.ModelInfo._fields_]
    size, offs = _c_layout(tmp_path, os.path.join(inc, "tplb200.h"), "tplb_model_info", fields)
    assert size == ctypes.sizeof(_cabi.ModelInfo)
    assert offs == [getattr(_cabi.ModelInfo, n).offset for n in fields]
    from tpl_b200 import sim
    fields = [n for n, _ in sim._Ego._fields_]
    size, offs = _c_layout(tmp_path, os.path.join(inc, "tplb200_prep.h"), "tplb_ego", fields)
    assert size == ctypes.sizeof(sim._Ego)
    assert offs == [getattr(sim._Ego, n).offset for n in fields]
