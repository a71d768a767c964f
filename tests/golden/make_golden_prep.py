#!/usr/bin/env python
"""Record golden vectors of the reference's two `rampify_profile` functions (SURVEY.md
section 8, row f2) by running the REAL reference code under numba:

    library/tpl/planning/utils.py:5-65                        (velocity profile)
    library/tpl/planning/path_vel_decomp/path_optim.py:11-55  (lateral corridor)

Neither module can be imported here (`tpl.planning/__init__` needs tplcpp, path_optim needs
objtoolbox), so the function definitions are taken from the reference files by name with `ast`
at run time and compiled as they stand (same decorator arguments except `cache`, which needs a
file-backed module).  Nothing of the reference is copied into the repository; only inputs and
outputs are stored: tests/golden/prep_velocity.npz, tests/golden/prep_lateral.npz.

    python tests/golden/make_golden_prep.py        # needs /root/reference and numba
"""
import ast
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(ROOT, "oracle", "_ref", "numba_cache"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from tpl_b200 import prep_scenarios as ps  # noqa: E402

REF = "/root/reference/library/tpl"


def reference_function(path, name):
    """Compile function `name` of the reference file `path` in a fresh namespace."""
    import numba
    src = open(path).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    seg = ast.get_source_segment(src, node)
    deco = "\n".join("@" + ast.get_source_segment(src, d).replace("cache=True, ", "").replace("cache=True", "")
                     for d in node.decorator_list)
    ns = {"numba": numba, "np": np}
    exec(compile(deco + "\n" + seg, path, "exec"), ns)
    return ns[name]


def reference_method(path, cls, name, extra):
    """Method `name` of class `cls` in the reference file `path`, as a plain function."""
    import textwrap
    src = open(path).read()
    tree = ast.parse(src)
    c = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
    node = next(n for n in c.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = dict(extra)
    exec(compile(textwrap.dedent(ast.get_source_segment(src, node, padded=True)), path, "exec"), ns)
    return ns[name]


def shift_interp_golden():
    """VelocityOptim.shift_interp (velocity_optim.py:98-104) run as it stands, with scipy's interp1d."""
    from types import SimpleNamespace
    from scipy.interpolate import interp1d
    fn = reference_method(os.path.join(REF, "planning", "path_vel_decomp", "velocity_optim.py"),
                          "VelocityOptim", "shift_interp", {"interp1d": interp1d, "np": np})
    out = {}
    for i, c in enumerate(ps.shift_cases()):
        n = c["arr"].shape[0]
        me = SimpleNamespace(ss=np.arange(0.0, n * c["step"], c["step"])[:n])      # update_shifts, :88
        me.shifts = me.ss + c["arc_len"]                                          # :92
        out[f"linear_{i}"] = fn(me, c["arr"])
        out[f"zero_{i}"] = fn(me, c["arr"], interp_kind="zero")
    np.savez_compressed(os.path.join(HERE, "prep_shift_interp.npz"), **out)
    print("shift_interp:", len(ps.shift_cases()), "cases")


def resample_path_golden():
    """util.interp_resampled_path (library/tpl/util.py:155-191, numba) run as it stands on the index
    tables the C restatement of `resample` produces (the C++ `resample` itself needs Eigen, which is
    not in this image): tests/golden/prep_path.npz."""
    from oracle import prep as oprep
    path = os.path.join(REF, "util.py")
    norm = reference_function(path, "normalize_angle")
    import numba
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"numba": numba, "np": np, "normalize_angle": norm}
    for name in ("short_angle_dist", "interp_resampled_path"):
        node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
        deco = "\n".join("@" + ast.get_source_segment(src, d).replace("cache=True, ", "").replace("cache=True", "")
                         for d in node.decorator_list)
        exec(compile(deco + "\n" + ast.get_source_segment(src, node), path, "exec"), ns)
    fn = ns["interp_resampled_path"]
    out = {}
    for i, (pth, step, steps, start, zero_end) in enumerate(ps.path_cases()):
        rsi = oprep.resample(pth[:, :2], step, steps, start, False)
        out[f"rsi_{i}"] = rsi
        out[f"rs_{i}"] = fn(np.array(pth), rsi, step, steps, zero_end, False)
    np.savez_compressed(os.path.join(HERE, "prep_path.npz"), **out)
    print("resample_path:", len(ps.path_cases()), "cases")


def update_ego_golden():
    """SimCore.update_ego (simulation/core.py:91-134) run as it stands: the method is taken from the
    reference file, `util.normalize_angle` from the reference's util.py (numba), `self` and `ego`
    are plain namespaces with the attributes the method touches."""
    from types import SimpleNamespace
    norm = reference_function(os.path.join(REF, "util.py"), "normalize_angle")
    fn = reference_method(os.path.join(REF, "simulation", "core.py"), "SimCore", "update_ego",
                          {"np": np, "util": SimpleNamespace(normalize_angle=norm)})
    out = {}
    for i, c in enumerate(ps.ego_cases()):
        me = SimpleNamespace(acc_buffer=[], steering_angle_buffer=[])
        ego = SimpleNamespace(**c["params"], **c["init"], control_acc=0.0, control_steer=0.0)
        rows = []
        for k in range(len(c["control_acc"])):
            ego.control_acc, ego.control_steer = float(c["control_acc"][k]), float(c["control_steer"][k])
            fn(me, ego, k * c["dt"], c["dt"])
            rows.append([ego.x, ego.y, ego.yaw, ego.v, ego.a, ego.steer_angle])
        out[f"states_{i}"] = np.array(rows)
    np.savez_compressed(os.path.join(HERE, "prep_update_ego.npz"), **out)
    print("update_ego:", len(ps.ego_cases()), "cases")


def main():
    if sys.argv[1:] == ["path"]:
        return resample_path_golden()
    resample_path_golden()
    shift_interp_golden()
    update_ego_golden()
    vel = reference_function(os.path.join(REF, "planning", "utils.py"), "rampify_profile")
    lat = reference_function(os.path.join(REF, "planning", "path_vel_decomp", "path_optim.py"), "rampify_profile")

    cases = ps.velocity_cases()
    out = {}
    for i, c in enumerate(cases):
        res = vel(c["v0"], c["a0"], c["lim_v"].copy(), c["a_min"], c["a_max"], c["j_min"], c["j_max"],
                  c["v_min"], c["step"])
        out[f"profile_{i}"] = np.asarray(res)
    np.savez_compressed(os.path.join(HERE, "prep_velocity.npz"), **out)
    print("velocity:", len(cases), "cases")

    cases = ps.lateral_cases()
    out = {}
    for i, c in enumerate(cases):
        res = lat(c["step"], c["horizon"], c["evasion_sharpness"], c["proj_distance"], c["path"], c["gap"],
                  c["lower"], c["upper"])
        out[f"d_offset_{i}"] = np.asarray(res)
    np.savez_compressed(os.path.join(HERE, "prep_lateral.npz"), **out)
    print("lateral:", len(cases), "cases")


if __name__ == "__main__":
    main()
