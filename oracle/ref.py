"""TEST INFRASTRUCTURE — loads the real reference solvers built by build_ref.py.

``load(name, flavour)`` returns the reference's own ``Optim`` class
(optim.c:1931-1944) from oracle/_ref/<name>/<flavour>/genopt<sha1>.so, or None
when oracle/_ref has not been built (then only the committed golden vectors and
the C restatement are available).
"""

import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "_ref")
_LOADED = {}


def available(name="trajectory_tracking_mpc_time"):
    return os.path.exists(os.path.join(ROOT, name, "meta.json"))


def load(name, flavour="fast"):
    key = (name, flavour)
    if key in _LOADED:
        return _LOADED[key]
    if not available(name):
        return None
    if flavour == "fast":
        from . import build_ref
        build_ref.ensure_native(name)       # -march=native objects are host specific
    with open(os.path.join(ROOT, name, "meta.json")) as fd:
        meta = json.load(fd)
    mod_name = "genopt" + meta["code_hash"]
    path = os.path.join(ROOT, name, flavour, mod_name + ".so")
    spec = importlib.util.spec_from_file_location(mod_name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _LOADED[key] = mod.Optim
    return mod.Optim
