#!/usr/bin/env python
"""TEST INFRASTRUCTURE — builds the *real* reference solvers into oracle/_ref/.

Recipe (runs only where /root/reference is mounted, i.e. in the build container):

1. import the reference's own, unmodified code generator
   (/root/reference/library/tpl/optim/{genopt,optimizers,symext}.py) and let
   ``genopt.build_module`` write ``optim.c`` for each problem
   (genopt.py:464-598).  The reference's own build system (cmake,
   genopt.py:600-610) is NOT run: the ``subprocess.Popen`` call is replaced by
   a stub for the duration of the call.
2. compile that single generated file directly with ``/usr/bin/gcc`` in two
   flavours:
     fast   -UNDEBUG -O3 -ffast-math -march=native   (the flags the reference
            ships, templates/CMakeLists.txt:16-20) — the contractual oracle and
            the CPU baseline;
     strict -O2 -fno-fast-math -ffp-contract=off      (bit-reproducible; used
            for the ill-conditioned 7x2 model and for debugging).
   ``-DPyArray_MoveInto=PyArray_CopyInto`` adapts optim.c:1343 to numpy 2.x.

Outputs go to oracle/_ref/<name>/ only (git-ignored, travels to the GPU box):
``optim.c`` (generated), ``fast/genopt<sha1>.so``, ``strict/genopt<sha1>.so``,
``meta.json``.  ``ensure_native(name)`` rebuilds the fast flavour from the
already generated ``optim.c`` when the host CPU differs from the build host
(``-march=native`` objects are not portable); that needs no reference tree.

Nothing in the product imports this file.
"""

import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference/library"
OUT = os.path.join(HERE, "_ref")
CC = "/usr/bin/gcc"

MODELS = (
    "trajectory_tracking_mpc", "trajectory_tracking_mpc_time", "lateral_profile",
    "velocity_profile_space", "ref_line_smoother_k", "ref_line_smoother_dk",
    "velocity_profile_time",
)

FLAVOURS = {
    "fast": ["-UNDEBUG", "-O3", "-ffast-math", "-march=native"],
    "strict": ["-UNDEBUG", "-O2", "-fno-fast-math", "-ffp-contract=off"],
}


def cpu_signature():
    """Identifies the ISA the -march=native objects were built for."""
    flags = ""
    try:
        with open("/proc/cpuinfo") as fd:
            for line in fd:
                if line.startswith("flags"):
                    flags = " ".join(sorted(line.split(":", 1)[1].split()))
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:16]


def _compile(src, out_so, flags):
    import numpy as np
    os.makedirs(os.path.dirname(out_so), exist_ok=True)
    cmd = [CC, "-shared", "-fPIC", "-w", *flags,
           "-DPyArray_MoveInto=PyArray_CopyInto",
           "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(),
           src, "-o", out_so, "-lm"]
    subprocess.run(cmd, check=True)


def reference_modules():
    """The reference's own (genopt, symext, optimizers), unmodified."""
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from tpl.optim import genopt, optimizers, symext
    return genopt, symext, optimizers


def generate(name, cfg=None):
    """Step 1: the reference's generator writes oracle/_ref/<name>/optim.c (`cfg`: a problem
    definition made with the reference's own genopt.Config / symext; default: the zoo config)."""
    genopt, _, optimizers = reference_modules()

    if cfg is None:
        cfg = getattr(optimizers, "config_" + name)()
    scratch = os.path.join(OUT, "_gen_" + name)
    shutil.rmtree(scratch, ignore_errors=True)
    cfg.output_dir = scratch
    cfg.use_cache = False

    class _NoBuild:                      # stands in for `cmake; make` (genopt.py:600-610)
        returncode = 0

        def __init__(self, *a, **k):
            pass

        def communicate(self):
            return b"", b""

    real_popen = genopt.subprocess.Popen
    genopt.subprocess.Popen = _NoBuild
    try:
        _, code_hash = genopt.build_module(cfg)
    finally:
        genopt.subprocess.Popen = real_popen

    dst = os.path.join(OUT, name)
    os.makedirs(dst, exist_ok=True)
    shutil.copyfile(os.path.join(scratch, code_hash, "optim.c"), os.path.join(dst, "optim.c"))
    shutil.rmtree(scratch, ignore_errors=True)
    return code_hash


def build(name, flavours=("fast", "strict"), cfg=None):
    code_hash = generate(name, cfg)
    dst = os.path.join(OUT, name)
    for fl in flavours:
        shutil.rmtree(os.path.join(dst, fl), ignore_errors=True)
        _compile(os.path.join(dst, "optim.c"),
                 os.path.join(dst, fl, f"genopt{code_hash}.so"), FLAVOURS[fl])
    meta = {"name": name, "code_hash": code_hash, "flags": FLAVOURS,
            "cc": CC, "cpu_signature": cpu_signature()}
    with open(os.path.join(dst, "meta.json"), "w") as fd:
        json.dump(meta, fd, indent=1)
    return meta


def ensure_native(name):
    """Rebuild the -march=native flavour if this host is not the build host.
    Uses only oracle/_ref/<name>/optim.c; returns False if that is missing."""
    dst = os.path.join(OUT, name)
    meta_path = os.path.join(dst, "meta.json")
    if not (os.path.exists(meta_path) and os.path.exists(os.path.join(dst, "optim.c"))):
        return False
    with open(meta_path) as fd:
        meta = json.load(fd)
    so = os.path.join(dst, "fast", f"genopt{meta['code_hash']}.so")
    if meta.get("cpu_signature") != cpu_signature() or not os.path.exists(so):
        _compile(os.path.join(dst, "optim.c"), so, FLAVOURS["fast"])
        meta["cpu_signature"] = cpu_signature()
        with open(meta_path, "w") as fd:
            json.dump(meta, fd, indent=1)
    return True


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("models", nargs="*", default=list(MODELS))
    args = ap.parse_args()
    if not os.path.isdir(REF_ROOT):
        print("reference tree not mounted; nothing to do")
        return 0
    os.makedirs(OUT, exist_ok=True)
    for name in args.models:
        meta = build(name)
        print(f"built {name}: {meta['code_hash']}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
