"""nvcc driver: problem definition -> ``libtplb200_<name>.so`` for sm_100a.

Counterpart of the reference's ``build_module`` (genopt.py:464-619), which
wrote ``optim.c`` and ran cmake.  Here the generated part is only the model
header (``csrc/generated/<name>.cuh``); the solver kernels are the hand-written
``csrc/solver.cuh`` and the C ABI is ``csrc/cabi.cu``.

Everything is built in-tree so the shared objects travel with the repository
snapshot: the model zoo into ``tpl_b200/lib/``, user-defined problems into
``tpl_b200/lib/cache/<sha1>/`` (or ``$TPLB_CACHE_DIR``; the reference uses
``~/.cache/genopt/<sha1>``, genopt.py:439).

    python -m tpl_b200.build            # build every zoo model that is stale
    python -m tpl_b200.build --regen    # regenerate the model headers too
"""

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
GENERATED = os.path.join(CSRC, "generated")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--shared",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-fno-gnu-unique",
    "--expt-relaxed-constexpr",
]
if os.environ.get("TPLB_LIBDEVICE_MATH"):                  # A/B: CUDA libdevice sin/cos/tan/div/sqrt
    NVCC_FLAGS.append("-DTPLB_LIBDEVICE_MATH")


def default_lib_dir():
    return os.path.join(PKG, "lib")


def nvcc_path():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; the CUDA toolkit is required to build tpl_b200 solvers")
    return cand


def _solver_sources_hash():
    h = hashlib.sha1()
    for fn in ("solver.cuh", "solo.cuh", "device_math.cuh", "fast_math.cuh", "cabi.cu"):
        with open(os.path.join(CSRC, fn), "rb") as fd:
            h.update(fd.read())
    with open(os.path.join(PKG, "..", "include", "tplb200.h"), "rb") as fd:
        h.update(fd.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _generator_hash():
    """sha1 over the sources that decide the text of a generated model header."""
    h = hashlib.sha1()
    for fn in ("codegen.py", "derive.py", "symext.py"):
        with open(os.path.join(PKG, fn), "rb") as fd:
            h.update(fd.read())
    return h.hexdigest()


@dataclass
class Prepared:
    name: str
    header: str
    lib_path: str
    stamp: str


def prepare_model_sources(config, name=None, lib_dir=None, regen=False) -> Prepared:
    """Generate (or reuse) the model header; decide where the library goes."""
    from . import codegen, derive

    dh = config.definition_hash()
    zoo = name is not None
    if zoo:
        header = os.path.join(GENERATED, name + ".cuh")
        lib_dir = lib_dir or default_lib_dir()
    else:
        name = "genopt" + dh
        cache = os.environ.get("TPLB_CACHE_DIR") or os.path.join(default_lib_dir(), "cache")
        root = os.path.join(os.path.expanduser(cache), dh)
        header = os.path.join(root, name + ".cuh")
        lib_dir = lib_dir if (lib_dir and lib_dir != default_lib_dir()) else root
    os.makedirs(os.path.dirname(header), exist_ok=True)
    os.makedirs(lib_dir, exist_ok=True)

    # a header is reused only if both the problem definition and the code generator are unchanged
    gen = _generator_hash()
    fresh = False
    if os.path.exists(header) and not regen:
        with open(header) as fd:
            top = fd.read(400)
            fresh = ("definition sha1: " + dh) in top and ("generator sha1: " + gen) in top
    if not fresh:
        text = codegen.emit_cuda_model(derive.derive(config), name, dh)
        first, rest = text.split("\n", 1)
        with open(header, "w") as fd:
            fd.write(first + "\n// generator sha1: " + gen + "\n" + rest)
    return Prepared(name, header, os.path.join(lib_dir, f"libtplb200_{name}.so"),
                    dh + ":" + _solver_sources_hash())


def compile_prepared(p: Prepared, force=False, verbose=False) -> str:
    stamp_file = p.lib_path + ".stamp"
    if not force and os.path.exists(p.lib_path) and os.path.exists(stamp_file):
        with open(stamp_file) as fd:
            if fd.read().strip() == p.stamp:
                return p.lib_path
    cmd = [nvcc_path(), *NVCC_FLAGS, "-I", CSRC,
           f'-DTPLB_MODEL_HEADER="{p.header}"',
           os.path.join(CSRC, "cabi.cu"), "-o", p.lib_path]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (p.name, res.stdout))
    if verbose:
        print(res.stdout)
    with open(stamp_file, "w") as fd:
        fd.write(p.stamp)
    return p.lib_path


def build_model_library(config, name=None, lib_dir=None, force=False) -> str:
    return compile_prepared(prepare_model_sources(config, name=name, lib_dir=lib_dir), force=force)


def zoo_library_path(name):
    return os.path.join(default_lib_dir(), f"libtplb200_{name}.so")


def build_prep(force=False, verbose=False) -> str:
    """libtplb200_prep.so: the model-independent profile shaping (csrc/prep.cu, include/tplb200_prep.h)."""
    lib_dir = default_lib_dir()
    os.makedirs(lib_dir, exist_ok=True)
    lib_path = os.path.join(lib_dir, "libtplb200_prep.so")
    h = hashlib.sha1()
    for fn in (os.path.join(CSRC, "prep.cu"), os.path.join(PKG, "..", "include", "tplb200_prep.h")):
        with open(fn, "rb") as fd:
            h.update(fd.read())
    h.update((" ".join(NVCC_FLAGS) + " -fmad=false").encode())
    stamp, stamp_file = h.hexdigest(), lib_path + ".stamp"
    if not force and os.path.exists(lib_path) and os.path.exists(stamp_file):
        with open(stamp_file) as fd:
            if fd.read().strip() == stamp:
                return lib_path
    # -fmad=false: the preparation kernels keep the reference's operation order (decisions on
    # sample intervals, tolerances of the resampling march) and are nowhere near FP64 bound
    cmd = [nvcc_path(), *NVCC_FLAGS, "-fmad=false", os.path.join(CSRC, "prep.cu"), "-o", lib_path]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for prep.cu:\n%s" % res.stdout)
    if verbose:
        print(res.stdout)
    with open(stamp_file, "w") as fd:
        fd.write(stamp)
    return lib_path


def build_zoo(names=None, force=False, regen=False, verbose=False):
    """Build the shipped problem definitions in-tree; returns {name: path}."""
    from . import optimizers

    names = list(names or optimizers.CONFIGS)
    prepared = [prepare_model_sources(optimizers.CONFIGS[n](), name=n, regen=regen) for n in names]
    with ThreadPoolExecutor(max_workers=min(8, len(prepared))) as pool:
        paths = list(pool.map(lambda p: compile_prepared(p, force=force, verbose=verbose), prepared))
    return dict(zip(names, paths))


def main(argv=None):
    ap = argparse.ArgumentParser(description="build the tpl_b200 solver libraries for sm_100a")
    ap.add_argument("names", nargs="*")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--regen", action="store_true", help="regenerate csrc/generated/*.cuh")
    ap.add_argument("-v", "--verbose", action="store_true", help="show ptxas resource usage")
    a = ap.parse_args(argv)
    for n, p in build_zoo(a.names or None, force=a.force, regen=a.regen, verbose=a.verbose).items():
        print(f"{n}: {os.path.relpath(p)}")
    if not a.names:
        print(f"prep: {os.path.relpath(build_prep(force=a.force, verbose=a.verbose))}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
