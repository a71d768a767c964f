#!/usr/bin/env python
"""Records tests/golden/ref_extra.npz from the REAL reference solvers (oracle/_ref): the cases of
tests/extra.py (ilr, shift, dynamics / ct_dynamics incl. the negative-index wrap, sticky state,
ref_line_smoother_dk, velocity_profile_time, a user-defined RK4 + augmented-Lagrangian problem,
prev_x / prev_k, a user-defined problem with lerp_wrap and blerp).  Run in the build container after `python oracle/build_ref.py`:

    python tests/golden/make_golden_extra.py
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import build_ref, ref            # noqa: E402
from tests import extra                      # noqa: E402
from tpl_b200 import _cabi, build            # noqa: E402


def main():
    genopt, symext, _ = build_ref.reference_modules()
    # the user-defined problem through the reference's own generator
    build_ref.build(extra.CUSTOM, cfg=extra.custom_definition(genopt, symext))
    build_ref.build(extra.TRACK, cfg=extra.track_definition(genopt, symext))

    def make(model):
        flavour = "strict" if model == "trajectory_tracking_mpc" else "fast"
        Ref = ref.load(model, flavour)
        if Ref is None:
            raise SystemExit(f"oracle/_ref/{model} is not built; run python oracle/build_ref.py first")
        return Ref()

    libs = build.build_zoo([n for n, _ in extra.ZOO])
    zoo_info = {n: _cabi.model_info(_cabi.load(p)) for n, p in libs.items()}
    out = extra.run_single(make, zoo_info)
    path = os.path.join(ROOT, "tests", "golden", "ref_extra.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} arrays -> {path} ({os.path.getsize(path) // 1024} KiB)")


if __name__ == "__main__":
    main()
