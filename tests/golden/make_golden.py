#!/usr/bin/env python
"""Records tests/golden/*.npz from the REAL reference solvers (oracle/_ref).

Run in the build container after `python oracle/build_ref.py`:

    python tests/golden/make_golden.py            # all cases of tests/common.py
    python tests/golden/make_golden.py lateral_al # one case

For every problem of every case it stores the per-iteration trace (x, u, cost,
alpha, mu_step, iterations, termination_condition, improved, trajectory_changed
after max_iterations = 0..I) and the derivative blocks / gains of the first
iteration, exactly as the reference's `Optim` object reports them.
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from oracle import ref                      # noqa: E402
from tests import common                    # noqa: E402


def main(names):
    for name in names or list(common.CASES):
        pb, iters, flavour = common.make_case(name)
        Ref = ref.load(pb.model, flavour)
        if Ref is None:
            raise SystemExit("oracle/_ref is not built; run python oracle/build_ref.py first")
        traces = [common.trace_single(Ref, pb, i, iters) for i in range(pb.batch)]
        derivs = [common.derivatives_single(Ref, pb, i) for i in range(pb.batch)]
        common.save_golden(name, traces, derivs)
        last = traces[0][-1]
        print(f"{name:22s} {pb.model:30s} B={pb.batch} T={pb.horizon} it={int(last['iterations'])} "
              f"term={int(last['termination_condition'])} cost={last['traj_costs']:.12g} "
              f"({os.path.getsize(common.golden_path(name)) // 1024} KiB)")


if __name__ == "__main__":
    main(sys.argv[1:])
