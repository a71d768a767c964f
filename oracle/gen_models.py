#!/usr/bin/env python
"""TEST INFRASTRUCTURE — writes oracle/models/<name>.h for the CPU oracle.

Each header holds the thirteen model routines of one problem under the
reference's routine names (genopt.py:105-110, 141-148, 168-172, 177-179),
printed as plain C99 over doubles with one common-subexpression pass per
routine, like the reference generator does (genopt.py:573-582).  The symbolic
expressions come from tpl_b200.derive, which tests/test_derive_vs_reference.py
proves identical to the reference pipeline's expressions.
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main(names=None):
    from tpl_b200 import codegen, derive, optimizers

    out_dir = os.path.join(HERE, "models")
    os.makedirs(out_dir, exist_ok=True)
    for name in names or list(optimizers.CONFIGS):
        cfg = optimizers.CONFIGS[name]()
        text = codegen.emit_c_model(derive.derive(cfg), name, cfg.definition_hash())
        path = os.path.join(out_dir, name + ".h")
        with open(path, "w") as fd:
            fd.write(text)
        print("wrote", os.path.relpath(path, HERE))


if __name__ == "__main__":
    main(sys.argv[1:])
