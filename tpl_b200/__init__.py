"""tpl_b200 — batched iLQR/DDP trajectory optimisation for NVIDIA B200 (sm_100a).

A from-scratch implementation of the optimal-control solve behind tpl's spatial
trajectory planner and MPC controllers (reference: library/tpl/optim/), applied
to batches of independent problems.  See DESIGN.md.
"""

__version__ = "0.1.0"
