"""CPU oracle for the PDE-inverse-problem hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (torch.func autodiff, float64 truth / float32
"reference-precision" twin) of the reference's JAX implementation of the
particle-ensemble loop.  Each function cites the reference file:line it
follows (paths relative to the reference checkout).

PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or
fixtures for this path (its only test asset, test_partial_s_log_density.py,
prints two RMSEs without asserting), and the reference itself cannot run in
this image (jax / flax / optax / hydra are not installed and not in the offline
wheelhouse).  The oracle is therefore pinned against analytic known-answer
tests derived from the reference semantics (tests/test_oracle_kat.py) and the
committed fixtures under tests/golden/ are *oracle* outputs (float64), produced
by tests/golden/make_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this package.  The product (pde_inverse_problem_b200) never does.
"""
