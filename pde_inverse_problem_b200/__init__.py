"""pde_inverse_problem_b200 — B200-native hot path of shenzebang/PDE-inverse-problem.

Hand-written sm_100a CUDA kernels behind a C ABI (include/pdeip.h, libpdeip.so) with a thin Python
host layer that mirrors the reference's api.py / registry.py / methods/consistency.py /
consistency_instances interfaces on torch CUDA tensors.  No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
