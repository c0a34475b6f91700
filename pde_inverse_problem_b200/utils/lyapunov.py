"""Host-side (float64, once per problem / time stamp) Gaussian moments of the linear SDEs.

Replaces jax.experimental.ode.odeint on  m' = F m,  P' = F P + P F^T + L
(example_problems/kinetic_fokker_planck_example_OU.py:73-93) by the exact solution
    P(t) = Pinf + e^{Ft} (P0 - Pinf) e^{F^T t},   F Pinf + Pinf F^T + L = 0,   m(t) = e^{Ft} m0,
valid because F = [[0, I], [-tilde_F, -gamma I]] is Hurwitz for SPD tilde_F and gamma > 0.
This is problem set-up, not the particle hot path.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import scipy.linalg


def kinetic_ou_mean_cov(t: float, cfg: Dict) -> Tuple[np.ndarray, np.ndarray]:
    F, Lm, m0, P0 = cfg["F"], cfg["L"], cfg["m_0"], cfg["P_0"]
    Pinf = scipy.linalg.solve_continuous_lyapunov(F, -Lm)
    E = scipy.linalg.expm(F * float(t))
    P = Pinf + E @ (P0 - Pinf) @ E.T
    return E @ m0, 0.5 * (P + P.T)


def cov_half(cov: np.ndarray) -> np.ndarray:
    """core/distribution.py:59-62."""
    U, S, _ = np.linalg.svd(cov)
    return U @ np.diag(np.sqrt(S)) @ U.T
