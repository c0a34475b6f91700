"""Oracle: analytic Gaussian moments (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates example_problems/kinetic_fokker_planck_example_OU.py:15-106 (kinetic OU
configuration, Lyapunov ODE  m' = F m,  P' = F P + P F^T + L) and
example_problems/fokker_planck_example.py:20-55 (overdamped OU closed form).
The reference integrates the Lyapunov ODE with jax.experimental.ode.odeint
(dopri5, fp32); the oracle evaluates the same ODE's exact solution with a
float64 matrix exponential (Van Loan block form) — KAT-1 — and the exact
discrete-time moment recursion of the reference's integrator — KAT-2.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import scipy.linalg


def kinetic_ou_configuration(domain_dim: int, seed: int = 2217, gamma_friction: float = 1.0,
                             P_x0_scale: float = 1.0, P_v0_scale: float = 1.0) -> Dict:
    """OU.py:15-70.  `_F ~ N(0,1)[d, d+1]`, tilde_F = _F _F^T; the JAX threefry
    stream of PRNGKey(2217) is not reproduced — tilde_F is an input chosen by our seed."""
    rng = np.random.default_rng(seed)
    d = domain_dim
    _F = rng.standard_normal((d, d + 1))
    tilde_F = _F @ _F.T
    Z, I = np.zeros((d, d)), np.eye(d)
    F = np.block([[Z, I], [-tilde_F, -gamma_friction * I]])
    L = np.block([[Z, Z], [Z, 2.0 * I]])
    m_0 = np.zeros(2 * d)
    P_0 = np.block([[P_x0_scale * I, Z], [Z, P_v0_scale * I]])
    return {"gamma_friction": gamma_friction, "tilde_F": tilde_F, "F": F, "L": L, "m_0": m_0,
            "P_0": P_0, "m_x_0": m_0[:d], "P_x_0": P_0[:d, :d]}


def lyapunov_mean_cov(t: float, cfg: Dict) -> Tuple[np.ndarray, np.ndarray]:
    """Exact solution at time t of OU.py:78-84 (KAT-1), Van Loan:
    expm([[F, L],[0, -F^T]] t) = [[Phi, Q Phi^-T... ]]  ->  P(t) = Phi P0 Phi^T + G12 Phi^T."""
    F, L = cfg["F"], cfg["L"]
    n = F.shape[0]
    M = np.block([[F, L], [np.zeros((n, n)), -F.T]]) * t
    E = scipy.linalg.expm(M)
    Phi = E[:n, :n]
    G12 = E[:n, n:]
    m = Phi @ cfg["m_0"]
    P = Phi @ cfg["P_0"] @ Phi.T + G12 @ Phi.T
    return m, 0.5 * (P + P.T)


def discrete_step_matrices(h: float, tilde_F: np.ndarray, gamma: float):
    """One reference step z' = A_h z + B_h xi (from sampling_utils.py:17,20)."""
    d = tilde_F.shape[0]
    I = np.eye(d)
    A = np.block([[I - h * h * tilde_F, h * (1 - gamma * h) * I],
                  [-h * tilde_F, (1 - gamma * h) * I]])
    B = np.vstack([h * math.sqrt(2 * h) * I, math.sqrt(2 * h) * I])
    return A, B


def discrete_mean_cov(n_steps: int, dt: float, cfg: Dict, tau0: float = 0.0):
    """Exact ensemble moments of the reference scheme for linear drift (KAT-2):
    step(tau0), (n_steps-1) x step(dt), step(dt - tau0); m <- A m, P <- A P A^T + B B^T."""
    m, P = cfg["m_0"].copy(), cfg["P_0"].copy()
    hs = [tau0] + [dt] * (n_steps - 1) + [dt - tau0]
    for h in hs:
        A, B = discrete_step_matrices(h, cfg["tilde_F"], cfg["gamma_friction"])
        m = A @ m
        P = A @ P @ A.T + B @ B.T
    return m, P


def overdamped_ou_configuration(domain_dim: int, seed: int = 2217) -> Dict:
    """fokker_planck_example.py:20-46: F = _F _F^T, L = 2 I, m0 = 1, P0 = 5 I."""
    rng = np.random.default_rng(seed)
    d = domain_dim
    _F = rng.standard_normal((d, d + 1))
    F = _F @ _F.T
    L = np.eye(d) * 2.0
    m_0 = np.ones(d)
    P_0 = np.eye(d) * 5.0
    U, s, _ = np.linalg.svd(F)
    return {"F": F, "L": L, "U": U, "ss": s + s[:, None], "B": U.T @ L @ U,
            "B_0": U.T @ P_0 @ U, "s": s, "m_0": m_0, "P_0": P_0}


def overdamped_ou_mean_cov(t: float, cfg: Dict):
    """fokker_planck_example.py:48-55 (closed form through the SVD of F)."""
    exp_t_s = np.diag(np.exp(-t * cfg["s"]))
    m_t = cfg["U"] @ exp_t_s @ cfg["U"].T @ cfg["m_0"]
    P_t_1 = exp_t_s @ cfg["B_0"] @ exp_t_s
    B_S = cfg["B"] / cfg["ss"]
    P_t_2 = B_S - exp_t_s @ B_S @ exp_t_s
    P_t = cfg["U"] @ (P_t_1 + P_t_2) @ cfg["U"].T
    return m_t, P_t


def gaussian_cov_half(cov: np.ndarray) -> np.ndarray:
    """core/distribution.py:59-62: U sqrt(S) U^T from the SVD."""
    U, S, _ = np.linalg.svd(cov)
    return U @ np.diag(np.sqrt(S)) @ U.T
