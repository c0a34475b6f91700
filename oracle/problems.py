"""Oracle: problem instances (TEST INFRASTRUCTURE, see oracle/__init__.py).

Minimal stand-ins for what the residual modules read from `pde_instance`
(example_problems/*.py): V_true_fn / Phi_true_fn, gamma_friction,
total_evolving_time, and the KMV density derivatives
(example_problems/kinetic_mckean_vlasov_example_quadratic.py:18-191).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import moments
from .potential import gmm_V


class KineticOUProblem:
    """example_problems/kinetic_fokker_planck_example_OU.py:109-139."""

    def __init__(self, d: int, T: float = 2.0, seed: int = 2217, dtype=torch.float64):
        self.dim = d
        self.dtype = dtype
        self.np_cfg = moments.kinetic_ou_configuration(d, seed)
        self.initial_configuration = {
            k: (torch.as_tensor(v, dtype=dtype) if isinstance(v, np.ndarray) else v)
            for k, v in self.np_cfg.items()}
        self.total_evolving_time = T

    def V_true_fn(self, x):
        """OU.py:130-131: x . tilde_F x / 2."""
        return torch.dot(x, self.initial_configuration["tilde_F"] @ x) / 2

    Phi_true_fn = V_true_fn  # kinetic_mckean_vlasov_example_quadratic.py:193-196

    def get_mean_cov(self, t: float):
        m, P = moments.lyapunov_mean_cov(float(t), self.np_cfg)
        return torch.as_tensor(m, dtype=self.dtype), torch.as_tensor(P, dtype=self.dtype)

    # --- KMV density derivatives, kinetic_mckean_vlasov_example_quadratic.py:51-69 ---
    def partial_s_log_density_fn(self, s, x):
        mean, cov = self.get_mean_cov(float(s))
        F, L = self.initial_configuration["F"], self.initial_configuration["L"]
        d = self.dim
        mean1, cov11 = mean[:d], cov[:d, :d]
        cov11_inv = torch.linalg.inv(cov11)
        dmds = F @ mean
        dm_1ds = dmds[:d]
        dPds = F @ cov + cov @ F.T + L
        dP_11ds = dPds[:d, :d]
        dP_11ds_inv = -cov11_inv @ dP_11ds @ cov11_inv

        def single(xx):
            term1 = -dm_1ds @ cov11_inv @ (mean1 - xx)
            term2 = -0.5 * torch.trace(dP_11ds @ cov11_inv)
            term3 = -0.5 * torch.dot(mean1 - xx, dP_11ds_inv @ (mean1 - xx))
            return term1 + term2 + term3

        if x.ndim == 1:
            return single(x)
        return torch.stack([single(xx) for xx in x])

    # --- kinetic_mckean_vlasov_example_quadratic.py:120-177 ---
    def partial_s2_log_density_fn(self, s, x):
        mean, cov = self.get_mean_cov(float(s))
        F, L = self.initial_configuration["F"], self.initial_configuration["L"]
        d = self.dim
        mean1, cov11 = mean[:d], cov[:d, :d]
        cov11_inv = torch.linalg.inv(cov11)
        dmds = F @ mean
        dm_1ds = dmds[:d]
        d2mds2 = F @ dmds
        d2m_1ds2 = d2mds2[:d]
        dPds = F @ cov + cov @ F.T + L
        dP_11ds = dPds[:d, :d]
        d2Pds2 = F @ dPds + dPds @ F.T
        d2P_11ds2 = d2Pds2[:d, :d]
        dinv_P_11ds = -cov11_inv @ dP_11ds @ cov11_inv
        dinv_P_11_2ds2 = (-cov11_inv @ d2P_11ds2 @ cov11_inv
                          + cov11_inv @ dP_11ds @ cov11_inv @ dP_11ds @ cov11_inv * 2)
        term3 = 0.5 * torch.trace(cov11_inv @ dP_11ds @ cov11_inv @ dP_11ds) \
            - 0.5 * torch.trace(cov11_inv @ d2P_11ds2)

        def single(xx):
            term1 = (-d2m_1ds2 @ cov11_inv @ (mean1 - xx)
                     - dm_1ds @ dinv_P_11ds @ (mean1 - xx)
                     - dm_1ds @ cov11_inv @ dm_1ds)
            term2 = (-0.5 * (xx - mean1) @ dinv_P_11_2ds2 @ (xx - mean1)
                     - (mean1 - xx) @ dinv_P_11ds @ dm_1ds)
            return term1 + term2 + term3

        if x.ndim == 1:
            return single(x)
        return torch.stack([single(xx) for xx in x])

    def log_density_x(self, s: float, x):
        """log rho_s(x) of the x-marginal (for the KAT-7 finite-difference check)."""
        mean, cov = self.get_mean_cov(float(s))
        d = self.dim
        mean1, cov11 = mean[:d], cov[:d, :d]
        diff = x - mean1
        quad = diff @ torch.linalg.inv(cov11) @ diff
        return -0.5 * (torch.logdet(cov11 * 2 * torch.pi) + quad)


class KineticGMMProblem:
    """example_problems/kinetic_fokker_planck_example_GMM.py:16-102 (n_Gaussian is a
    parameter here; the reference hard-codes 3 at :19)."""

    def __init__(self, d: int, n_gaussian: int = 3, T: float = 2.0, seed: int = 1,
                 dtype=torch.float64):
        g = torch.Generator().manual_seed(seed)
        self.dim = d
        self.mus = ((torch.rand(n_gaussian, d, generator=g, dtype=torch.float64) * 8.0) - 4.0).to(dtype)
        self.sigma = 1.0
        self.initial_configuration = {
            "n_Gaussian": n_gaussian, "gamma_friction": 0.5,
            "m_0": torch.zeros(2 * d, dtype=dtype),
            "P_0": torch.diag(torch.cat([torch.full((d,), 4.0), torch.full((d,), 0.1)])).to(dtype),
        }
        self.total_evolving_time = T

    def V_true_fn(self, x):
        return gmm_V(x, self.mus, self.sigma)


class OverdampedOUProblem:
    """example_problems/fokker_planck_example.py:63-82."""

    def __init__(self, d: int, T: float = 5.0, seed: int = 2217, dtype=torch.float64):
        self.dim = d
        self.np_cfg = moments.overdamped_ou_configuration(d, seed)
        self.initial_configuration = {"F": torch.as_tensor(self.np_cfg["F"], dtype=dtype)}
        self.total_evolving_time = T

    def V_true_fn(self, x):
        return torch.dot(x, self.initial_configuration["F"] @ x) / 2
