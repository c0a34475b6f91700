"""Mirror of example_problems/kinetic_mckean_vlasov_example_quadratic.py: the quadratic-interaction
McKean-Vlasov problem reuses the kinetic OU law (:14-17) and adds the time derivatives of the log density of
the x-marginal (:18-191).  Those are per-time-stamp d x d linear algebra (host, float64) followed by one
quadratic form per sample."""
from __future__ import annotations

import numpy as np
import torch

from ..core.model import V_parametric_quadratic
from ..utils import lyapunov
from .kinetic_fokker_planck_example_OU import KineticFokkerPlanck


class KineticMcKeanVlasov(KineticFokkerPlanck):
    def _density_coefficients(self, s: float):
        """Coefficients of d_s log rho = a0 + a1.(x) + x^T A2 x style forms (quadratic_example.py:51-69) and of
        d_s^2 log rho (:120-177), evaluated in float64 at time s."""
        cfg = self.np_configuration
        d = self.dim
        mean, cov = lyapunov.kinetic_ou_mean_cov(float(s), cfg)
        F, Lm = cfg["F"], cfg["L"]
        mean1, cov11 = mean[:d], cov[:d, :d]
        inv = np.linalg.inv(cov11)
        dmds = F @ mean
        dm1 = dmds[:d]
        d2m1 = (F @ dmds)[:d]
        dP = F @ cov + cov @ F.T + Lm
        dP11 = dP[:d, :d]
        d2P = F @ dP + dP @ F.T
        d2P11 = d2P[:d, :d]
        dinv = -inv @ dP11 @ inv
        d2inv = -inv @ d2P11 @ inv + inv @ dP11 @ inv @ dP11 @ inv * 2
        return dict(mean1=mean1, inv=inv, dm1=dm1, d2m1=d2m1, dP11=dP11, d2P11=d2P11, dinv=dinv, d2inv=d2inv)

    def partial_s_log_density_fn(self, s, x: torch.Tensor):
        """quadratic_example.py:18-83 for scalar s and x [d] or [n,d]."""
        c = self._density_coefficients(float(s))
        xs = x.detach().double().cpu().numpy()
        single = xs.ndim == 1
        xs = np.atleast_2d(xs)
        diff = c["mean1"][None] - xs
        term1 = -(diff @ (c["inv"].T @ c["dm1"]))
        term2 = -0.5 * np.trace(c["dP11"] @ c["inv"])
        term3 = -0.5 * np.einsum("ni,ij,nj->n", diff, c["dinv"], diff)
        out = torch.as_tensor(term1 + term2 + term3, dtype=torch.float32, device=x.device)
        return out[0] if single else out

    def partial_s2_log_density_fn(self, s, x: torch.Tensor):
        """quadratic_example.py:85-191 for scalar s and x [d] or [n,d]."""
        c = self._density_coefficients(float(s))
        xs = x.detach().double().cpu().numpy()
        single = xs.ndim == 1
        xs = np.atleast_2d(xs)
        diff = c["mean1"][None] - xs
        term1 = (-(diff @ (c["inv"].T @ c["d2m1"])) - (diff @ (c["dinv"].T @ c["dm1"]))
                 - c["dm1"] @ c["inv"] @ c["dm1"])
        term2 = (-0.5 * np.einsum("ni,ij,nj->n", -diff, c["d2inv"], -diff) - diff @ (c["dinv"] @ c["dm1"]))
        term3 = (0.5 * np.trace(c["inv"] @ c["dP11"] @ c["inv"] @ c["dP11"]) - 0.5 * np.trace(c["inv"] @ c["d2P11"]))
        out = torch.as_tensor(term1 + term2 + term3, dtype=torch.float32, device=x.device)
        return out[0] if single else out

    def Phi_true_fn(self, x: torch.Tensor):
        """quadratic_example.py:193-203."""
        return self.V_true_fn(x)

    def create_parametric_model(self):
        return V_parametric_quadratic(self.dim)
