"""Mirror of example_problems/kinetic_mckean_vlasov_example_quadratic.py: the quadratic-interaction
McKean-Vlasov problem reuses the kinetic OU law (:14-17) and adds the time derivatives of the log density of
the x-marginal (:18-191).  Those are per-time-stamp d x d linear algebra (host, float64, computed ONCE per time
stamp and cached: the time stamps come from a fixed grid) followed by two quadratic forms per sample, which run on
the device (pdeip_kmv_density_terms): no device-to-host copy and no host arithmetic per iteration."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
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

    def density_coefficient_row(self, s: float) -> np.ndarray:
        """[mean1 | a1 | a2 | k1 | k2 | M1 | M2] of pdeip_kmv_density_terms at time s (float64, cached per time stamp):
        d_s log rho = -a1.diff + k1 - diff'M1 diff/2 (:51-69), d_ss log rho = -a2.diff + k2 - diff'M2 diff/2
        (:120-177), diff = mean1 - x."""
        key = round(float(s), 12)
        cache = self.__dict__.setdefault("_coef_rows", {})
        row = cache.get(key)
        if row is None:
            c = self._density_coefficients(float(s))
            a1 = c["inv"].T @ c["dm1"]
            k1 = -0.5 * np.trace(c["dP11"] @ c["inv"])
            a2 = c["inv"].T @ c["d2m1"] + c["dinv"].T @ c["dm1"] + c["dinv"] @ c["dm1"]
            k2 = (-c["dm1"] @ c["inv"] @ c["dm1"] + 0.5 * np.trace(c["inv"] @ c["dP11"] @ c["inv"] @ c["dP11"])
                  - 0.5 * np.trace(c["inv"] @ c["d2P11"]))
            row = np.concatenate([c["mean1"], a1, a2, [k1, k2], c["dinv"].reshape(-1), c["d2inv"].reshape(-1)])
            cache[key] = row
        return row

    def density_coefficients(self, tau, device) -> torch.Tensor:
        """coef [nt, 3d + 2 + 2d^2] for the time stamps `tau` (a host sequence of floats, or a tensor read ONCE)."""
        taus = tau.tolist() if isinstance(tau, torch.Tensor) else [float(t) for t in tau]
        rows = np.stack([self.density_coefficient_row(t) for t in taus], 0)
        return torch.as_tensor(rows, dtype=torch.float32, device=device)

    def density_terms(self, tau, xv: torch.Tensor, gamma: float) -> torch.Tensor:
        """c[nt, n] = d_ss log rho + (d_s log rho)^2 + gamma d_s log rho at (tau_t, x[:, t]) on the device."""
        return ops.kmv_density_terms(xv, self.density_coefficients(tau, xv.device), gamma)

    def _density_parts(self, s, x: torch.Tensor):
        single = x.ndim == 1
        x2 = x[None] if single else x
        n, d = x2.shape
        xv = torch.zeros((n, 1, 2 * d), device=x2.device, dtype=torch.float32)
        xv[:, 0, :d] = x2
        _, ps, ps2 = ops.kmv_density_terms(xv, self.density_coefficients([float(s)], x2.device), 0.0, want_parts=True)
        return (ps[0, 0], ps2[0, 0]) if single else (ps[0], ps2[0])

    def partial_s_log_density_fn(self, s, x: torch.Tensor):
        """quadratic_example.py:18-83 for scalar s and x [d] or [n,d]."""
        return self._density_parts(s, x)[0]

    def partial_s2_log_density_fn(self, s, x: torch.Tensor):
        """quadratic_example.py:85-191 for scalar s and x [d] or [n,d]."""
        return self._density_parts(s, x)[1]

    def Phi_true_fn(self, x: torch.Tensor):
        """quadratic_example.py:193-203."""
        return self.V_true_fn(x)

    def create_parametric_model(self):
        return V_parametric_quadratic(self.dim)
