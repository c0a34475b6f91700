"""Oracle: potentials / drifts (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates core/potential.py:11-61 (GMM and quadratic potentials) with
torch.func autodiff, exactly as the reference obtains the gradient from
jax.grad (core/potential.py:37,46).
"""
from __future__ import annotations

import torch
from torch.func import grad, vmap


def gmm_V(x: torch.Tensor, mus: torch.Tensor, sigma) -> torch.Tensor:
    """core/potential.py:32-35: -logsumexp(-|x-mu_k|^2 / (2 sigma^2))."""
    a = -torch.sum((x - mus) ** 2, dim=1) / (2 * sigma ** 2)
    return -torch.logsumexp(a, dim=0)


g_gmm_V = grad(gmm_V)  # core/potential.py:37
vg_gmm_V = vmap(g_gmm_V, in_dims=(0, None, None))  # core/potential.py:46


class GMMPotential:
    """core/potential.py:48-61."""

    def __init__(self, mus: torch.Tensor, sigma):
        self.mus = mus
        self.sigma = sigma

    def value(self, x):
        if x.ndim == 1:
            return gmm_V(x, self.mus, self.sigma)
        return vmap(gmm_V, in_dims=(0, None, None))(x, self.mus, self.sigma)

    def gradient(self, x):
        if x.ndim == 1:
            return g_gmm_V(x, self.mus, self.sigma)
        return vg_gmm_V(x, self.mus, self.sigma)


def gmm_gradient_closed_form(x, mus, sigma=1.0):
    """Closed form of the same gradient: (x - sum_k softmax(a)_k mu_k)/sigma^2.

    Used only to cross-check the autodiff restatement (KAT-3)."""
    a = -((x[:, None, :] - mus[None]) ** 2).sum(-1) / (2 * sigma ** 2)
    w = torch.softmax(a, dim=1)
    return (x - w @ mus) / sigma ** 2


class QuadraticPotential:
    """core/potential.py:11-24: grad = inv(cov) @ (x - mu)."""

    def __init__(self, mu, cov):
        self.mu = mu
        self.cov = cov
        self.inv_cov = torch.linalg.inv(cov)

    def gradient(self, x):
        return (x - self.mu) @ self.inv_cov.T


class LinearDrift:
    """grad U = F~ x (kinetic OU, example_problems/kinetic_fokker_planck_example_OU.py:15-20,
    README.md:64-71 for the -A X McKean-Vlasov equivalence); optional mean-field
    centring A (x - xbar)."""

    def __init__(self, F_tilde, mean_field: bool = False):
        self.F = F_tilde
        self.mean_field = mean_field

    def gradient(self, x):
        if self.mean_field:
            x = x - x.mean(0, keepdim=True)
        return x @ self.F.T
