"""Mirror of core/distribution.py for the classes the live path uses: Gaussian (:52-84) and
Uniform (:162-186)."""
from __future__ import annotations

import warnings

import numpy as np
import torch

from .. import _lib as L
from .. import ops
from ..utils import rng as jrandom


class Distribution:
    def sample(self, batch_size: int, key):
        raise NotImplementedError


class Gaussian(Distribution):
    """core/distribution.py:52-65.  cov_half = U sqrt(S) U^T from the SVD (:59-62), computed once on the
    host in float64; sample = cov_half xi + mu with xi from Philox on the device (:64-65)."""

    def __init__(self, mu: torch.Tensor, cov: torch.Tensor):
        assert mu.ndim == 1 and cov.ndim == 2 and cov.shape[0] == cov.shape[1] and cov.shape[0] == mu.shape[0]
        self.dim = mu.shape[0]
        self.mu = mu.float().contiguous()
        self.cov = cov.float().contiguous()
        c = cov.detach().double().cpu().numpy()
        U, S, _ = np.linalg.svd(c)
        self.cov_half = torch.as_tensor(U @ np.diag(np.sqrt(S)) @ U.T, dtype=torch.float32,
                                        device=mu.device).contiguous()
        self.inv_cov = torch.as_tensor(np.linalg.inv(c), dtype=torch.float32, device=mu.device)
        self.log_det = float(np.log(np.linalg.det(c * 2 * np.pi)))

    def sample(self, batch_size: int, key, particle_offset: int = 0, layout: int = L.LAYOUT_AOS):
        return ops.gaussian_sample(batch_size, self.dim, self.mu, self.cov_half, int(key),
                                   particle_offset=particle_offset, layout=layout, device=self.mu.device)

    def score(self, x: torch.Tensor):
        return ops.linear_grad((self.mu - x).contiguous(), self.inv_cov)


class Uniform(Distribution):
    """core/distribution.py:162-186 (scalar interval used for distribution_time, api.py:35-37)."""

    def __init__(self, mins, maxs):
        self.mins = float(mins)
        self.maxs = float(maxs)

    def sample(self, batch_size: int, key, device="cuda"):
        u = ops.philox_uniforms(batch_size, int(key), device=device)
        return self.mins + (self.maxs - self.mins) * u
