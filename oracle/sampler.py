"""Oracle: offline 0T sub-sampler and Gaussian sampler (TEST INFRASTRUCTURE).

Restates methods/consistency.py:90-118 (offline mode: every 5th time stamp with a random
shift, a random fifth of the trajectories, flattened) and core/distribution.py:52-65
(Gaussian.sample) with the random draws injected.
"""
from __future__ import annotations

import torch

from . import moments


def offline_subsample_0T(dataset_0T: torch.Tensor, shift: int, perm: torch.Tensor,
                         interval_time: int = 5, interval_sample: int = 5) -> torch.Tensor:
    """dataset_0T [n_traj, n_time, 2d] -> [n_traj//5 * n_time//5, 2d].

    consistency.py:102-105: time_index = arange(n_time // 5) * 5 + shift
    consistency.py:111-113: random_sample_index = perm[: n_traj // 5]
    consistency.py:115-118: index trajectories, then time, then flatten the first two dims.
    """
    n_traj, n_time, _ = dataset_0T.shape
    time_index = torch.arange(n_time // interval_time) * interval_time + shift
    sample_index = perm[: n_traj // interval_sample]
    data = dataset_0T[sample_index]
    data = data[:, time_index, :]
    return data.reshape(data.shape[0] * data.shape[1], data.shape[2])


def gaussian_sample(mu: torch.Tensor, cov: torch.Tensor, xi: torch.Tensor) -> torch.Tensor:
    """core/distribution.py:64-65: v_matmul(cov_half, xi) + mu, cov_half from the SVD (:59-62)."""
    cov_half = torch.as_tensor(moments.gaussian_cov_half(cov.double().numpy()), dtype=xi.dtype)
    return xi @ cov_half.T + mu
