"""Mirror of utils/sampling_utils.py:25-52 on the K1 CUDA integrator."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import _lib as L
from .. import ops


def drift_spec(potential_grad):
    """Map the `potential_grad` callable the reference passes (Potential.gradient, GMM.py:120) to a drift
    kind + parameter tensor of the C ABI."""
    owner = getattr(potential_grad, "__self__", None)
    if owner is None or not hasattr(owner, "drift_spec"):
        raise NotImplementedError(
            "potential_grad must be the .gradient method of a pde_inverse_problem_b200.core.potential.Potential; "
            "arbitrary Python callables cannot run inside the fused CUDA integrator")
    return owner.drift_spec()


def underdamped_langevin_dynamics_scan(q0_p0: torch.Tensor, n_steps: int, dt: float, key, potential_grad,
                                       gamma_friction: float, *, noise: Optional[torch.Tensor] = None,
                                       tau0: Optional[torch.Tensor] = None, particle_offset: int = 0,
                                       traj_layout: int = L.TRAJ_PARTICLE_MAJOR, want_trajectory: bool = True,
                                       ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (last_sample [N,2d], trajectory [N,n_steps,2d], tau_trajectory [N,n_steps]) exactly like the
    vmapped reference (sampling_utils.py:52).  `key` is an integer seed (the reference passes one jax key
    per particle; here the per-particle stream is Philox(seed, particle_offset + n)).  `noise`/`tau0` inject
    the random draws for parity runs."""
    kind, params, n_gaussian, sigma = drift_spec(potential_grad)
    z_last, traj, tau = ops.kl_integrate(
        q0_p0, n_steps, float(dt), float(gamma_friction), kind, params, n_gaussian=n_gaussian, sigma=sigma,
        noise=noise, tau0=tau0, seed=int(key), particle_offset=particle_offset,
        schedule=L.SCHEDULE_REFERENCE, traj_layout=traj_layout, want_traj=want_trajectory, want_tau=True)
    return z_last, traj, tau
