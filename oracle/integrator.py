"""Oracle: kinetic-Langevin integrator (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates utils/sampling_utils.py:6-52 with all randomness *injected* so the
CUDA kernel and the oracle consume identical noise:

  noise [N, S+1, d]   the S+1 standard-normal draws of one trajectory
                      (reference: one random.normal per update_step, :14)
  tau0  [N]           the per-particle initial time shift tau_0 = U(0,1)*dt (:32)
"""
from __future__ import annotations

import math
from typing import Callable, Tuple

import torch


def update_step(q, p, xi, h, potential_grad: Callable, gamma_friction: float):
    """One step, utils/sampling_utils.py:6-22.

    p' = p - h*grad_U(q) + sqrt(h)*sqrt(2)*xi - gamma*p*h      (:17)
    q' = q + h*p'                                              (:20, uses p')
    `h` may be a per-particle tensor [N,1] (tau_0 / dt - tau_0) or a scalar.
    """
    grad_U = potential_grad(q)
    noise = math.sqrt(2.0) * xi
    sqrt_h = torch.sqrt(h) if torch.is_tensor(h) else math.sqrt(h)
    p_new = p - h * grad_U + sqrt_h * noise - gamma_friction * p * h
    q_new = q + h * p_new
    return q_new, p_new


def underdamped_langevin_dynamics_scan(
    q0_p0: torch.Tensor,
    n_steps: int,
    dt: float,
    noise: torch.Tensor,
    tau0: torch.Tensor,
    potential_grad: Callable,
    gamma_friction: float,
    want_trajectory: bool = True,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """utils/sampling_utils.py:25-52, vmapped over particles (axis 0).

    Returns (last[N,2d], traj[N,S,2d], tau[N,S]) — particle-major, as the
    vmapped reference does (:25, :52).  S+1 update_steps per particle:
    step(tau_0) (:33), S-1 x step(dt) (:37-42), step(dt - tau_0) (:45-46).
    """
    N, two_d = q0_p0.shape
    d = two_d // 2
    assert noise.shape == (N, n_steps + 1, d), noise.shape
    assert tau0.shape == (N,)
    q, p = q0_p0[:, :d].clone(), q0_p0[:, d:].clone()
    h0 = tau0[:, None]
    q, p = update_step(q, p, noise[:, 0], h0, potential_grad, gamma_friction)
    traj = []
    if want_trajectory:
        traj.append(torch.cat([q, p], -1))
    for s in range(1, n_steps):
        q, p = update_step(q, p, noise[:, s], dt, potential_grad, gamma_friction)
        if want_trajectory:
            traj.append(torch.cat([q, p], -1))
    h_end = dt - h0
    q, p = update_step(q, p, noise[:, n_steps], h_end, potential_grad, gamma_friction)
    last = torch.cat([q, p], -1)
    tau = tau0[:, None] + torch.arange(n_steps, dtype=q0_p0.dtype)[None, :] * dt  # :48
    trajectory = torch.stack(traj, 1) if want_trajectory else q0_p0.new_zeros((N, 0, two_d))
    return last, trajectory, tau


def fixed_step_scan(z0, n_steps, dt, noise, potential_grad, gamma_friction):
    """n_steps plain steps of size dt (tau_0 = 0 limit of the scheme; KAT-2)."""
    d = z0.shape[1] // 2
    q, p = z0[:, :d].clone(), z0[:, d:].clone()
    for s in range(n_steps):
        q, p = update_step(q, p, noise[:, s], dt, potential_grad, gamma_friction)
    return torch.cat([q, p], -1)


def interacting_langevin_scan(z0, n_steps, dt, noise, A, gamma_friction):
    """The interacting particle system of README.md:54-62 with Phi(x) = x'Ax/2: drift grad U = A (x - xbar_t), xbar_t the
    EMPIRICAL mean of the whole ensemble, recomputed before every step.  The reference never integrates it (it samples
    the equivalent -A x law, README.md:64-71), so this is the definition the CUDA mean-field table is checked against.
    Common time grid: utils/sampling_utils.py:32-46 with tau_0 = 0, i.e. step 0 has h = 0 (sample 0 = z0), steps 1..S have
    h = dt.  Returns (last [N,2d], traj [N,S,2d], xbar [S+1,d] = mean before step s, drift [N,S,d] = grad U at the emitted
    samples)."""
    d = z0.shape[1] // 2
    q, p = z0[:, :d].clone(), z0[:, d:].clone()
    traj, xbars, drifts = [], [], []
    for s in range(n_steps + 1):
        h = 0.0 if s == 0 else dt
        xbar = q.mean(0)
        xbars.append(xbar)
        shift = xbar
        grad_fn = lambda x: (x - shift) @ A.T
        if s >= 1:
            drifts.append(grad_fn(q))
        q, p = update_step(q, p, noise[:, s], h, grad_fn, gamma_friction)
        if s < n_steps:
            traj.append(torch.cat([q, p], -1))
    return torch.cat([q, p], -1), torch.stack(traj, 1), torch.stack(xbars, 0), torch.stack(drifts, 1)


def meanfield_mean_recursion(sum_noise, sum_q0, sum_p0, n_global, dt, gamma_friction):
    """Closed recursion of the ensemble mean of interacting_langevin_scan: the interaction A (xbar - xbar) vanishes in the
    mean, so pbar' = (1 - gamma h) pbar + sqrt(2 h) xibar_s, qbar' = qbar + h pbar' with xibar_s = sum_noise[s] / N.
    sum_noise [S+1, d] (summed over ALL ranks' particles).  Returns xbar [S+1, d] (mean before step s)."""
    q, p = sum_q0 / n_global, sum_p0 / n_global
    out = []
    for s in range(sum_noise.shape[0]):
        out.append(q.clone())
        h = 0.0 if s == 0 else dt
        p = (1.0 - gamma_friction * h) * p + math.sqrt(2.0 * h) * sum_noise[s] / n_global
        q = q + h * p
    return torch.stack(out, 0)
