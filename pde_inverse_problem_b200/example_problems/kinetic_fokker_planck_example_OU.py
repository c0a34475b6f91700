"""Mirror of example_problems/kinetic_fokker_planck_example_OU.py (kinetic OU, exact Gaussian law)."""
from __future__ import annotations

import warnings
from math import prod

import numpy as np
import torch

from .. import _lib as L
from .. import ops
from ..api import ProblemInstance
from ..core.distribution import Gaussian
from ..core.model import V_parametric_quadratic
from ..core.potential import LinearPotential
from ..utils import lyapunov
from ..utils import rng as jrandom


def initialize_configuration(domain_dim: int, seed: int = 2217):
    """OU.py:15-70 (float64 numpy on the host; _F ~ N(0,1)[d,d+1] from our generator, not jax's
    PRNGKey(2217) stream)."""
    g = np.random.default_rng(seed)
    d = domain_dim
    _F = g.standard_normal((d, d + 1))
    tilde_F = _F @ _F.T
    gamma_friction = 1.0
    Z, I = np.zeros((d, d)), np.eye(d)
    m_0 = np.zeros(2 * d)
    P_0 = np.block([[I, Z], [Z, I]])
    F = np.block([[Z, I], [-tilde_F, -gamma_friction * I]])
    Lm = np.block([[Z, Z], [Z, 2.0 * I]])
    return {"gamma_friction": gamma_friction, "tilde_F": tilde_F, "F": F, "L": Lm, "m_0": m_0, "P_0": P_0,
            "m_x_0": m_0[:d], "P_x_0": P_0[:d, :d]}


def OU_process(t_space, configuration):
    """OU.py:73-93 (exact solution instead of odeint)."""
    t_space = np.atleast_1d(np.asarray(t_space, dtype=np.float64))
    assert t_space.size >= 2
    ms, Ps = zip(*[lyapunov.kinetic_ou_mean_cov(t, configuration) for t in t_space[1:]])
    if t_space.size == 2:
        return ms[0], Ps[0]
    return np.stack(ms), np.stack(Ps)


def get_mean_cov(t, configuration):
    """OU.py:96-106."""
    t = np.asarray(t, dtype=np.float64)
    if t.size == 1:
        return OU_process(np.array([0.0, float(t)]), configuration)
    assert t.ndim == 1
    warnings.warn("The user is responsible for ensuring t[0] == 0")
    return OU_process(t, configuration)


class KineticFokkerPlanck(ProblemInstance):
    """OU.py:109-220."""

    def __init__(self, cfg, rng, device="cuda"):
        super().__init__(cfg, rng, device)
        self.np_configuration = initialize_configuration(cfg.pde_instance.domain_dim)
        self.initial_configuration = {
            k: (torch.as_tensor(v, dtype=torch.float32, device=self.device).contiguous()
                if isinstance(v, np.ndarray) else v) for k, v in self.np_configuration.items()}
        self.get_mean_cov = lambda t: get_mean_cov(t, self.np_configuration)
        self.potential = LinearPotential(self.initial_configuration["tilde_F"])
        self.distribution_initial = Gaussian(self.initial_configuration["m_0"], self.initial_configuration["P_0"])
        self.distribution_initial_x = Gaussian(self.initial_configuration["m_x_0"],
                                               self.initial_configuration["P_x_0"])
        m_T, P_T = self.get_mean_cov(self.total_evolving_time)
        self.distribution_terminal = Gaussian(torch.as_tensor(m_T, dtype=torch.float32, device=self.device),
                                              torch.as_tensor(P_T, dtype=torch.float32, device=self.device))
        if self.sample_mode == "offline":
            raise NotImplementedError

    def V_true_fn(self, x: torch.Tensor):
        """x . tilde_F x / 2 (OU.py:130-131)."""
        if x.ndim not in (1, 2):
            raise ValueError("x should be either 1D (unbatched) or 2D (batched) array.")
        xx = x[None] if x.ndim == 1 else x
        val = 0.5 * (xx * self.potential.gradient(xx.contiguous())).sum(-1)
        return val[0] if x.ndim == 1 else val

    def true_grad_spec(self):
        return ops.TrueGrad(L.DRIFT_LINEAR, self.initial_configuration["tilde_F"])

    def _sample_at_times(self, times: np.ndarray, per_time: int, rng) -> torch.Tensor:
        """[len(times), per_time, 2d] exact Gaussian samples, one (mean, cov_half) per time stamp."""
        d2 = 2 * self.dim
        mus = np.empty((len(times), d2))
        halves = np.empty((len(times), d2, d2))
        for i, t in enumerate(times):
            if t <= 0.0:
                m, P = self.np_configuration["m_0"], self.np_configuration["P_0"]
            else:
                m, P = lyapunov.kinetic_ou_mean_cov(float(t), self.np_configuration)
            mus[i], halves[i] = m, lyapunov.cov_half(P)
        return ops.gaussian_sample_grouped(
            len(times), per_time, d2, torch.as_tensor(mus, dtype=torch.float32, device=self.device),
            torch.as_tensor(halves, dtype=torch.float32, device=self.device), int(rng))

    def sample_ground_truth(self, rng, batch_size):
        """OU.py:140-190."""
        if isinstance(batch_size, int):
            sample_per_time = 100
            assert batch_size >= sample_per_time * 2
            n_random_time = batch_size // sample_per_time
            rng_time, rng_x = jrandom.split(rng, 2)
            times = self.distribution_time.sample(n_random_time, rng_time, device=self.device).cpu().numpy()
            samples = self._sample_at_times(times, sample_per_time, rng_x)
        else:
            rng_time_shift, rng = jrandom.split(rng)
            n_time_stamps, sample_per_time = batch_size
            time_stamps = self._grid_times(rng_time_shift, n_time_stamps, full=True)
            # OU.py:168-170: prepend t = 0, drop the last stamp; get_mean_cov then drops the leading 0 again
            time_stamps = np.concatenate([np.zeros(1), time_stamps[:-1]])[1:]
            assert n_time_stamps == 1  # OU.py:176 (defect D2 of SURVEY.md §2.3 kept)
            samples = self._sample_at_times(time_stamps, sample_per_time, rng)
            samples = samples.reshape(sample_per_time, n_time_stamps, -1)  # OU.py:184-186
        return samples.reshape((prod(samples.shape[:2]), *samples.shape[2:]))

    def _grid_times(self, rng_time_shift, n_time_stamps, full=False):
        shift = ops.philox_uniforms(n_time_stamps + 1, int(rng_time_shift), device=self.device).cpu().numpy()
        shift = shift * (self.total_evolving_time / n_time_stamps)
        ts = np.linspace(0, self.total_evolving_time, n_time_stamps + 1) + shift
        return ts if full else ts[:-1]

    def get_time_sample_ground_truth(self, rng, batch_size):
        """OU.py:192-207."""
        if isinstance(batch_size, int):
            raise NotImplementedError
        rng_time_shift, rng = jrandom.split(rng)
        ts = self._grid_times(rng_time_shift, batch_size[0])
        self.last_grid_times_host = [float(t) for t in ts]  # the same stamps as Python floats: no device read later
        return torch.as_tensor(ts, dtype=torch.float32, device=self.device)

    def create_parametric_model(self):
        return V_parametric_quadratic(self.dim)
