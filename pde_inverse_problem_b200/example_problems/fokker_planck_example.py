"""Mirror of example_problems/fokker_planck_example.py (overdamped OU, closed-form law)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
from .. import ops
from ..api import ProblemInstance
from ..core.distribution import Gaussian
from ..core.potential import LinearPotential
from ..utils import rng as jrandom


def initialize_configuration(domain_dim: int, seed: int = 2217):
    """fokker_planck_example.py:20-46 (float64 numpy; F = _F _F^T from our generator)."""
    g = np.random.default_rng(seed)
    d = domain_dim
    _F = g.standard_normal((d, d + 1))
    F = _F @ _F.T * 1.0
    Lm = np.eye(d) * 2.0
    m_0 = np.ones(d) * 1.0
    P_0 = np.eye(d) * 5.0
    U, s, _ = np.linalg.svd(F)
    return {"F": F, "L": Lm, "U": U, "ss": s + s[:, None], "B": U.T @ Lm @ U, "B_0": U.T @ P_0 @ U, "s": s,
            "m_0": m_0, "P_0": P_0}


def OU_process(t, configuration):
    """fokker_planck_example.py:48-55."""
    exp_t_s = np.diag(np.exp(-t * configuration["s"]))
    U = configuration["U"]
    m_t = U @ exp_t_s @ U.T @ configuration["m_0"]
    P_t_1 = exp_t_s @ configuration["B_0"] @ exp_t_s
    B_S = configuration["B"] / configuration["ss"]
    P_t_2 = B_S - exp_t_s @ B_S @ exp_t_s
    return m_t, U @ (P_t_1 + P_t_2) @ U.T


class FokkerPlanck(ProblemInstance):
    """fokker_planck_example.py:63-96."""

    def __init__(self, cfg, rng, device="cuda"):
        super().__init__(cfg, rng, device)
        self.np_configuration = initialize_configuration(cfg.pde_instance.domain_dim)
        dev = self.device
        f32 = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev).contiguous()
        self.initial_configuration = {k: f32(v) for k, v in self.np_configuration.items()}
        self._mt0 = f32(self.np_configuration["U"].T @ self.np_configuration["m_0"])
        self.potential = LinearPotential(self.initial_configuration["F"])
        self.distribution_initial = self.get_distribution(0.0)
        self.distribution_terminal = self.get_distribution(self.total_evolving_time)

    def get_distribution(self, t):
        m, P = OU_process(float(t), self.np_configuration)
        return Gaussian(torch.as_tensor(m, dtype=torch.float32, device=self.device),
                        torch.as_tensor(P, dtype=torch.float32, device=self.device))

    def V_true_fn(self, x: torch.Tensor):
        if x.ndim not in (1, 2):
            raise ValueError("x should be either 1D (unbatched) or 2D (batched) array.")
        xx = x[None] if x.ndim == 1 else x
        val = 0.5 * (xx * self.potential.gradient(xx.contiguous())).sum(-1)
        return val[0] if x.ndim == 1 else val

    def true_grad_spec(self):
        return ops.TrueGrad(L.DRIFT_LINEAR, self.initial_configuration["F"])

    def sample_ground_truth(self, rng, batch_size: int):
        """fokker_planck_example.py:84-96: each sample has its own t ~ U(1e-4, T) and x ~ N(m_t, P_t)."""
        c = self.initial_configuration
        return ops.ou_exact_sample(batch_size, c["U"], c["s"], c["B_0"], c["B"], self._mt0,
                                   self.distribution_time.mins, self.distribution_time.maxs, int(rng))
