"""Mirror of example_problems/kinetic_fokker_planck_example_GMM.py on the CUDA integrator."""
from __future__ import annotations

from math import prod

import torch

from .. import _lib as L
from .. import ops
from ..api import ProblemInstance
from ..core.distribution import Gaussian
from ..core.model import V_parametric_GMM
from ..core.potential import GMMPotential
from ..utils import rng as jrandom
from ..utils.sampling_utils import underdamped_langevin_dynamics_scan


def initialize_configuration(domain_dim: int, rng, n_Gaussian: int = 3, device="cuda"):
    """GMM.py:16-63.  n_Gaussian is hard-coded to 3 in the reference (:19); it is a parameter here
    (cfg.pde_instance.n_gaussian).  mus ~ U[-4, 4]^d from a torch generator seeded by `rng`."""
    gamma_friction = 0.5
    g = torch.Generator().manual_seed(int(rng) & 0x7FFFFFFFFFFFFFFF)
    mus = (torch.rand(n_Gaussian, domain_dim, generator=g, dtype=torch.float64) * 8.0 - 4.0).float()
    m_0 = torch.zeros(2 * domain_dim)
    P_0 = torch.diag(torch.cat([torch.full((domain_dim,), 4.0), torch.full((domain_dim,), 0.1)]))
    return {
        "n_Gaussian": n_Gaussian,
        "gamma_friction": gamma_friction,
        "m_0": m_0.to(device), "P_0": P_0.to(device),
        "m_x_0": m_0[:domain_dim].to(device), "P_x_0": P_0[:domain_dim, :domain_dim].to(device),
        "GMM": {"mus": mus.to(device).contiguous()},
    }


class KineticFokkerPlanck(ProblemInstance):
    """GMM.py:66-211."""

    def __init__(self, cfg, rng, device="cuda"):
        super().__init__(cfg, rng, device)
        rng_initial_config, rng_dataset = jrandom.split(rng)
        self.initial_configuration = initialize_configuration(
            cfg.pde_instance.domain_dim, rng_initial_config,
            n_Gaussian=int(cfg.pde_instance.get("n_gaussian", 3)), device=self.device)
        self.potential = GMMPotential(self.initial_configuration["GMM"]["mus"], 1.0)  # GMM.py:76-78
        self.sample_scheme = "SDE"
        self.sample_mode = self.cfg.pde_instance.sample_mode
        self.distribution_initial = Gaussian(self.initial_configuration["m_0"], self.initial_configuration["P_0"])
        self.distribution_initial_x = Gaussian(self.initial_configuration["m_x_0"],
                                               self.initial_configuration["P_x_0"])
        if self.sample_mode == "offline":
            self.dataset = self.generate_ground_truth_dataset(rng_dataset)

    def V_true_fn(self, x: torch.Tensor):
        if x.ndim in (1, 2):
            return self.potential.value(x)
        raise ValueError("x should be either 1D (unbatched) or 2D (batched) array.")

    def true_grad_spec(self):
        return ops.TrueGrad(L.DRIFT_GMM, self.potential.mus, sigma=self.potential.sigma)

    def sample_ground_truth(self, rng, batch_size):
        """GMM.py:104-142 (with the 3-tuple of the integrator unpacked correctly; the reference's 2-way
        unpack at :115/:133 is defect D1 of SURVEY.md §2.3)."""
        rng, rng2, rng_init, rng_init2, rng_init3 = jrandom.split(rng, 5)
        multiple_init = 30
        multiple_terminal = 30
        n_steps = self.cfg.pde_instance.n_steps
        dt = self.total_evolving_time / n_steps
        gamma = self.initial_configuration["gamma_friction"]
        q0_p0 = self.distribution_initial.sample(batch_size, rng_init)
        _, sample_0T, _ = underdamped_langevin_dynamics_scan(q0_p0, n_steps, dt, rng, self.potential.gradient, gamma)
        sample_0T = sample_0T.reshape((prod(sample_0T.shape[:2]), *sample_0T.shape[2:]))
        sample_initial = self.distribution_initial.sample(batch_size * multiple_init, rng_init2)
        q0_p02 = self.distribution_initial.sample(batch_size * multiple_terminal, rng_init3)
        sample_final, _, _ = underdamped_langevin_dynamics_scan(q0_p02, n_steps, dt, rng2, self.potential.gradient,
                                                                gamma, want_trajectory=False)
        return sample_initial, sample_final, sample_0T

    def generate_ground_truth_dataset(self, rng):
        """GMM.py:158-204."""
        rng_initial, rng_terminal, rng_0T = jrandom.split(rng, 3)
        gamma = self.initial_configuration["gamma_friction"]
        dataset = {"initial": self.distribution_initial.sample(self.cfg.pde_instance.sample_initial_size, rng_initial)}
        rng_terminal_0, rng_terminal_1 = jrandom.split(rng_terminal)
        n_steps = self.cfg.pde_instance.n_steps_terminal
        dt = self.total_evolving_time / n_steps
        q0_p0 = self.distribution_initial.sample(self.cfg.pde_instance.sample_terminal_size, rng_terminal_0)
        dataset["terminal"], _, _ = underdamped_langevin_dynamics_scan(
            q0_p0, n_steps, dt, rng_terminal_1, self.potential.gradient, gamma, want_trajectory=False)
        rng_0T_0, rng_0T_1 = jrandom.split(rng_0T)
        n_steps = self.cfg.pde_instance.n_steps_0T
        dt = self.total_evolving_time / n_steps
        q0_p0 = self.distribution_initial.sample(self.cfg.pde_instance.sample_0T_size, rng_0T_0)
        _, dataset["0T"], dataset["tau_0T"] = underdamped_langevin_dynamics_scan(
            q0_p0, n_steps, dt, rng_0T_1, self.potential.gradient, gamma)
        return dataset

    def create_parametric_model(self):
        return V_parametric_GMM(dim=self.dim, n_Gaussians=self.initial_configuration["n_Gaussian"])


V_parametric = V_parametric_GMM
