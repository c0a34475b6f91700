"""Mirror of api.py: ProblemInstance (:15-64) and Method (:67-104), on torch CUDA tensors."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Tuple, Union

import torch

from .core.distribution import Distribution, Uniform


class ProblemInstance:
    distribution_initial: Distribution
    distribution_initial_x: Distribution
    distribution_terminal: Distribution
    distribution_time: Distribution
    total_evolving_time: float = 1.0
    diffusion_coefficient: float = 0.0
    instance_name: str
    dim: int

    def __init__(self, cfg, rng, device="cuda"):
        """api.py:27-41."""
        self.cfg = cfg
        self.device = torch.device(device)
        self.instance_name = f"{cfg.pde_instance.domain_dim}D-{cfg.pde_instance.name}"
        self.dim = cfg.pde_instance.domain_dim
        self.diffusion_coefficient = float(cfg.pde_instance.diffusion_coefficient)
        self.total_evolving_time = float(cfg.pde_instance.total_evolving_time)
        self.distribution_time = Uniform(0.0001, self.total_evolving_time)  # api.py:35-37
        self.sample_scheme = "exact"  # should either be "exact" or "SDE"
        self.sample_mode = "online"   # should either be "online" or "offline"

    def sample_ground_truth(self, rng, batch_size: Union[int, Tuple[int, int]]):
        pass

    def get_time_sample_ground_truth(self, rng, batch_size: Union[int, Tuple[int, int]]):
        pass

    def generate_ground_truth_dataset(self, rng):
        pass

    def create_parametric_model(self):
        pass

    def true_grad_spec(self):
        """grad V_true for the residual kernels (ops.TrueGrad)."""
        raise NotImplementedError


@dataclass
class Method:
    """api.py:67-104."""
    pde_instance: ProblemInstance
    cfg: Any
    rng: Any

    def value_and_grad_fn(self, forward_fn, params, rng):
        raise NotImplementedError

    def test_fn(self, forward_fn, params, rng):
        pass

    def plot_fn(self, forward_fn, params, rng):
        return  # api.py:81-82: plotting is disabled in the reference

    def create_model_fn(self) -> Tuple[Any, Dict]:
        raise NotImplementedError
