"""Mirror of methods/consistency.py:17-122 (ConsistencyBased: create_model_fn, test_fn, value_and_grad_fn,
sample_data)."""
from __future__ import annotations

from functools import partial
from math import prod

import torch

from .. import ops
from ..api import Method
from ..utils import rng as jrandom
from .consistency_instances import fokker_planck, kinetic_fokker_planck, kinetic_mckean_vlasov

INSTANCES = {
    "Fokker-Planck": fokker_planck,
    "Kinetic-Fokker-Planck": kinetic_fokker_planck,
    "Kinetic-McKean-Vlasov": kinetic_mckean_vlasov,
}


class ConsistencyBased(Method):
    def create_model_fn(self):
        if self.cfg.pde_instance.name in INSTANCES:
            return INSTANCES[self.cfg.pde_instance.name].create_model_fn(self.pde_instance)
        raise NotImplementedError

    def test_fn(self, forward_fn, params, rng):
        forward_fn = partial(forward_fn, params)
        if self.cfg.pde_instance.name in INSTANCES:
            return INSTANCES[self.cfg.pde_instance.name].test_fn(forward_fn=forward_fn,
                                                                 pde_instance=self.pde_instance, rng=rng)
        raise NotImplementedError

    def value_and_grad_fn(self, forward_fn, params, rng):
        rng_sample, rng_vg = jrandom.split(rng, 2)
        data = self.sample_data(rng_sample)
        if self.cfg.pde_instance.name in INSTANCES:
            return INSTANCES[self.cfg.pde_instance.name].value_and_grad_fn(
                forward_fn=forward_fn, params=params, data=data, rng=rng_vg, pde_instance=self.pde_instance)
        raise NotImplementedError

    def sample_data(self, rng):
        """methods/consistency.py:52-122."""
        pde, cfg = self.pde_instance, self.cfg
        if pde.sample_mode == "online":
            rng_initial, rng_terminal, rng_0T = jrandom.split(rng, 3)
            if pde.sample_scheme == "exact":
                batch_size_0T = {
                    "random_time": cfg.solver.train.batch_size_0T,
                    "grid_time": (cfg.solver.train.n_time_stamps, cfg.solver.train.sample_per_time),
                }
                bs = batch_size_0T[cfg.solver.train.sample_mode]
                data = {
                    "initial": pde.distribution_initial.sample(cfg.solver.train.batch_size_init, rng_initial),
                    "terminal": pde.distribution_terminal.sample(cfg.solver.train.batch_size_terminal, rng_terminal),
                    "0T": pde.sample_ground_truth(rng_0T, bs),
                }
                if isinstance(bs, tuple):  # the reference calls this for every mode; only grid_time implements it
                    data["tau_0T"] = pde.get_time_sample_ground_truth(rng_0T, bs)
                    host = getattr(pde, "last_grid_times_host", None)
                    if host is not None:  # KMV residual: per-time-stamp coefficients are looked up on the host
                        data["tau_0T_host"] = host
            elif pde.sample_scheme == "SDE":
                data = {}
                data["initial"], data["terminal"], data["0T"] = pde.sample_ground_truth(
                    rng_0T, cfg.solver.train.batch_size_0T)
            else:
                raise ValueError("unknown sampling scheme")
        elif pde.sample_mode == "offline":
            data = {"initial": pde.dataset["initial"], "terminal": pde.dataset["terminal"]}
            rng_time, rng_sample = jrandom.split(rng)
            n_trajectories, n_time_stamps_0T, _ = pde.dataset["0T"].shape
            interval_time = 5
            shift = jrandom.randint(rng_time, 0, interval_time)                      # :104
            interval_sample = 5
            g = torch.Generator().manual_seed(int(rng_sample) & 0x7FFFFFFFFFFFFFFF)
            perm = torch.randperm(n_trajectories, generator=g)                      # :111
            sample_index = perm[: n_trajectories // interval_sample].to(pde.dataset["0T"].device)
            # :115-118 — dataset[idx][:, time_idx, :].reshape(-1, 2d) as one device gather
            data["0T"] = ops.gather_0T(pde.dataset["0T"], sample_index, interval_time, shift)
        else:
            raise ValueError("unknown sampling mode")
        return data
