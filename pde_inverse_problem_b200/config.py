"""Attribute-style configuration mirroring the reference's Hydra tree (configurations/**/*.yaml).

Hydra / omegaconf are not part of this image; `Config` gives the same `cfg.a.b.c` access the
reference code uses everywhere (e.g. methods/consistency.py:57-60, core/model.py:110-123) and
`make_config` composes the reference's defaults (configurations/config.yaml:1-5) with
`key=value` overrides like the launch scripts do (scripts/*.sh).
"""
from __future__ import annotations

import copy
from typing import Any, Dict


class Config(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(d: Any) -> Any:
        if isinstance(d, dict):
            return Config({k: Config.wrap(v) for k, v in d.items()})
        return d


# configurations/pde_instance/*.yaml
_PDE = {
    "fokker_planck": dict(domain_dim=4, domain_min=-10.0, domain_max=10.0, boundary_condition="None",
                          diffusion_coefficient=2.0, total_evolving_time=2.0, name="Fokker-Planck",
                          potential="Quadratic"),
    "kinetic_fokker_planck": dict(domain_dim=4, domain_min=-10.0, domain_max=10.0, boundary_condition="None",
                                  diffusion_coefficient=2.0, total_evolving_time=2.0,
                                  name="Kinetic-Fokker-Planck", potential="Quadratic", sample_mode="online",
                                  n_steps=100, sample_initial_size=500000, sample_terminal_size=200000,
                                  sample_0T_size=20000, n_steps_terminal=400, n_steps_0T=400,
                                  n_gaussian=3),  # n_gaussian: a parameter here; hard-coded 3 at GMM.py:19
    "kinetic_mckean_vlasov": dict(domain_dim=4, domain_min=-10.0, domain_max=10.0, boundary_condition="None",
                                  diffusion_coefficient=2.0, total_evolving_time=2.0,
                                  name="Kinetic-McKean-Vlasov", potential="Quadratic", sample_mode="online",
                                  n_steps=100, sample_initial_size=500000, sample_terminal_size=200000,
                                  sample_0T_size=20000, n_steps_terminal=400, n_steps_0T=400),
}
# configurations/solver/ConsistencyBased.yaml
_SOLVER = dict(name="ConsistencyBased",
               train=dict(batch_size_init=50000, batch_size_terminal=50000, batch_size_0T=50000,
                          n_time_stamps=200, sample_per_time=250, sample_mode="random_time"))
# configurations/neural_network/MLP.yaml
_NN = dict(time_embedding_dim=0, n_resblocks=0, initialization="kaiming", hidden_dim=20, layers=8,
           activation="relu")
# configurations/config.yaml
_ROOT = dict(
    backend=dict(use_pmap_train=False, use_pmap_test=False),
    save_and_load=dict(load_model=False, save_model=False, save_frequency=2000, model_directory="./checkpoint"),
    test=dict(batch_size=50000, frequency=100, verbose=False),
    baseline=dict(name="particle method", batch_size=5000),
    plot=dict(batch_size=50000, frequency=2000),
    train=dict(number_of_iterations=80000, batch_size=64,
               optimizer=dict(use_ema=False, method="SGD", momentum=0.9, weight_decay=0.001,
                              learning_rate=dict(initial=0.001, scheduling="None"),
                              grad_clipping=dict(type="adaptive", threshold=1))),
    ODE_tolerance=1e-5, seed=1, estimation_mode="parametric",
)


def _set(d: Dict, dotted: str, value: Any) -> None:
    keys = dotted.split(".")
    for k in keys[:-1]:
        d = d.setdefault(k, {})
    d[keys[-1]] = value


def make_config(pde_instance: str = "fokker_planck", **overrides: Any) -> Config:
    """Compose the defaults list of configurations/config.yaml:1-5, then apply dotted overrides, e.g.
    make_config("kinetic_fokker_planck", **{"pde_instance.potential": "GMM", "neural_network.layers": 2})."""
    cfg = copy.deepcopy(_ROOT)
    cfg["pde_instance"] = copy.deepcopy(_PDE[pde_instance])
    cfg["solver"] = copy.deepcopy(_SOLVER)
    cfg["neural_network"] = copy.deepcopy(_NN)
    for k, v in overrides.items():
        _set(cfg, k, v)
    return Config.wrap(cfg)
