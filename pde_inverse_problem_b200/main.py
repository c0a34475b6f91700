"""Mirror of main.py:32-74 without Hydra / wandb: build problem, method, model, optimizer, trainer and fit.

    python -m pde_inverse_problem_b200.main kinetic_fokker_planck pde_instance.potential=GMM \
        estimation_mode=non-parametric neural_network.hidden_dim=32 neural_network.layers=2 \
        train.number_of_iterations=100
"""
from __future__ import annotations

import sys

from .config import make_config
from .core.optimizer import get_optimizer
from .core.trainer import JaxTrainer
from .registry import get_method, get_pde_instance
from .utils import rng as jrandom


def run(cfg, log_fn=None, device="cuda"):
    seeds_keys = ["rng_problem", "rng_method", "rng_trainer", "rng_log_density"]
    seeds = dict(zip(seeds_keys, jrandom.split(jrandom.PRNGKey(cfg.seed), len(seeds_keys))))   # main.py:43-44
    pde_instance = get_pde_instance(cfg)(cfg=cfg, rng=seeds["rng_problem"], device=device)     # :47
    method = get_method(cfg)(pde_instance=pde_instance, cfg=cfg, rng=seeds["rng_method"])      # :53
    net, params = method.create_model_fn()                                                     # :56
    optimizer = get_optimizer(cfg.train.optimizer)                                             # :59
    trainer = JaxTrainer(cfg=cfg, method=method, rng=seeds["rng_trainer"], forward_fn=net.apply,
                         params=params, optimizer=optimizer, log_fn=log_fn)                    # :62-63
    return trainer.fit(), pde_instance, net                                                    # :66


def _parse(v: str):
    for cast in (int, float):
        try:
            return cast(v)
        except ValueError:
            pass
    return {"True": True, "False": False}.get(v, v)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    pde = "fokker_planck"
    overrides = {}
    for a in argv:
        if "=" in a:
            k, v = a.split("=", 1)
            if k == "pde_instance":
                pde = v
            else:
                overrides[k] = _parse(v)
        else:
            pde = a
    cfg = make_config(pde, **overrides)
    losses = []
    run(cfg, log_fn=lambda d, step: losses.append((step, {k: float(v) for k, v in d.items()})))
    print(losses[-1])


if __name__ == "__main__":
    main()
