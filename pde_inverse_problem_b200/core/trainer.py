"""Mirror of core/trainer.py:14-132 (JaxTrainer.fit) driving the CUDA path.

Differences that do not change results: wandb is replaced by a `log_fn` callback (wandb is not part of this
image); the per-iteration NaN assert (trainer.py:112) is kept but can be evaluated every `nan_check_every`
iterations so the loop does not force a device->host sync each step (nan_check_every=1 reproduces the
reference); data parallelism is one process per GPU + one packed all-reduce instead of pmap + host mean
(trainer.py:44-53).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from .. import parallel
from ..utils import rng as jrandom
from .optimizer import AdamL2


class JaxTrainer:
    """Name kept for drop-in compatibility with the reference's `from core.trainer import JaxTrainer`."""

    EMA_START_EPOCH = 40000  # core/trainer.py:87,97 (hard-coded in the reference; an attribute here so tests can reach it)

    def __init__(self, cfg, method, rng, optimizer: AdamL2, forward_fn, params,
                 log_fn: Optional[Callable[[Dict, int], None]] = None, nan_check_every: int = 1):
        self.cfg = cfg
        self.forward_fn = forward_fn
        self.params = params
        self.optimizer = optimizer
        self.method = method
        self.rng = rng
        self.log_fn = log_fn
        self.nan_check_every = max(1, int(nan_check_every))

    def value_and_grad_fn_efficient(self, params, rng):
        """trainer.py:44-56.  With use_pmap_train and more than one rank, every rank draws an independent
        batch (its own rng) and the whole output dict is averaged (pmap + tree mean, :45-52)."""
        shard = parallel.Shard.current()
        if self.cfg.backend.use_pmap_train and shard.world > 1:
            rngs = jrandom.split(rng, shard.world)
            out = self.method.value_and_grad_fn(self.forward_fn, params, rngs[shard.rank])
            red = parallel.allreduce_mean_dict(
                {"loss": out["loss"], "grad_norm": out["grad_norm"], "loss ground truth": out["loss ground truth"],
                 "grad": out["grad"]["_flat"]}, ["loss", "grad_norm", "loss ground truth", "grad"])
            out["grad"]["_flat"].copy_(red["grad"])
            for k in ("loss", "grad_norm", "loss ground truth"):
                out[k] = red[k]
            return out
        return self.method.value_and_grad_fn(self.forward_fn, params, rng)

    def fit(self):
        cfg = self.cfg
        opt_state = self.optimizer.init(self.params)
        ema = torch.zeros_like(self.params["_flat"])  # optax.ema(0.999).init (trainer.py:36-37)
        norms = torch.empty(2, device=self.params["_flat"].device, dtype=torch.float32)
        n_iter = cfg.train.number_of_iterations
        rngs = jrandom.split(self.rng, n_iter)
        pending_loss = []
        for epoch in range(n_iter):
            rng_train, rng_test, rng_plot = jrandom.split(rngs[epoch], 3)
            v_g_etc = self.value_and_grad_fn_efficient(self.params, rng_train)
            use_ema = False
            if cfg.train.optimizer.use_ema and epoch >= self.EMA_START_EPOCH:      # trainer.py:87-103
                if epoch == self.EMA_START_EPOCH:
                    ema.copy_(self.params["_flat"])                  # EmaState(count=0, ema=params)
                use_ema = True
            self.optimizer.step(self.params, v_g_etc["grad"], opt_state, ema=ema, use_ema=use_ema, norms=norms)
            v_g_etc.pop("grad")
            v_g_etc["params_norm"] = norms[1].clone()               # trainer.py:110
            pending_loss.append(v_g_etc["loss"])
            if (epoch + 1) % self.nan_check_every == 0 or epoch == n_iter - 1:
                losses = torch.stack(pending_loss)
                assert not bool(torch.isnan(losses).any()), "loss is NaN"  # trainer.py:112
                pending_loss = []
            if self.log_fn is not None:
                self.log_fn(v_g_etc, epoch)                         # wandb.log(v_g_etc, step=epoch), :113
            if (epoch % cfg.test.frequency == 0 and self.method.test_fn is not None) or epoch >= n_iter - 3:
                result_epoch = self.method.test_fn(self.forward_fn, self.params, rng_test)
                if self.log_fn is not None and result_epoch:
                    self.log_fn(result_epoch, epoch)
                if cfg.test.verbose:
                    msg = f"In epoch {epoch + 1: 5d}, "
                    for key in v_g_etc:
                        msg = msg + f"{key} is {float(v_g_etc[key]): .3e}, "
                    for key in (result_epoch or {}):
                        msg = msg + f"{key} is {float(result_epoch[key]): .3e}, "
                    print(msg)
            if (epoch + 1) % cfg.plot.frequency == 0 and self.method.plot_fn is not None:
                self.method.plot_fn(self.forward_fn, self.params, rng_plot)
        return self.params


Trainer = JaxTrainer
