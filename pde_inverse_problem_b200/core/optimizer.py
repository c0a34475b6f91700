"""Mirror of main.py:11-29 (get_optimizer): optax.chain(add_decayed_weights(wd), adam(lr, b1=0.9, eps=1e-4))
with a constant or cosine-decay learning rate, executed by the K6 CUDA kernel."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch

from .. import ops


def cosine_decay_schedule(init_value: float, decay_steps: int, alpha: float) -> Callable[[int], float]:
    """optax.cosine_decay_schedule as called at main.py:16 (decay_steps=20000, alpha=0.001)."""
    def schedule(count: int) -> float:
        c = min(count, decay_steps)
        return init_value * ((1 - alpha) * 0.5 * (1 + math.cos(math.pi * c / decay_steps)) + alpha)
    return schedule


@dataclass
class OptState:
    count: int
    m: torch.Tensor
    v: torch.Tensor


class AdamL2:
    """`optimizer.init(params)` / `optimizer.update_and_apply(...)` — one fused kernel does what
    optimizer.update + optax.apply_updates (+ ema.update) do in core/trainer.py:61-70."""

    def __init__(self, lr_schedule: Callable[[int], float], weight_decay: float, b1: float = 0.9,
                 b2: float = 0.999, eps: float = 1e-4):
        self.lr_schedule, self.weight_decay, self.b1, self.b2, self.eps = lr_schedule, weight_decay, b1, b2, eps

    def init(self, params: Dict) -> OptState:
        flat = params["_flat"]
        return OptState(0, torch.zeros_like(flat), torch.zeros_like(flat))

    def step(self, params: Dict, grad: Dict, state: OptState, ema: Optional[torch.Tensor] = None,
             use_ema: bool = False, grad_scale: float = 1.0, norms: Optional[torch.Tensor] = None) -> torch.Tensor:
        """In place on params["_flat"].  Returns norms = [grad_norm, params_norm] (device tensor)."""
        lr = self.lr_schedule(state.count)  # schedule read at the 0-based count
        state.count += 1                    # Adam bias correction uses count + 1
        return ops.adam_l2_step(params["_flat"], grad["_flat"], state.m, state.v, count=state.count, lr=lr,
                                b1=self.b1, b2=self.b2, eps=self.eps, weight_decay=self.weight_decay,
                                grad_scale=grad_scale, ema=ema, use_ema=use_ema, norms=norms)


def get_optimizer(optimizer_cfg) -> AdamL2:
    """main.py:11-29."""
    if optimizer_cfg.method == "SGD":  # (sic) the reference's "SGD" branch builds Adam
        if optimizer_cfg.learning_rate.scheduling == "None":
            init = float(optimizer_cfg.learning_rate.initial)
            lr_schedule = lambda count: init
        elif optimizer_cfg.learning_rate.scheduling == "cosine":
            lr_schedule = cosine_decay_schedule(float(optimizer_cfg.learning_rate.initial), 20000, 0.001)
        else:
            raise NotImplementedError
        return AdamL2(lr_schedule, float(optimizer_cfg.weight_decay))
    raise NotImplementedError
