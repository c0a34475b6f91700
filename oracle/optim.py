"""Oracle: optimizer step (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates main.py:11-29 (optax.chain(add_decayed_weights(wd), adam(lr, b1=0.9,
eps=1e-4))) and core/trainer.py:61-70 (step with optional EMA).  optax is an
un-pinned third-party dependency absent from this image; its published
algorithms are restated here:

  add_decayed_weights:  g <- g + wd * p                       (before Adam: L2, not AdamW)
  scale_by_adam:        m <- b1 m + (1-b1) g ;  v <- b2 v + (1-b2) g^2
                        count <- count + 1
                        mhat = m / (1 - b1^count) ; vhat = v / (1 - b2^count)
                        u = mhat / (sqrt(vhat) + eps)          (eps outside sqrt, eps_root = 0)
  scale_by_learning_rate / scale_by_schedule: step = -lr(count_before_increment) * u
  cosine_decay_schedule(init, decay_steps, alpha):
        lr(t) = init * ((1-alpha) * 0.5*(1+cos(pi*min(t,decay_steps)/decay_steps)) + alpha)
  ema(decay, debias=True): ema <- decay*ema + (1-decay)*p ; the trainer reads the RAW
        ema_state.ema (core/trainer.py:68-69), not the debiased output.
"""
from __future__ import annotations

import math
from typing import Dict

import torch


def cosine_decay_schedule(init_value: float, decay_steps: int = 20000, alpha: float = 0.001):
    """main.py:16."""
    def lr(count: int) -> float:
        c = min(count, decay_steps)
        cosine = 0.5 * (1 + math.cos(math.pi * c / decay_steps))
        return init_value * ((1 - alpha) * cosine + alpha)
    return lr


def constant_schedule(value: float):
    return lambda count: value


class AdamL2State:
    def __init__(self, flat_params: torch.Tensor):
        self.count = 0
        self.m = torch.zeros_like(flat_params)
        self.v = torch.zeros_like(flat_params)


def adam_l2_step(p: torch.Tensor, g: torch.Tensor, st: AdamL2State, lr_schedule,
                 weight_decay: float = 1e-3, b1: float = 0.9, b2: float = 0.999,
                 eps: float = 1e-4) -> torch.Tensor:
    """One optimizer.update + apply_updates on flat tensors (main.py:20-26, trainer.py:63-64)."""
    g = g + weight_decay * p
    st.m = b1 * st.m + (1 - b1) * g
    st.v = b2 * st.v + (1 - b2) * g * g
    lr = lr_schedule(st.count)
    st.count += 1
    mhat = st.m / (1 - b1 ** st.count)
    vhat = st.v / (1 - b2 ** st.count)
    u = mhat / (torch.sqrt(vhat) + eps)
    return p - lr * u


def ema_update(ema: torch.Tensor, p_new: torch.Tensor, decay: float = 0.999) -> torch.Tensor:
    """core/trainer.py:67-69 — raw EMA accumulator, which then overwrites params."""
    return decay * ema + (1 - decay) * p_new
