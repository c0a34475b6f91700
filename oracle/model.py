"""Oracle: hypothesis models (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates core/model.py:32-62 (V_hypothesis) and the parametric models
example_problems/kinetic_fokker_planck_example_GMM.py:214-234 (V_parametric, GMM)
and example_problems/kinetic_fokker_planck_example_OU.py:209-220 (quadratic).

Parameter trees mirror Flax's: {"params": {"layers_0": {"kernel": [in,out],
"bias": [out]}, ...}}.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .potential import gmm_V

OUT_DIM = 40  # core/model.py:43 (hard-coded)


def init_mlp_params(d: int, hidden_dim: int, layers: int, seed: int = 11,
                    dtype=torch.float64) -> Dict:
    """Dense widths d -> [hidden]*layers -> 40; kaiming-normal kernels
    (variance_scaling(2, fan_in, normal): std = sqrt(2 / fan_in)), zero bias
    (core/model.py:42).  JAX's threefry stream is not reproduced: the weights
    are inputs chosen by our seed."""
    g = torch.Generator().manual_seed(seed)
    dims = [d] + [hidden_dim] * layers + [OUT_DIM]
    tree = {}
    for i in range(len(dims) - 1):
        fan_in, fan_out = dims[i], dims[i + 1]
        w = torch.randn(fan_in, fan_out, generator=g, dtype=torch.float64) * math.sqrt(2.0 / fan_in)
        tree[f"layers_{i}"] = {"kernel": w.to(dtype), "bias": torch.zeros(fan_out, dtype=dtype)}
    return {"params": tree}


def mlp_apply(params: Dict, y_input: torch.Tensor) -> torch.Tensor:
    """core/model.py:51-62: Dense/tanh stack, output sum(x**2)[None]."""
    tree = params["params"]
    n = len(tree)
    x = y_input
    for i in range(n):
        layer = tree[f"layers_{i}"]
        x = x @ layer["kernel"] + layer["bias"]
        if i < n - 1:
            x = torch.tanh(x)
    return torch.sum(x ** 2, dim=-1)[None]


def gmm_parametric_apply(params: Dict, y_input: torch.Tensor) -> torch.Tensor:
    """GMM.py:214-234: learnable mus[K,d], sigma = 1."""
    mus = params["params"]["mus"]
    return gmm_V(y_input, mus, 1.0)[None]


def quadratic_parametric_apply(params: Dict, y_input: torch.Tensor) -> torch.Tensor:
    """OU.py:209-220: sum(y * Dense(d)(y)) with Dense kernel [d,d] and bias [d]."""
    p = params["params"]["tilde_F"]
    return torch.sum(y_input * (y_input @ p["kernel"] + p["bias"]), dim=-1)[None]


def flatten_params(params: Dict) -> torch.Tensor:
    """Flat order: sorted layer names, kernel (row-major [in,out]) then bias."""
    out = []
    tree = params["params"]
    for name in sorted(tree.keys(), key=lambda s: (len(s), s)):
        leaf = tree[name]
        if isinstance(leaf, dict):
            out.append(leaf["kernel"].reshape(-1))
            out.append(leaf["bias"].reshape(-1))
        else:
            out.append(leaf.reshape(-1))
    return torch.cat(out)
