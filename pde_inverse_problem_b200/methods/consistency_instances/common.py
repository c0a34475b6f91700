"""Shared driver of the three residual modules: point sets -> libpdeip residual kernels -> result dict."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from ... import _lib as L
from ... import ops
from ...core.model import model_of

_ACC_CACHE: Dict[Tuple, ops.ResidualAccumulator] = {}

# arithmetic path of the MLP residual: PATH_FP32 (parity, rtol 1e-5) or PATH_TENSOR (tcgen05, rtol 1e-2)
DEFAULT_PATH = {"path": L.PATH_FP32}


def accumulator_for(model, device) -> ops.ResidualAccumulator:
    key = (model.spec.kind, model.spec.args(), str(device))
    acc = _ACC_CACHE.get(key)
    if acc is None:
        acc = ops.ResidualAccumulator(model.spec, device=device)
        _ACC_CACHE[key] = acc
    return acc


def result_dict(model, params, sums: torch.Tensor, grad_flat: torch.Tensor) -> Dict:
    """{"loss", "grad" (same tree as params), "grad_norm", "loss ground truth"} —
    kinetic_fokker_planck.py:64-69 / fokker_planck.py:63 / kinetic_mckean_vlasov.py:115-120."""
    grad_tree = model.tree(grad_flat.clone())
    return {
        "loss": sums[L.SUM_LOSS].clone(),
        "grad": {"params": grad_tree["params"], "_flat": grad_tree["_flat"]},
        "grad_norm": sums[L.SUM_GRADNORM].clone(),
        "loss ground truth": sums[L.SUM_GT].clone(),
    }
