"""Mirror of utils/common_utils.py for the symbols the hot path uses (:6-14, :64, :74-76)."""
from __future__ import annotations

import torch

from .. import ops


def compute_pytree_norm(pytree) -> torch.Tensor:
    """utils/common_utils.py:74-76.  Leaves that are views of one flat buffer reduce to one norm; the
    residual / optimizer kernels already emit this value (PDEIP_SUM_GRADNORM, adam norms), this helper is
    for callers that hold an arbitrary tree."""
    leaves = _leaves(pytree)
    flat = torch.cat([x.reshape(-1) for x in leaves])
    # sum of squares through the K7 kernel: the [n,1] "ensemble" has second moment sum x^2
    _, s2 = ops.ensemble_moments(flat.view(-1, 1).contiguous())
    return torch.sqrt(s2.reshape(()))


def _leaves(tree):
    if isinstance(tree, dict):
        out = []
        for k in tree:
            out += _leaves(tree[k])
        return out
    return [tree]


def v_matmul(A: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """utils/common_utils.py:64: vmap(matmul, (None, 0)) == rows x_n -> A x_n."""
    return ops.linear_grad(x, A)


def hessian_vector_product(model, params_flat: torch.Tensor, x: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """v^T H v of the model at x (utils/common_utils.py:6-14 composed with jnp.dot(v, .) as every caller
    does, kinetic_fokker_planck.py:20-23)."""
    return ops.model_eval(model.spec, params_flat, x, v, want=("vHv",))["vHv"]
