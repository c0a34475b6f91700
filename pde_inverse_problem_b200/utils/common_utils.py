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


def hessian_vector_product(f, x: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """utils/common_utils.py:6-14: (f, x, v) -> H_f(x) v, with f = functools.partial(net.apply, params) (or any
    callable that model_of / params_of can resolve), x, v: [d] or [N, d].  The fused kernels evaluate the quadratic
    form w'H w (Taylor stream of order 2); the vector comes from it by polarisation,
    (Hv)_i = ((v + e_i)'H(v + e_i) - (v - e_i)'H(v - e_i)) / 4, as ONE model_eval launch over 2 d N directions.
    The residual kernels never go through here (they use v'Hv directly, kinetic_fokker_planck.py:20-23)."""
    from ..core.model import model_of
    model = model_of(f)
    params = getattr(f, "args", (None,))[0] if hasattr(f, "func") else None
    if params is None:
        raise NotImplementedError("hessian_vector_product needs f = functools.partial(net.apply, params)")
    single = x.ndim == 1
    x2, v2 = (x[None], v[None]) if single else (x, v)
    n, d = x2.shape
    eye = torch.eye(d, device=x2.device, dtype=x2.dtype)
    dirs = torch.cat([v2[:, None, :] + eye[None], v2[:, None, :] - eye[None]], dim=1)  # [n, 2d, d]
    xx = x2[:, None, :].expand(n, 2 * d, d).reshape(-1, d).contiguous()
    q = ops.model_eval(model.spec, model.flat(params), xx, dirs.reshape(-1, d).contiguous(), want=("vHv",))["vHv"]
    q = q.view(n, 2, d)
    hv = 0.25 * (q[:, 0] - q[:, 1])
    return hv[0] if single else hv


def vHv(model, params_flat: torch.Tensor, x: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """v'H v of the model at x: jnp.dot(v, hessian_vector_product(V, x, v)) as every reference caller composes it
    (kinetic_fokker_planck.py:20-23), one Taylor stream of order 2 per point."""
    return ops.model_eval(model.spec, params_flat, x, v, want=("vHv",))["vHv"]
