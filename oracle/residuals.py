"""Oracle: self-consistency residuals (TEST INFRASTRUCTURE, see oracle/__init__.py).

Literal restatements, with torch.func.{grad,jvp,jacfwd,vmap}, of

  methods/consistency_instances/kinetic_fokker_planck.py:11-69
  methods/consistency_instances/fokker_planck.py:33-63
  methods/consistency_instances/kinetic_mckean_vlasov.py:11-120
  utils/common_utils.py:6-14 (hessian_vector_product), :74-76 (compute_pytree_norm)

`pde` is any object exposing what the reference's residual modules read from
`pde_instance`: V_true_fn / Phi_true_fn (single point -> scalar),
initial_configuration["gamma_friction"], total_evolving_time and (KMV only)
partial_s_log_density_fn / partial_s2_log_density_fn.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch
from torch.func import grad, jacfwd, jvp, vmap


def hessian_vector_product(f: Callable, x, v):
    """utils/common_utils.py:6-14: jvp of grad f."""
    grad_f = grad(f)
    _, hvp = jvp(grad_f, (x,), (v,))
    return hvp


def compute_pytree_norm(tree) -> torch.Tensor:
    """utils/common_utils.py:74-76."""
    leaves = _leaves(tree)
    return torch.sqrt(sum(torch.vdot(g.reshape(-1), g.reshape(-1)) for g in leaves))


def _leaves(tree):
    if isinstance(tree, dict):
        out = []
        for k in tree:
            out += _leaves(tree[k])
        return out
    return [tree]


def _value_and_grad(loss_fn, params):
    g, val = torch.func.grad_and_value(loss_fn)(params)
    return val, g


def kfp_value_and_grad_fn(forward_fn, params, data: Dict, pde, chunk=None) -> Dict:
    """kinetic_fokker_planck.py:11-69."""
    d2 = data["0T"].shape[-1]
    d = d2 // 2
    x_initial, v_initial = data["initial"][:, :d], data["initial"][:, d:]
    x_terminal, v_terminal = data["terminal"][:, :d], data["terminal"][:, d:]
    x_0T, v_0T = data["0T"][:, :d], data["0T"][:, d:]
    gamma = pde.initial_configuration["gamma_friction"]
    T = pde.total_evolving_time

    V = lambda x, p: forward_fn(p, x)[0]
    nabla_V = grad(V, argnums=0)

    def my_prod_single(x, v, p):
        f = lambda xx: V(xx, p)
        return torch.dot(v, hessian_vector_product(f, x, v))

    my_prod = vmap(my_prod_single, in_dims=(0, 0, None), chunk_size=chunk)
    nabla_V_vmap_x = vmap(nabla_V, in_dims=(0, None), chunk_size=chunk)
    nabla_V_true_vmap_x = vmap(grad(pde.V_true_fn), chunk_size=chunk)

    def loss_fn(p):
        loss_initial = torch.mean(torch.sum(nabla_V_vmap_x(x_initial, p) * v_initial, -1))
        loss_terminal = torch.mean(torch.sum(nabla_V_vmap_x(x_terminal, p) * v_terminal, -1))
        loss_nabla = torch.mean(torch.sum(nabla_V_vmap_x(x_0T, p) ** 2, -1))
        loss_Hessian = torch.mean(my_prod(x_0T, v_0T, p))
        loss_friction = torch.mean(torch.sum(nabla_V_vmap_x(x_0T, p) * v_0T, -1)) * gamma
        loss_nabla_true = torch.mean(torch.sum(nabla_V_true_vmap_x(x_0T) ** 2, -1))
        return (loss_nabla - 2 * loss_Hessian + 2 * loss_friction + loss_nabla_true) + (
            -2 * loss_initial + 2 * loss_terminal) / T

    def loss_ground_truth_fn(p):
        return torch.mean(torch.sum((nabla_V_true_vmap_x(x_0T) - nabla_V_vmap_x(x_0T, p)) ** 2, -1), 0)

    loss, g = _value_and_grad(loss_fn, params)
    return {"loss": loss, "grad": g, "grad_norm": compute_pytree_norm(g),
            "loss ground truth": loss_ground_truth_fn(params)}


def fp_value_and_grad_fn(forward_fn, params, data: Dict, pde, chunk=None) -> Dict:
    """fokker_planck.py:33-63 (exact Laplacian from jacfwd of grad)."""
    T = pde.total_evolving_time
    V = lambda x, p: forward_fn(p, x)[0]
    nabla_V = grad(V, argnums=0)
    hessian_V = jacfwd(nabla_V, argnums=0)
    laplacian_V = lambda x, p: torch.sum(torch.diagonal(hessian_V(x, p)))

    V_map_x = vmap(V, in_dims=(0, None), chunk_size=chunk)
    nabla_V_vmap_x = vmap(nabla_V, in_dims=(0, None), chunk_size=chunk)
    laplacian_V_vmap_x = vmap(laplacian_V, in_dims=(0, None), chunk_size=chunk)
    nabla_V_true_vmap_x = vmap(grad(pde.V_true_fn), chunk_size=chunk)

    def loss_fn(p):
        loss_initial = torch.mean(V_map_x(data["initial"], p))
        loss_terminal = torch.mean(V_map_x(data["terminal"], p))
        loss_nabla = torch.mean(torch.sum(nabla_V_vmap_x(data["0T"], p) ** 2, -1))
        loss_laplacian = torch.mean(laplacian_V_vmap_x(data["0T"], p))
        loss_nabla_true = torch.mean(torch.sum(nabla_V_true_vmap_x(data["0T"]) ** 2, -1))
        return (loss_nabla - 2 * loss_laplacian + loss_nabla_true) + (
            2 * loss_terminal - 2 * loss_initial) / T

    def loss_ground_truth_fn(p):
        return torch.mean(torch.sum((nabla_V_true_vmap_x(data["0T"]) - nabla_V_vmap_x(data["0T"], p)) ** 2, -1), 0)

    loss, g = _value_and_grad(loss_fn, params)
    return {"loss": loss, "grad": g, "grad_norm": compute_pytree_norm(g),
            "loss ground truth": loss_ground_truth_fn(params)}


def fp_test_fn(forward_fn, params, data_initial, data_terminal, pde) -> Dict:
    """fokker_planck.py:66-85: relative L2 error of grad V on given samples."""
    V = lambda x: forward_fn(params, x)[0]
    nabla_V_vmap_x = vmap(grad(V))
    nabla_V_true_vmap_x = vmap(grad(pde.V_true_fn))
    out = {}
    for name, data in (("initial", data_initial), ("terminal", data_terminal)):
        pred, true = nabla_V_vmap_x(data), nabla_V_true_vmap_x(data)
        out[f"relative error of gradient estimation {name}"] = torch.sqrt(
            torch.mean(torch.sum((pred - true) ** 2, -1)) / torch.mean(torch.sum(true ** 2, -1)))
    return out


def kmv_value_and_grad_fn(forward_fn, params, data: Dict, pde, m=None) -> Dict:
    """kinetic_mckean_vlasov.py:11-120 (pairwise [m,n,n_time,d] residual, m = n).  `m` (not in the reference, which
    fixes ref_0T = x_0T at :21): use only the first m trajectories as the reference set, the generalisation the
    commented lines :15-16 (`data["ref"]`) point at."""
    d2 = data["0T"].shape[-1]
    d = d2 // 2
    x_0T, v_0T = data["0T"][:, :d], data["0T"][:, d:]
    tau_0T = data["tau_0T"]
    nt = tau_0T.shape[0]
    x_0T = x_0T.reshape(-1, nt, d)  # :19
    v_0T = v_0T.reshape(-1, nt, d)  # :20
    ref_0T = x_0T if m is None else x_0T[:m]  # :21
    x_minus_ref = x_0T[None] - ref_0T[:, None]  # :23  [m,n,nt,d]
    gamma = pde.initial_configuration["gamma_friction"]

    Phi = lambda x, p: forward_fn(p, x)[0]
    nabla_Phi = grad(Phi, argnums=0)

    def my_prod(x, v, p):
        f = lambda xx: Phi(xx, p)
        return torch.dot(v, hessian_vector_product(f, x, v))

    Phi_v = vmap(vmap(vmap(Phi, (0, None)), (0, None)), (0, None))  # :36-38
    nabla_Phi_v = vmap(vmap(vmap(nabla_Phi, (0, None)), (0, None)), (0, None))  # :41-43
    my_prod_v = vmap(vmap(vmap(my_prod, (0, 0, None)), (0, 0, None)), (0, None, None))  # :46-48

    nabla_Phi_true = vmap(vmap(vmap(grad(pde.Phi_true_fn))))  # :50-55

    # :57-72 — vmap over (tau, x[:, t]) gives [n_time, n]; the reference then
    # *reshapes* (not transposes) to [n, n_time] (defect D4, kept verbatim).
    psl = torch.stack([pde.partial_s_log_density_fn(tau_0T[t], x_0T[:, t]) for t in range(nt)], 0)
    psl = psl.reshape(-1, nt)
    ps2l = torch.stack([pde.partial_s2_log_density_fn(tau_0T[t], x_0T[:, t]) for t in range(nt)], 0)
    ps2l = ps2l.reshape(-1, nt)

    def loss_fn(p):
        loss_nabla = nabla_Phi_v(x_minus_ref, p)
        loss_nabla = torch.mean(loss_nabla, 0)
        loss_nabla = torch.sum(loss_nabla ** 2, -1)
        loss_nabla = torch.mean(loss_nabla)

        loss_Hessian = my_prod_v(x_minus_ref, v_0T, p)
        loss_Hessian = torch.mean(torch.mean(loss_Hessian, 0))

        loss_value = Phi_v(x_minus_ref, p)
        loss_value = torch.mean(loss_value, 0)
        loss_value = loss_value * (ps2l + psl ** 2 + gamma * psl)
        loss_value = torch.mean(loss_value)

        loss_nabla_true = torch.mean(torch.sum(torch.mean(nabla_Phi_true(x_minus_ref), 0) ** 2, -1))
        return loss_nabla - 2 * loss_Hessian + 2 * loss_value + loss_nabla_true

    def loss_ground_truth_fn(p):
        return torch.mean(torch.sum(
            (torch.mean(nabla_Phi_true(x_minus_ref), 0) - torch.mean(nabla_Phi_v(x_minus_ref, p), 0)) ** 2, -1))

    loss, g = _value_and_grad(loss_fn, params)
    return {"loss": loss, "grad": g, "grad_norm": compute_pytree_norm(g),
            "loss ground truth": loss_ground_truth_fn(params)}
