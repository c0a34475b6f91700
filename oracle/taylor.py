"""Oracle twin: hand-derived Taylor-mode residual (TEST INFRASTRUCTURE, see oracle/__init__.py).

The CUDA residual kernels do not run autodiff; they evaluate the closed-form
forward Taylor streams and one reverse pass of SURVEY.md §9.1-9.4.  This module
states those formulas in plain batched torch so that tests can check them against
the literal autodiff restatement in oracle/residuals.py (which follows the
reference's jax.grad / jvp / jacfwd graph) before any GPU is involved.  It is a
checker, not a product path.

Per-point scalar handled by `point_set`:
    l(x) = sum_k [ alpha_k D^2_{w_k} V + beta_k D_{w_k} V ] + kappa V + c_g |grad V|^2
with the stop-gradient identity d|g|^2 = 2 d(D_g V)|_{g const} for the last term.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch


def unpack(params: Dict) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    tree = params["params"]
    n = len(tree)
    return [tree[f"layers_{i}"]["kernel"] for i in range(n)], [tree[f"layers_{i}"]["bias"] for i in range(n)]


def forward_primal(W, b, x):
    a = [x]
    z = None
    L = len(W)
    for l in range(L):
        z = a[l] @ W[l] + b[l]
        if l < L - 1:
            a.append(torch.tanh(z))
    return a, z  # a[0..L-1] layer inputs, u = z


def input_gradient(W, a, u):
    L = len(W)
    za = 2 * u
    for l in range(L - 1, -1, -1):
        aa = za @ W[l].T
        if l > 0:
            za = aa * (1 - a[l] ** 2)
    return aa


def point_set(W, b, x, dirs: Sequence[Tuple[torch.Tensor, float, float]], kappa: float, c_g: float,
              scale: float):
    """Returns (sum of l over points * scale, dW list, db list, dict of per-term sums, g).

    dirs: list of (w[N,d], alpha, beta)."""
    L = len(W)
    N = x.shape[0]
    a, u = forward_primal(W, b, x)
    g = input_gradient(W, a, u)
    dW = [torch.zeros_like(w) for w in W]
    db = [torch.zeros_like(bb) for bb in b]
    ubar = torch.zeros_like(u)
    tbar = [None] + [torch.zeros_like(a[l]) for l in range(1, L)]
    V = (u * u).sum(-1)
    total = kappa * V + c_g * (g * g).sum(-1)
    all_dirs = [(w, al, be, True) for (w, al, be) in dirs]
    if c_g != 0.0:
        # stop-gradient direction w = g: contributes to the gradient only (value already has c_g |g|^2)
        all_dirs.append((g.detach(), 0.0, 2.0 * c_g, False))
    terms = {"V": V.sum(), "g2": (g * g).sum(), "D1": [], "D2": []}
    for (w, alpha, beta, in_value) in all_dirs:
        a1 = [w]
        a2 = [torch.zeros_like(w)]
        z1s, z2s = [], []
        for l in range(L):
            z1 = a1[l] @ W[l]
            z2 = a2[l] @ W[l]
            z1s.append(z1)
            z2s.append(z2)
            if l < L - 1:
                t = a[l + 1]
                s1 = 1 - t * t
                s2 = -2 * t * s1
                a1.append(s1 * z1)
                a2.append(s1 * z2 + s2 * z1 * z1)
        u1, u2 = z1s[-1], z2s[-1]
        D1 = 2 * (u * u1).sum(-1)
        D2 = 2 * ((u1 * u1).sum(-1) + (u * u2).sum(-1))
        terms["D1"].append(D1.sum())
        terms["D2"].append(D2.sum())
        if in_value:
            total = total + alpha * D2 + beta * D1
        ubar = ubar + 2 * alpha * u2 + 2 * beta * u1
        zb1 = 4 * alpha * u1 + 2 * beta * u
        zb2 = 2 * alpha * u
        for l in range(L - 1, -1, -1):
            dW[l] += a1[l].T @ zb1 + a2[l].T @ zb2
            if l > 0:
                ab1 = zb1 @ W[l].T
                ab2 = zb2 @ W[l].T
                t = a[l]
                s1 = 1 - t * t
                s2 = -2 * t * s1
                z1p, z2p = z1s[l - 1], z2s[l - 1]
                tbar[l] = tbar[l] + ab1 * z1p * (-2 * t) + ab2 * (z2p * (-2 * t) + z1p * z1p * (6 * t * t - 2))
                zb1 = ab1 * s1 + ab2 * 2 * s2 * z1p
                zb2 = ab2 * s1
    zb = ubar + 2 * kappa * u
    for l in range(L - 1, -1, -1):
        dW[l] += a[l].T @ zb
        db[l] += zb.sum(0)
        if l > 0:
            ab = zb @ W[l].T + tbar[l]
            zb = ab * (1 - a[l] ** 2)
    dW = [scale * m for m in dW]
    db = [scale * m for m in db]
    return scale * total.sum(), dW, db, terms, g


def _pack(dWs, dbs):
    tree = {}
    for i, (w, bb) in enumerate(zip(dWs, dbs)):
        tree[f"layers_{i}"] = {"kernel": w, "bias": bb}
    return {"params": tree}


def kfp_value_and_grad(params, data, gamma: float, T: float, grad_true_0T: Optional[torch.Tensor] = None):
    """Closed-form twin of kinetic_fokker_planck.py:33-61 (loss + grad)."""
    W, b = unpack(params)
    d = data["0T"].shape[1] // 2
    out_W = [torch.zeros_like(w) for w in W]
    out_b = [torch.zeros_like(bb) for bb in b]
    loss = 0.0
    sets = [
        ("0T", [(-2.0, 2.0 * gamma)], 1.0),
        ("terminal", [(0.0, 2.0 / T)], 0.0),
        ("initial", [(0.0, -2.0 / T)], 0.0),
    ]
    g0T = None
    for name, ab, c_g in sets:
        z = data[name]
        x, v = z[:, :d], z[:, d:]
        val, dW, db_, _, g = point_set(W, b, x, [(v, ab[0][0], ab[0][1])], 0.0, c_g, 1.0 / z.shape[0])
        loss = loss + val
        for l in range(len(W)):
            out_W[l] += dW[l]
            out_b[l] += db_[l]
        if name == "0T":
            g0T = g
    res = {"grad": _pack(out_W, out_b)}
    if grad_true_0T is not None:
        loss = loss + (grad_true_0T ** 2).sum(-1).mean()
        res["loss ground truth"] = ((grad_true_0T - g0T) ** 2).sum(-1).mean()
    res["loss"] = loss
    return res


def fp_value_and_grad(params, data, T: float, grad_true_0T: Optional[torch.Tensor] = None):
    """Closed-form twin of fokker_planck.py:47-59 (loss + grad)."""
    W, b = unpack(params)
    d = data["0T"].shape[1]
    out_W = [torch.zeros_like(w) for w in W]
    out_b = [torch.zeros_like(bb) for bb in b]
    loss = 0.0
    g0T = None
    for name, kappa, c_g in (("0T", 0.0, 1.0), ("terminal", 2.0 / T, 0.0), ("initial", -2.0 / T, 0.0)):
        x = data[name]
        dirs = []
        if name == "0T":
            eye = torch.eye(d, dtype=x.dtype)
            dirs = [(eye[i].expand(x.shape[0], d), -2.0, 0.0) for i in range(d)]
        val, dW, db_, _, g = point_set(W, b, x, dirs, kappa, c_g, 1.0 / x.shape[0])
        loss = loss + val
        for l in range(len(W)):
            out_W[l] += dW[l]
            out_b[l] += db_[l]
        if name == "0T":
            g0T = g
    res = {"grad": _pack(out_W, out_b)}
    if grad_true_0T is not None:
        loss = loss + (grad_true_0T ** 2).sum(-1).mean()
        res["loss ground truth"] = ((grad_true_0T - g0T) ** 2).sum(-1).mean()
    res["loss"] = loss
    return res


# ---------------------------------------------------------------------------------------------------
# parametric models (closed forms of SURVEY.md §9.5), checked against autodiff in tests/test_oracle_kat.py
# ---------------------------------------------------------------------------------------------------
def gmm_param_point_set(mus, y, v, alpha: float, beta: float, c_g: float, scale: float):
    """V = -logsumexp(-|y - mu_k|^2 / 2) with learnable mus (GMM.py:214-234).
    l = alpha D_v^2 V + beta D_v V + c_g |grad V|^2.  Returns (sum l * scale, dmus, g, D1, D2)."""
    r = y[:, None, :] - mus[None]                      # [N,K,d]
    a = -0.5 * (r * r).sum(-1)
    w = torch.softmax(a, dim=1)                        # [N,K]
    g = (w[..., None] * r).sum(1)                      # E[r]
    c = (r * v[:, None, :]).sum(-1)                    # c_k = r_k . v
    Ec = (w * c).sum(1, keepdim=True)
    Ec2 = (w * c * c).sum(1, keepdim=True)
    D1 = Ec[:, 0]
    D2 = (v * v).sum(-1) - (Ec2[:, 0] - Ec[:, 0] ** 2)
    cg = (r * g[:, None, :]).sum(-1)
    Ecg = (w * cg).sum(1, keepdim=True)
    beta_g = 2.0 * c_g
    # d/dmu_k D_u V (u const) = -w_k u + w_k (c^u_k - E[c^u]) r_k
    # d/dmu_k D_v^2 V = 2 w_k c_k v - w_k (c_k^2 - E[c^2]) r_k + 2 E[c] (-w_k v + w_k (c_k - E[c]) r_k)
    s_g = -beta_g * w                                                       # coefficient of g
    s_v = -beta * w + alpha * (2 * w * c - 2 * Ec * w)                      # coefficient of v
    s_r = (beta_g * w * (cg - Ecg) + beta * w * (c - Ec)
           + alpha * (-w * (c * c - Ec2) + 2 * Ec * w * (c - Ec)))          # coefficient of r_k
    dmus = s_g.T @ g + s_v.T @ v + s_r.T @ y - mus * s_r.sum(0)[:, None]
    total = alpha * D2 + beta * D1 + c_g * (g * g).sum(-1)
    return scale * total.sum(), scale * dmus, g, D1, D2


def quad_param_point_set(W, b, y, v, alpha: float, beta: float, c_g: float, scale: float):
    """V = y.(yW + b) (OU.py:209-220; Flax Dense kernel W[in,out]).  grad V = (W + W^T) y + b,
    D_v V = grad V . v, D_v^2 V = 2 v^T W v."""
    g = y @ (W + W.T).T + b
    D1 = (g * v).sum(-1)
    D2 = 2 * ((v @ W) * v).sum(-1)
    # d|g|^2/dW_ij = 2 (g_i y_j + g_j y_i); d(v^T W v)/dW_ij = v_i v_j; d(g.v)/dW_ij = v_i y_j + v_j y_i
    dW = c_g * 2 * (g.T @ y + y.T @ g) + alpha * 2 * (v.T @ v) + beta * (v.T @ y + y.T @ v)
    db = c_g * 2 * g.sum(0) + beta * v.sum(0)
    total = alpha * D2 + beta * D1 + c_g * (g * g).sum(-1)
    return scale * total.sum(), scale * dW, scale * db, g, D1, D2
