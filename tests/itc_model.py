"""CPU twin of the tcgen05 GMM-drift arithmetic of csrc/integrator_tc.cu (test infrastructure).

Restates, with torch bfloat16 casts and float32 accumulation, exactly what the kernel feeds the tensor cores:
centres scaled to log2 units and split hi + lo, the bias -|mu~|^2 / (2 c) as three bf16 terms through a ones column,
x split hi + lo, GEMM 1 = x_hi mu_hi^T + x_lo mu_hi^T + x_hi mu_lo^T + bias (the lo x lo term is dropped), softmax
weights e = 2^(L - max L) split hi + lo, GEMM 2 = e_hi mu_hi + e_lo mu_hi + e_hi mu_lo, and
grad U = (x - A / (c sum e)) / sigma^2   (core/potential.py:32-37 in closed form).
"""
import math

import torch


def _split(v: torch.Tensor, terms: int = 2):
    out, r = [], v.float()
    for _ in range(terms):
        h = r.to(torch.bfloat16).float()
        out.append(h)
        r = r - h
    return out


def gmm_grad_tensor_model(x: torch.Tensor, mus: torch.Tensor, sigma: float = 1.0, k_pad: int = 0) -> torch.Tensor:
    """x [N, d], mus [K, d] (any float dtype) -> grad U [N, d] float32, by the kernel's split-operand formulation."""
    inv_s2 = 1.0 / (sigma * sigma)
    cs = torch.tensor(inv_s2 * 1.4426950408889634, dtype=torch.float32)
    mu_hi, mu_lo = _split(cs * mus.float())
    mu_rec = mu_hi + mu_lo
    b = -0.5 * (mu_rec * mu_rec).sum(-1) / cs
    if k_pad > mus.shape[0]:  # padding centres: zero rows, bias -1e30 (weight exactly 0)
        z = torch.zeros(k_pad - mus.shape[0], mus.shape[1])
        mu_hi, mu_lo = torch.cat([mu_hi, z]), torch.cat([mu_lo, z])
        b = torch.cat([b, torch.full((k_pad - mus.shape[0],), -1.0e30)])
    b0, b1, b2 = _split(b, 3)
    x_hi, x_lo = _split(x)
    logits = x_hi @ mu_hi.T + x_lo @ mu_hi.T + x_hi @ mu_lo.T + (b0 + b1 + b2)
    m = logits.max(-1, keepdim=True).values
    e = torch.exp2(logits - m)
    se = e.sum(-1, keepdim=True)
    e_hi, e_lo = _split(e)
    acc = e_hi @ mu_hi + e_lo @ mu_hi + e_hi @ mu_lo
    return (x.float() - acc / (cs * se)) * inv_s2
