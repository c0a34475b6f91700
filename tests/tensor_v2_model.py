"""Plain-torch statement of the phase schedule of the tcgen05 residual kernel (csrc/residual_tensor.cu, v2).

TEST INFRASTRUCTURE: it restates, phase by phase (P = GEMM phase, E = epilogue), exactly what the kernel computes,
in float64 and without any bf16 rounding, so that (a) the algebra of the schedule (rescaled streams a2^ = -a2/2,
g^ = g/2, input-gradient chain za^ = 4 za; adjoints of the order-2 and g streams taken from the input-gradient chain; pz terms; merged band
c = a2~ + ag~) is checked against oracle/taylor.py on the CPU, and (b) the per-phase probe dumps of the kernel can be
compared with `trace` on the GPU.  KFP 0T set (kinetic_fokker_planck.py:40-45): per point
|g|^2 - 2 D_v^2 V + 2 gamma D_v V, 3-Dense-layer MLP d -> H -> H -> O.
"""
from __future__ import annotations

import torch


def _bf16(t):
    return t.to(torch.float32).to(torch.bfloat16).to(t.dtype)


def kfp_0T_schedule(W, b, x, v, gamma: float, weight: float, mask=None, emulate_bf16: bool = False):
    """W, b: lists of 3 kernels [in,out] / biases; x, v [n,d].  Returns dict(loss, dW, db, D1, D2, g, trace).
    emulate_bf16: round every shared-memory operand and every parked TMEM value to bf16 where the kernel does
    (weights are split hi + lo in the kernel, i.e. kept to ~16 mantissa bits: treated as exact here)."""
    r = _bf16 if emulate_bf16 else (lambda t: t)
    W0, W1, W2 = W
    b0, b1, b2 = b
    n = x.shape[0]
    m = torch.ones(n, 1, dtype=x.dtype) if mask is None else mask.reshape(n, 1).to(x.dtype)
    tr = {}
    x, v = r(x), r(v)
    # P0 / E1
    z0, z10 = x @ W0, v @ W0
    t1 = torch.tanh(z0 + b0)
    s11 = 1 - t1 * t1
    a11 = s11 * z10
    a2t1 = t1 * a11 * z10                    # a2^ = -a2 / 2
    t1, s11, a11, a2t1 = r(t1), r(s11), r(a11), r(a2t1)   # operand bands / parked s1 (epilogues later read these)
    tr.update(t1=t1, s1_1=s11, a1_1=a11, a2t_1=a2t1)
    # P1 / E2
    z1, z11, z21t = t1 @ W1, a11 @ W1, a2t1 @ W1
    t2 = torch.tanh(z1 + b1)
    s12 = 1 - t2 * t2
    a12 = s12 * z11
    a2t2 = s12 * z21t + t2 * a12 * z11
    t2, s12, a12, a2t2 = r(t2), r(s12), r(a12), r(a2t2)
    tr.update(t2=t2, s1_2=s12, a1_2=a12, a2t_2=a2t2)
    # P2 / E3
    u, u1, u2t = t2 @ W2, a12 @ W2, a2t2 @ W2
    uu = u + b2
    za2 = r(8 * uu * m)                      # za^ = 4 za
    D1 = 2 * (uu * u1).sum(-1, keepdim=True)
    D2 = 2 * ((u1 * u1).sum(-1, keepdim=True) - 2 * (uu * u2t).sum(-1, keepdim=True))
    s1v = r((-8 * u1 + 4 * gamma * uu) * m)
    s0p = r((8 * u2t + 4 * gamma * u1) * m)
    tr.update(za2=za2, s1v=s1v, s0p=s0p)
    # P3 / E4
    aa2, ab12 = za2 @ W2.T, s1v @ W2.T
    dW2 = a12.T @ s1v
    za1 = r(aa2 * s12)
    q = aa2 * a12
    zb1p = r(s12 * ab12 + 2 * t2 * q)
    pz2 = r(a12 * (q - 2 * t2 * ab12))
    tr.update(za1=za1, zb1p=zb1p, pz2=pz2)
    # P4 / E5
    aa1, ab11 = za1 @ W1.T, zb1p @ W1.T
    dW1 = a11.T @ zb1p
    za0 = r(aa1 * s11)
    q = aa1 * a11
    zb1pp = r(s11 * ab11 + 2 * t1 * q)
    pz1 = r(a11 * (q - 2 * t1 * ab11))
    tr.update(za0=za0, zb1pp=zb1pp, pz1=pz1)
    # P5 / E6
    g = 0.25 * (za0 @ W0.T)
    dW0 = v.T @ zb1pp
    gt = r(0.5 * g)                          # g^ = g / 2
    tr.update(g=g)
    # P6 / E7
    zg0 = gt @ W0
    ag1 = r(s11 * zg0)
    c1 = r(a2t1 + ag1)
    # P7 / E8
    zg1 = ag1 @ W1
    ag2 = r(s12 * zg1)
    c2 = r(a2t2 + ag2)
    tr.update(c_1=c1, c_2=c2)
    # P8 / E9
    ug = ag2 @ W2
    s0f = s0p + 8 * ug
    db2 = s0f.sum(0)
    s0 = r(s0f)
    tr.update(s0=s0)
    # P9 / E10
    ab2, aa2r = s0 @ W2.T, za2 @ W2.T
    dW2 = dW2 + t2.T @ s0 + c2.T @ za2
    zb0pf = s12 * ab2 + pz2 - 2 * t2 * aa2r * c2
    db1 = zb0pf.sum(0)
    zb0p = r(zb0pf)
    tr.update(zb0p=zb0p)
    # P10 / E11
    ab1_, aa1r = zb0p @ W1.T, za1 @ W1.T
    dW1 = dW1 + t1.T @ zb0p + c1.T @ za1
    zb0ppf = s11 * ab1_ + pz1 - 2 * t1 * aa1r * c1
    db0 = zb0ppf.sum(0)
    zb0pp = r(zb0ppf)
    tr.update(zb0pp=zb0pp)
    # P11
    dW0 = dW0 + x.T @ zb0pp + gt.T @ za0
    g2 = (g * g).sum(-1, keepdim=True)
    loss = weight * (m * (g2 - 2 * D2 + 2 * gamma * D1)).sum()
    return {"loss": loss, "dW": [weight * dW0, weight * dW1, weight * dW2],
            "db": [weight * db0, weight * db1, weight * db2], "D1": D1, "D2": D2, "g": g, "trace": tr}
