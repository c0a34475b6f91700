"""Oracle: Philox4x32-10 counter RNG in numpy (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference draws its noise from JAX's threefry2x32 (utils/sampling_utils.py:8,14,27,32).
The B200 path draws noise in registers from Philox4x32-10 (Salmon et al., SC'11) keyed by
(seed, global particle id, step); the stream is *not* the JAX stream (parity runs inject
identical noise tensors on both sides instead).  This module restates the
device generator bit-for-bit for the integer part, so tests can check (i) the uint32 stream
against the published Philox known-answer vectors and (ii) the device normals.

Counter layout (must match csrc/philox.cuh):
    c0 = particle id low 32 bits      c1 = particle id high 32 bits
    c2 = step index (0..S) or a TAG   c3 = block index j (4 outputs per block)
    key = (seed low 32, seed high 32)
Normals: Box-Muller on pairs (x0,x1), (x2,x3):
    u1 = ((x >> 8) + 1) * 2^-24  in (0,1],  u2 = (x' >> 8) * 2^-24 in [0,1)
    r = sqrt(-2 ln u1), theta = pi * (2 u2 - 1);  n0 = r cos(theta), n1 = r sin(theta)
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
TAG_TAU0 = 0xFFFFFFFF
TAG_INIT = 0xFFFFFFFE


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def _box_muller(xa, xb):
    u1 = ((xa >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
    u2 = (xb >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    r = np.sqrt(-2.0 * np.log(u1))
    th = np.pi * (2.0 * u2 - 1.0)
    return r * np.cos(th), r * np.sin(th)


def normals(seed: int, particle_ids: np.ndarray, step: int, d: int) -> np.ndarray:
    """[len(ids), d] float64 normals for one step (device layout: component 4j+i <- block j)."""
    ids = np.asarray(particle_ids, dtype=np.uint64)
    lo = (ids & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (ids >> np.uint64(32)).astype(np.uint32)
    out = np.empty((len(ids), 4 * ((d + 3) // 4)))
    for j in range((d + 3) // 4):
        x0, x1, x2, x3 = philox4x32_10(lo, hi, np.uint32(step), np.uint32(j),
                                       seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        n0, n1 = _box_muller(x0, x1)
        n2, n3 = _box_muller(x2, x3)
        out[:, 4 * j + 0], out[:, 4 * j + 1], out[:, 4 * j + 2], out[:, 4 * j + 3] = n0, n1, n2, n3
    return out[:, :d]


def uniform01(seed: int, particle_ids: np.ndarray, tag: int = TAG_TAU0) -> np.ndarray:
    """[len(ids)] uniforms in [0,1): (x0 >> 8) * 2^-24 of block 0 at step = tag."""
    ids = np.asarray(particle_ids, dtype=np.uint64)
    lo = (ids & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (ids >> np.uint64(32)).astype(np.uint32)
    x0, _, _, _ = philox4x32_10(lo, hi, np.uint32(tag), np.uint32(0),
                                seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return (x0 >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
