"""Counter-style seed handling standing in for jax.random keys.

The reference threads jax PRNG keys through every call (`random.split`, main.py:43-44,
core/trainer.py:80-83, methods/consistency.py:37,54).  Here a key is a Python int (uint64); `split`
derives independent child seeds with SplitMix64, and the device consumes seeds through Philox4x32-10
(csrc/philox.cuh).  Streams are deterministic but are not JAX's threefry streams.
"""
from __future__ import annotations

from typing import List

_MASK = 0xFFFFFFFFFFFFFFFF


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _MASK
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
    return z ^ (z >> 31)


def PRNGKey(seed: int) -> int:
    return _splitmix64(int(seed) & _MASK)


def split(key: int, num: int = 2) -> List[int]:
    key = int(key) & _MASK
    base = _splitmix64(key ^ 0xA5A5A5A5A5A5A5A5)
    return [_splitmix64((base + (i + 1) * 0xD1342543DE82EF95) & _MASK) for i in range(num)]


def fold_in(key: int, data: int) -> int:
    base = _splitmix64((int(key) & _MASK) ^ 0x5A5A5A5A5A5A5A5A)
    return _splitmix64((base + ((int(data) & _MASK) + 1) * 0x9FB21C651E98DF25) & _MASK)


def randint(key: int, low: int, high: int) -> int:
    return low + _splitmix64(key) % (high - low)
