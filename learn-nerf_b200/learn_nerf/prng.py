"""jax.random-style keys for the host mirror (render.py:55,142; train.py:137).

``PRNGKey``, ``split``, ``fold_in`` and ``uniform`` follow JAX's (non-partitionable)
Threefry-2x32 construction, so an integer seed produces the uniforms JAX would produce --
as far as that can be established without JAX: the construction is [recalled] and pinned by the
Random123 known-answer vectors and JAX's documented ``split(PRNGKey(0))`` / ``uniform(PRNGKey(0))``
values (tests/test_oracle.py).  Key bookkeeping runs on the host in Python integers (a handful of
32-bit operations per split); the uniforms themselves are generated on the device
(lnrf_threefry_uniform).  Parity tests keep using explicit uniforms.
"""
from dataclasses import dataclass
from typing import Tuple, Union

import torch

_M32 = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _threefry_pair(k0: int, k1: int, x0: int, x1: int) -> Tuple[int, int]:
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0 = (x0 + ks[0]) & _M32
    x1 = (x1 + ks[1]) & _M32
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & _M32
            x1 = ((x1 << r) | (x1 >> (32 - r))) & _M32
            x1 ^= x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M32
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & _M32
    return x0, x1


def _threefry_2x32(k0: int, k1: int, counts):
    counts = list(counts)
    n = len(counts)
    if n % 2:
        counts.append(0)
    half = len(counts) // 2
    a, b = [], []
    for j in range(half):
        y0, y1 = _threefry_pair(k0, k1, counts[j], counts[half + j])
        a.append(y0)
        b.append(y1)
    return (a + b)[:n]


@dataclass(frozen=True)
class PRNGKey:
    """``PRNGKey(seed)`` as jax.random.PRNGKey: words (seed >> 32, seed & 0xFFFFFFFF)."""

    k0: int
    k1: int = None

    def __post_init__(self):
        if self.k1 is None:  # constructed from a seed
            seed = int(self.k0)
            object.__setattr__(self, "k0", (seed >> 32) & _M32)
            object.__setattr__(self, "k1", seed & _M32)

    @property
    def seed(self) -> int:
        """One integer identifying the key (used to seed torch generators for weight init)."""
        return ((self.k0 << 32) | self.k1) & ((1 << 63) - 1)


KeyLike = Union[PRNGKey, int]


def _as_key(key: KeyLike) -> PRNGKey:
    return key if isinstance(key, PRNGKey) else PRNGKey(int(key))


def split(key: KeyLike, num: int = 2) -> Tuple[PRNGKey, ...]:
    k = _as_key(key)
    bits = _threefry_2x32(k.k0, k.k1, range(2 * num))
    return tuple(PRNGKey(bits[2 * i], bits[2 * i + 1]) for i in range(num))


def fold_in(key: KeyLike, data: int) -> PRNGKey:
    k = _as_key(key)
    y = _threefry_2x32(k.k0, k.k1, [0, int(data) & _M32])
    return PRNGKey(y[0], y[1])


def uniform(key: KeyLike, shape, device) -> torch.Tensor:
    """fp32 uniforms in [0, 1) (multiples of 2^-23) generated on ``device`` (CUDA only)."""
    from . import _native
    k = _as_key(key)
    return _native.threefry_uniform(k.k0, k.k1, shape, device)
