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


class DeviceKey:
    """A Threefry key whose two words live in DEVICE memory (an int32 tensor of 2 elements): the
    form a CUDA-graph replay needs, where the host rewrites the words before every replay.  Only
    ``uniform`` accepts it; splitting is done on the host before the words are uploaded."""

    def __init__(self, words: torch.Tensor):
        assert words.numel() == 2 and words.is_cuda and words.dtype in (torch.int32, torch.uint32)
        self.words = words


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
    if isinstance(key, DeviceKey):
        return _native.threefry_uniform_dk(key.words, shape)
    k = _as_key(key)
    return _native.threefry_uniform(k.k0, k.k1, shape, device)


# ---------------------------------------------------------------------------- host-side streams
# Vectorised (numpy) Threefry for the dataset feeder (dataset.py:206-248), where whole arrays of
# random bits are needed on the host.  Same stream as the device kernel / the scalar code above.
def random_bits_host(key: KeyLike, size: int):
    """uint32[size] = threefry_2x32(key, arange(size)) (jax.random.bits for 32-bit words)."""
    import numpy as np
    k = _as_key(key)
    count = np.arange(size, dtype=np.uint32)
    if size % 2:
        count = np.concatenate([count, np.zeros(1, np.uint32)])
    half = count.size // 2
    x0, x1 = count[:half].copy(), count[half:].copy()
    ks = [np.uint32(k.k0), np.uint32(k.k1), np.uint32(k.k0 ^ k.k1 ^ 0x1BD11BDA)]
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 += x1
                x1 = (x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))
                x1 ^= x0
            x0 += ks[(i + 1) % 3]
            x1 += ks[(i + 2) % 3] + np.uint32(i + 1)
    return np.concatenate([x0, x1])[:size]


def randint_host(key: KeyLike, size: int, minval: int, maxval: int):
    """jax.random.randint(key, [size], minval, maxval) [recalled from jax/_src/random.py, 0.4 series;
    not verifiable here]: two 32-bit draws hi, lo from split(key);
    offset = ((hi % span) * (2^32 % span) + lo % span) % span, i.e. a 64-bit draw modulo span."""
    import numpy as np
    k1, k2 = split(key)
    span = np.uint32(maxval - minval)
    hi, lo = random_bits_host(k1, size), random_bits_host(k2, size)
    mult = np.uint32((int(2 ** 16 % int(span)) ** 2) % int(span))
    with np.errstate(over="ignore"):
        off = ((hi % span) * mult + (lo % span)) % span
    return (off.astype(np.int64) + minval).astype(np.int32)


def permutation_host(key: KeyLike, n: int):
    """Index permutation in the manner of jax.random.permutation / _shuffle [recalled; not
    verifiable here]: ceil(3 ln n / ln(2^32 - 1)) rounds, each a stable sort by fresh 32-bit keys."""
    import math
    import numpy as np
    idx = np.arange(n)
    if n <= 1:
        return idx
    rounds = int(math.ceil(3 * math.log(max(1, n)) / math.log(2 ** 32 - 1)))
    k = _as_key(key)
    for _ in range(rounds):
        k, sub = split(k)
        idx = idx[np.argsort(random_bits_host(sub, n), kind="stable")]
    return idx
