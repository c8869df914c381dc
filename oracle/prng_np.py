"""numpy restatement of the jax.random pieces the hot path touches (render.py:55,142; train.py:137).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY: JAX is not installable here, so this
follows the published Threefry-2x32 algorithm (Salmon et al., "Parallel random numbers: as easy
as 1, 2, 3", SC'11; 20 rounds) and JAX's use of it as [recalled] from jax/_src/prng.py of the
0.3/0.4 series the reference imports (`jax._src.prng.PRNGKeyArray`, non-partitionable threefry):

  PRNGKey(seed)        -> uint32[2] = [seed >> 32, seed & 0xFFFFFFFF]
  threefry_2x32(k, c)  -> c padded to even length, split in two halves x0 | x1, 20 rounds, concat
  split(key, num)      -> threefry_2x32(key, arange(2 num)).reshape(num, 2)
  fold_in(key, data)   -> threefry_2x32(key, [0, data])
  uniform(key, shape)  -> bits = threefry_2x32(key, arange(size)); ((bits >> 9) | 0x3F800000) as f32 - 1

Pinned by the Random123 known-answer vectors that JAX's own test-suite uses and by the
documented outputs of `split(PRNGKey(0))` and `uniform(PRNGKey(0))` (tests/test_oracle.py).
"""
import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def prng_key(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], np.uint32)


def _rotl(x, r):
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def threefry_2x32(key, count) -> np.ndarray:
    key = np.asarray(key, np.uint32)
    count = np.asarray(count, np.uint32).ravel()
    n = count.size
    if n % 2:
        count = np.concatenate([count, np.zeros(1, np.uint32)])
    half = count.size // 2
    x0, x1 = count[:half].copy(), count[half:].copy()
    ks = [key[0], key[1], key[0] ^ key[1] ^ np.uint32(0x1BD11BDA)]
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 += x1
                x1 = _rotl(x1, r)
                x1 ^= x0
            x0 += ks[(i + 1) % 3]
            x1 += ks[(i + 2) % 3] + np.uint32(i + 1)
    out = np.concatenate([x0, x1])
    return out[:n]


def split(key, num: int = 2) -> np.ndarray:
    return threefry_2x32(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)


def fold_in(key, data: int) -> np.ndarray:
    return threefry_2x32(key, np.array([0, int(data) & 0xFFFFFFFF], np.uint32))


def random_bits(key, size: int) -> np.ndarray:
    return threefry_2x32(key, np.arange(size, dtype=np.uint32))


def uniform(key, shape) -> np.ndarray:
    size = int(np.prod(shape))
    bits = random_bits(key, size)
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    return np.maximum(np.float32(0.0), f).reshape(shape)
