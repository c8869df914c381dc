"""Counter-style PRNG keys standing in for jax.random keys.

The reference threads ``jax.random`` keys through render_rays/step_fn
(render.py:55,142; train.py:137).  JAX's threefry2x32 stream is not reproduced
(needs a real JAX to validate), so a key -> uniforms mapping here
is deterministic but NOT bit-identical to JAX's.  Parity with the oracle is always
established through the explicit-uniforms entry points instead.

``uniform`` matches jax.random.uniform's fp32 construction: 23 random mantissa bits,
i.e. multiples of 2^-23 in [0, 1).
"""
from dataclasses import dataclass
from typing import Tuple, Union

import torch

_MASK = (1 << 63) - 1


@dataclass(frozen=True)
class PRNGKey:
    seed: int

    def __post_init__(self):
        object.__setattr__(self, "seed", int(self.seed) & _MASK)


KeyLike = Union[PRNGKey, int]


def _as_key(key: KeyLike) -> PRNGKey:
    return key if isinstance(key, PRNGKey) else PRNGKey(int(key))


def _mix(x: int) -> int:  # splitmix64 finaliser
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


def split(key: KeyLike, num: int = 2) -> Tuple[PRNGKey, ...]:
    k = _as_key(key)
    return tuple(PRNGKey(_mix(k.seed * 0x100000001B3 + i + 1)) for i in range(num))


def uniform(key: KeyLike, shape, device) -> torch.Tensor:
    """fp32 uniforms k * 2^-23, k in [0, 2^23), generated on `device`."""
    gen = torch.Generator(device=device)
    gen.manual_seed(_as_key(key).seed)
    bits = torch.randint(0, 1 << 23, tuple(shape), generator=gen, device=device, dtype=torch.int32)
    return bits.to(torch.float32) * (2.0 ** -23)
