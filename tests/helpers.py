"""Shared synthetic inputs for oracle and GPU parity tests."""
import numpy as np

F = np.float32
BBOX_MIN = np.array([-1.0, -1.0, -1.0], F)
BBOX_MAX = np.array([1.0, 1.0, 1.0], F)


def make_rays(n, seed=0, miss_frac=0.0, with_targets=True):
    """Origins on the radius-4 sphere, directions toward a uniform point of the bbox
    (or away from it for a `miss_frac` share), targets U(-1,1)."""
    rs = np.random.RandomState(seed)
    o = rs.randn(n, 3)
    o = 4.0 * o / np.linalg.norm(o, axis=1, keepdims=True)
    tgt = rs.uniform(-1, 1, (n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    miss = rs.uniform(size=n) < miss_frac
    d[miss] = -d[miss]
    cols = [o, d]
    if with_targets:
        cols.append(rs.uniform(-1, 1, (n, 3)))
    return np.stack(cols, axis=1).astype(F)


def make_uniforms(n, t, seed=1):
    rs = np.random.RandomState(seed)
    return (rs.randint(0, 2 ** 23, (n, t)).astype(np.float64) * 2.0 ** -23).astype(F)
