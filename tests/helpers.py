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


def reproducible(fn, tries=4):
    """Evaluate the CPU oracle `fn()` until two consecutive evaluations agree bit for bit, and return that result.

    Why: on the GPU boxes' hosts the multi-threaded torch CPU matmul has been caught returning a slightly different
    result for one thread's block of rows (rows 512..767 of 4097, 6e-5 abs in an O(1) output) in ONE evaluation of
    the fp32 oracle -- the next evaluation of the same expression on the same inputs, and every GPU evaluation, were
    bit-identical (captured in a fresh-box run of test_ngp_model_apply[6]).  The 1e-5 comparisons must not depend on
    that, so they take an oracle value that reproduced."""
    import torch

    def leaves(t):
        if isinstance(t, dict):
            return [x for k in sorted(t) for x in leaves(t[k])]
        if isinstance(t, (tuple, list)):
            return [x for v in t for x in leaves(v)]
        return [t]

    def same(a, b):
        if isinstance(a, torch.Tensor):
            return torch.equal(a.detach(), b.detach())
        return bool(np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True))

    prev = fn()
    for _ in range(tries):
        cur = fn()
        if all(same(a, b) for a, b in zip(leaves(prev), leaves(cur))):
            return cur
        prev = cur
    return prev
