"""Deterministic fp32 exp shared bit-for-bit between oracle and CUDA kernels.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference calls ``jnp.exp`` at render.py:279 and render.py:284.  XLA's exp
is a polynomial whose last-ulp behaviour is neither documented nor available
here, and CUDA's ``expf`` differs from numpy's in the last ulp, which would
break bit-exact fine-sample positions (north star).  So both sides restate exp
as the SAME sequence of individually rounded fp32 multiplies and adds (no FMA):
Cody-Waite reduction by ln2 (hi/lo split) + the degree-5 Cephes ``expf``
polynomial + exact scaling by 2^n.  Measured max error vs correctly rounded
exp on [-87, 0]: <= 1 ulp (tests/test_oracle.py::test_expf_accuracy).

The CUDA twin is ``lnrf_expf`` in learn-nerf_b200/csrc/lnrf_math.cuh and must be
kept in lock-step with this file.
"""
import numpy as np

F = np.float32

LOG2E = F(1.44269504088896341)
LN2_HI = F(0.693359375)            # 0x3f318000: 8 trailing zero bits -> n*LN2_HI exact for |n| < 2^15
LN2_LO = F(-2.12194440e-4)         # ln2 - LN2_HI
P0 = F(1.9875691500e-4)
P1 = F(1.3981999507e-3)
P2 = F(8.3334519073e-3)
P3 = F(4.1665795894e-2)
P4 = F(1.6666665459e-1)
P5 = F(5.0000001201e-1)
X_MIN = F(-87.0)                   # below this the result is defined as +0
X_MAX = F(88.0)                    # above this the result is defined as +inf


def expf(x):
    """exp(x) in fp32 by the shared deterministic algorithm (array or scalar)."""
    x = np.asarray(x, dtype=F)
    with np.errstate(over="ignore", invalid="ignore"):
        xc = np.minimum(np.maximum(x, X_MIN), X_MAX)
        n = np.rint(xc * LOG2E).astype(F)
        r = (xc - n * LN2_HI).astype(F)
        r = (r - n * LN2_LO).astype(F)
        r2 = (r * r).astype(F)
        p = (P0 * r + P1).astype(F)
        p = (p * r + P2).astype(F)
        p = (p * r + P3).astype(F)
        p = (p * r + P4).astype(F)
        p = (p * r + P5).astype(F)
        p = (p * r2).astype(F)
        p = (p + r).astype(F)
        p = (p + F(1.0)).astype(F)
        # scale by 2^n exactly; n in [-126, 127] after the clamp
        bits = ((n.astype(np.int32) + 127) << 23).astype(np.int32)
        scale = bits.view(F)
        out = (p * scale).astype(F)
        out = np.where(x < X_MIN, F(0.0), out)
        out = np.where(x > X_MAX, F(np.inf), out)
        out = np.where(np.isnan(x), F(np.nan), out)
    return out.astype(F)
