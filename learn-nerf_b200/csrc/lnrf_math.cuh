// Device math shared by the renderer kernels.
#pragma once
#include <cuda_runtime.h>

namespace lnrf {

// Deterministic fp32 exp: the CUDA twin of oracle/expf.py (keep in lock-step).
// Only individually rounded multiplies/adds (no FMA contraction), so the numpy
// oracle reproduces it bit for bit.  Replaces jnp.exp at render.py:279,284.
__device__ __forceinline__ float lnrf_expf(float x) {
  const float LOG2E = 1.44269504088896341f;
  const float LN2_HI = 0.693359375f;
  const float LN2_LO = -2.12194440e-4f;
  if (x != x) return x;
  if (x < -87.0f) return 0.0f;
  if (x > 88.0f) return __int_as_float(0x7f800000);
  float n = rintf(__fmul_rn(x, LOG2E));
  float r = __fsub_rn(x, __fmul_rn(n, LN2_HI));
  r = __fsub_rn(r, __fmul_rn(n, LN2_LO));
  float r2 = __fmul_rn(r, r);
  float p = __fadd_rn(__fmul_rn(1.9875691500e-4f, r), 1.3981999507e-3f);
  p = __fadd_rn(__fmul_rn(p, r), 8.3334519073e-3f);
  p = __fadd_rn(__fmul_rn(p, r), 4.1665795894e-2f);
  p = __fadd_rn(__fmul_rn(p, r), 1.6666665459e-1f);
  p = __fadd_rn(__fmul_rn(p, r), 5.0000001201e-1f);
  p = __fmul_rn(p, r2);
  p = __fadd_rn(p, r);
  p = __fadd_rn(p, 1.0f);
  float scale = __int_as_float((static_cast<int>(n) + 127) << 23);
  return __fmul_rn(p, scale);
}

__device__ __forceinline__ float softplus_f(float x) {
  // logaddexp(x, 0) as flax nn.softplus (model.py:57)
  return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace lnrf
