// sinusoidal_emb (learn_nerf/model.py:65-77) as a standalone kernel for the fp32 paths.
#pragma once
#include "lnrf_common.cuh"

namespace lnrf {

// Per coordinate: [sin 2^0..2^{F-1}, cos 2^0..2^{F-1}].  One thread per (sample, dim, freq).
// Ray mode (v == nullptr) forms the point as o + d*t with two roundings (render.py:153);
// `which` = 0 embeds the position, 1 the ray direction.
template <int FREQS>
__global__ void __launch_bounds__(256)
embed_kernel(const float* __restrict__ v, const float* __restrict__ rays, const float* __restrict__ ts,
             int T, int which, int64_t m, float* __restrict__ out) {
  const int64_t total = m * 3 * FREQS;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t s = i / (3 * FREQS);
    const int rem = int(i - s * 3 * FREQS);
    const int dim = rem / FREQS, f = rem - dim * FREQS;
    float c;
    if (v) {
      c = __ldg(v + s * 3 + dim);
    } else {
      const int64_t r = s / T;
      const float dd = __ldg(rays + r * 6 + 3 + dim);
      c = which ? dd : __fadd_rn(__ldg(rays + r * 6 + dim), __fmul_rn(dd, __ldg(ts + s)));
    }
    const float a = c * float(1 << f);
    float sn, cs;
    sincosf(a, &sn, &cs);
    float* o = out + s * (6 * FREQS) + dim * 2 * FREQS;
    o[f] = sn;
    o[FREQS + f] = cs;
  }
}

static inline unsigned ew_blocks(int64_t work_items, int per_block) {
  int64_t b = ceil_div(work_items, per_block);
  int64_t cap = int64_t(sm_count()) * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace lnrf
