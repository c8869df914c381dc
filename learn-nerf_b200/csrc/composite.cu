// K3 / K5: alpha compositing (transmittance scan) and its reverse-scan gradient,
// plus the MSE loss head.  One warp per ray; each lane owns a contiguous chunk of
// the ray's samples, chunk totals are combined with warp-shuffle scans; all [n,T]
// traffic goes through lane-contiguous (coalesced) loads/stores staged in smem.
//
// Reference: RaySamples.termination_probs / render_rays / render_alpha and the
// coords render, learn_nerf/render.py:155-190, 270-287, 329-331.
#include "lnrf_common.cuh"
#include "lnrf_math.cuh"

namespace lnrf {

// exp of the compositing kernels: the hardware ex2 (MUFU.EX2, ~2 ulp).  Compositing is compared by tolerance
// (1e-5 abs on outputs / alphas / coords), so it does not need the bit-exact 25-instruction lnrf_expf of the
// sampling kernels (whose weights decide sample POSITIONS and must match the oracle bit for bit): with two to
// three exponentials per sample that polynomial made K3 / K5 issue-bound at 0.3 - 0.5 of the HBM roofline.
__device__ __forceinline__ float comp_expf(float x) { return __expf(x); }

constexpr int kCompWarps = 4;

// exclusive prefix sum of one value per lane; `total` = sum over the warp
__device__ __forceinline__ float warp_excl_scan(float v, int lane, float& total) {
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

// exclusive suffix sum: sum of the values held by HIGHER lanes
__device__ __forceinline__ float warp_excl_suffix(float v, int lane) {
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(0xffffffffu, inc, o);
    if (lane + o < 32) inc += t;
  }
  return inc - v;
}

// With ts in s_ts and the densities in s_a: fills delta / a = dens*delta / acc_prev.
// Returns the ray's total optical depth.
__device__ __forceinline__ float stage_ray_smem(int T, float t_min, float t_max, int lane, int c0, int c1,
                                                float* s_ts, float* s_delta, float* s_a, float* s_acc);

// Loads one ray's samples into smem and fills delta / a = dens*delta / acc_prev.
// Returns the ray's total optical depth.
__device__ __forceinline__ float stage_ray(const float* __restrict__ ts, const float* __restrict__ dens,
                                           int64_t r, int T, float t_min, float t_max, int lane,
                                           int c0, int c1, float* s_ts, float* s_delta, float* s_a,
                                           float* s_acc) {
  for (int i = lane; i < T; i += 32) {
    s_ts[i] = __ldg(ts + r * T + i);
    s_a[i] = __ldg(dens + r * T + i);
  }
  __syncwarp();
  return stage_ray_smem(T, t_min, t_max, lane, c0, c1, s_ts, s_delta, s_a, s_acc);
}

__device__ __forceinline__ float stage_ray_smem(int T, float t_min, float t_max, int lane, int c0, int c1,
                                                float* s_ts, float* s_delta, float* s_a, float* s_acc) {
  float local = 0.0f;
  for (int i = c0; i < c1; ++i) {  // starts/ends/deltas render.py:259-268, density_dt :271
    float t = s_ts[i];
    float start = (i == 0) ? t_min : (t + s_ts[i - 1]) * 0.5f;
    float end = (i == T - 1) ? t_max : (s_ts[i + 1] + t) * 0.5f;
    float delta = end - start;
    float a = s_a[i] * delta;
    s_delta[i] = delta;
    s_acc[i] = local;  // chunk-local exclusive prefix for now
    local += a;
    s_a[i] = a;
  }
  float total;
  float base = warp_excl_scan(local, lane, total);
  for (int i = c0; i < c1; ++i) s_acc[i] += base;  // acc_densities_prev, :275-278
  return total;
}

__global__ void __launch_bounds__(kCompWarps * 32)
composite_fwd_kernel(const float* __restrict__ rays, const float* __restrict__ ts,
                     const float* __restrict__ t_min_in, const float* __restrict__ t_max_in,
                     const uint8_t* __restrict__ mask_in, const float* __restrict__ dens,
                     const float* __restrict__ rgb, const float* __restrict__ background, int64_t n,
                     int T, float* __restrict__ outputs, float* __restrict__ alphas,
                     float* __restrict__ coords) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* s_ts = smem + size_t(wib) * 7 * T;
  float* s_delta = s_ts + T;
  float* s_a = s_delta + T;
  float* s_acc = s_a + T;
  float* s_rgb = s_acc + T;  // 3T
  const int C = (T + 31) / 32;
  const int c0 = min(lane * C, T), c1 = min(c0 + C, T);
  const float bg0 = __ldg(background), bg1 = __ldg(background + 1), bg2 = __ldg(background + 2);
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += nwarps) {
    if (!mask_in[r]) {  // render.py:174-176, :190: masked rays show the background
      if (lane < 3) outputs[r * 3 + lane] = lane == 0 ? bg0 : (lane == 1 ? bg1 : bg2);
      if (lane == 0 && alphas) alphas[r] = 0.0f;
      if (lane < 3 && coords) coords[r * 3 + lane] = 0.0f;
      continue;
    }
    for (int i = lane; i < 3 * T; i += 32) s_rgb[i] = __ldg(rgb + r * 3 * T + i);
    float total = stage_ray(ts, dens, r, T, __ldg(t_min_in + r), __ldg(t_max_in + r), lane, c0, c1,
                            s_ts, s_delta, s_a, s_acc);
    __syncwarp();
    float o[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o[k] = __ldg(rays + r * 6 + k);
      d[k] = __ldg(rays + r * 6 + 3 + k);
    }
    float c_rgb[3] = {0.f, 0.f, 0.f}, c_xyz[3] = {0.f, 0.f, 0.f};
    for (int i = c0; i < c1; ++i) {
      float p = comp_expf(-s_acc[i]) * (1.0f - comp_expf(-s_a[i]));  // :279-287
      float t = s_ts[i];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        c_rgb[k] += p * s_rgb[i * 3 + k];
        c_xyz[k] += p * __fadd_rn(o[k], __fmul_rn(d[k], t));  // points, :153
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      c_rgb[k] = warp_sum(c_rgb[k]);
      c_xyz[k] = warp_sum(c_xyz[k]);
    }
    const float p_esc = comp_expf(-total);  // last column of termination_probs
    if (lane == 0) {
      outputs[r * 3 + 0] = c_rgb[0] + p_esc * bg0;
      outputs[r * 3 + 1] = c_rgb[1] + p_esc * bg1;
      outputs[r * 3 + 2] = c_rgb[2] + p_esc * bg2;
      if (alphas) alphas[r] = 1.0f - p_esc;
      if (coords) {
        coords[r * 3 + 0] = c_xyz[0];
        coords[r * 3 + 1] = c_xyz[1];
        coords[r * 3 + 2] = c_xyz[2];
      }
    }
    __syncwarp();
  }
}

// Register-prefetching variants for T <= 32 C (C = 2: the 64 coarse samples, C = 6: the 192 fine
// samples): the next ray's ts / densities / colours and per-ray scalars are requested before the
// current ray is processed, so a warp always has one ray of loads (1.3 / 3.9 KB) in flight behind
// its arithmetic.  The generic kernels above were latency-bound (25 - 45 % of the HBM copy peak).
template <int C>
struct RayRegs {
  float ts[C], dens[C], rgb[3 * C];
  float tmin, tmax;
  int mask;
};
template <int C>
__device__ __forceinline__ void ray_prefetch(RayRegs<C>& q, const float* __restrict__ ts, const float* __restrict__ dens,
                                             const float* __restrict__ rgb, const float* __restrict__ t_min_in,
                                             const float* __restrict__ t_max_in, const uint8_t* __restrict__ mask_in,
                                             int64_t r, int T, int lane) {
  q.mask = mask_in[r];
  q.tmin = __ldg(t_min_in + r);
  q.tmax = __ldg(t_max_in + r);
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int i = lane + 32 * j;
    q.ts[j] = i < T ? __ldg(ts + r * T + i) : 0.0f;
    q.dens[j] = i < T ? __ldg(dens + r * T + i) : 0.0f;
  }
#pragma unroll
  for (int j = 0; j < 3 * C; ++j) {
    const int i = lane + 32 * j;
    q.rgb[j] = i < 3 * T ? __ldg(rgb + r * 3 * T + i) : 0.0f;
  }
}
template <int C>
__device__ __forceinline__ void ray_to_smem(const RayRegs<C>& q, int T, int lane, float* s_ts, float* s_a, float* s_rgb) {
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int i = lane + 32 * j;
    if (i < T) {
      s_ts[i] = q.ts[j];
      s_a[i] = q.dens[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 3 * C; ++j) {
    const int i = lane + 32 * j;
    if (i < 3 * T) s_rgb[i] = q.rgb[j];
  }
}

template <int C>
__global__ void __launch_bounds__(kCompWarps * 32)
composite_fwd_pf_kernel(const float* __restrict__ rays, const float* __restrict__ ts,
                        const float* __restrict__ t_min_in, const float* __restrict__ t_max_in,
                        const uint8_t* __restrict__ mask_in, const float* __restrict__ dens,
                        const float* __restrict__ rgb, const float* __restrict__ background, int64_t n,
                        int T, float* __restrict__ outputs, float* __restrict__ alphas,
                        float* __restrict__ coords) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* s_ts = smem + size_t(wib) * 7 * T;
  float* s_delta = s_ts + T;
  float* s_a = s_delta + T;
  float* s_acc = s_a + T;
  float* s_rgb = s_acc + T;  // 3T
  const int CH = (T + 31) / 32;
  const int c0 = min(lane * CH, T), c1 = min(c0 + CH, T);
  const float bg0 = __ldg(background), bg1 = __ldg(background + 1), bg2 = __ldg(background + 2);
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  RayRegs<C> q;
  float od[6];
  if (warp < n) {
    ray_prefetch<C>(q, ts, dens, rgb, t_min_in, t_max_in, mask_in, warp, T, lane);
#pragma unroll
    for (int k = 0; k < 6; ++k) od[k] = __ldg(rays + warp * 6 + k);
  }
  for (int64_t r = warp; r < n; r += nwarps) {
    const int mask = q.mask;
    const float t_min = q.tmin, t_max = q.tmax;
    float o[3] = {od[0], od[1], od[2]}, d[3] = {od[3], od[4], od[5]};
    if (mask) ray_to_smem<C>(q, T, lane, s_ts, s_a, s_rgb);
    if (r + nwarps < n) {  // next ray's loads go out before this ray's arithmetic
      ray_prefetch<C>(q, ts, dens, rgb, t_min_in, t_max_in, mask_in, r + nwarps, T, lane);
#pragma unroll
      for (int k = 0; k < 6; ++k) od[k] = __ldg(rays + (r + nwarps) * 6 + k);
    }
    if (!mask) {  // render.py:174-176, :190: masked rays show the background
      if (lane < 3) outputs[r * 3 + lane] = lane == 0 ? bg0 : (lane == 1 ? bg1 : bg2);
      if (lane == 0 && alphas) alphas[r] = 0.0f;
      if (lane < 3 && coords) coords[r * 3 + lane] = 0.0f;
      continue;
    }
    __syncwarp();
    float total = stage_ray_smem(T, t_min, t_max, lane, c0, c1, s_ts, s_delta, s_a, s_acc);
    __syncwarp();
    float c_rgb[3] = {0.f, 0.f, 0.f}, c_xyz[3] = {0.f, 0.f, 0.f};
    for (int i = c0; i < c1; ++i) {
      float p = comp_expf(-s_acc[i]) * (1.0f - comp_expf(-s_a[i]));  // :279-287
      float t = s_ts[i];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        c_rgb[k] += p * s_rgb[i * 3 + k];
        c_xyz[k] += p * __fadd_rn(o[k], __fmul_rn(d[k], t));  // points, :153
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      c_rgb[k] = warp_sum(c_rgb[k]);
      c_xyz[k] = warp_sum(c_xyz[k]);
    }
    const float p_esc = comp_expf(-total);  // last column of termination_probs
    if (lane == 0) {
      outputs[r * 3 + 0] = c_rgb[0] + p_esc * bg0;
      outputs[r * 3 + 1] = c_rgb[1] + p_esc * bg1;
      outputs[r * 3 + 2] = c_rgb[2] + p_esc * bg2;
      if (alphas) alphas[r] = 1.0f - p_esc;
      if (coords) {
        coords[r * 3 + 0] = c_xyz[0];
        coords[r * 3 + 1] = c_xyz[1];
        coords[r * 3 + 2] = c_xyz[2];
      }
    }
    __syncwarp();
  }
}

// Gradient: with a_k = dens_k*delta_k, T_k = exp(-sum_{j<k} a_j),
// p_k = T_k (1 - exp(-a_k)), g_k = rgb_k . dO, g_bg = bg . dO:
//   dL/drgb_k = p_k dO,   dL/dbg += p_esc dO,
//   dL/da_k   = T_{k+1} g_k - (sum_{i>k} p_i g_i + p_esc g_bg)      (suffix scan)
//   dL/ddens_k = dL/da_k * delta_k.
// Masked rays: outputs == background, so dL/dbg += dO and the rest is zero.
__global__ void __launch_bounds__(kCompWarps * 32)
composite_bwd_kernel(const float* __restrict__ ts, const float* __restrict__ t_min_in,
                     const float* __restrict__ t_max_in, const uint8_t* __restrict__ mask_in,
                     const float* __restrict__ dens, const float* __restrict__ rgb,
                     const float* __restrict__ background, const float* __restrict__ d_outputs,
                     int64_t n, int T, float* __restrict__ d_dens, float* __restrict__ d_rgb,
                     float* __restrict__ d_background) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* s_ts = smem + size_t(wib) * 7 * T;
  float* s_delta = s_ts + T;
  float* s_a = s_delta + T;
  float* s_acc = s_a + T;
  float* s_rgb = s_acc + T;  // 3T
  const int C = (T + 31) / 32;
  const int c0 = min(lane * C, T), c1 = min(c0 + C, T);
  const float bg0 = __ldg(background), bg1 = __ldg(background + 1), bg2 = __ldg(background + 2);
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float dbg[3] = {0.f, 0.f, 0.f};  // lane 0 accumulates this warp's d_background
  for (int64_t r = warp; r < n; r += nwarps) {
    const float g0 = __ldg(d_outputs + r * 3), g1 = __ldg(d_outputs + r * 3 + 1),
                g2 = __ldg(d_outputs + r * 3 + 2);
    if (!mask_in[r]) {
      for (int i = lane; i < T; i += 32) d_dens[r * T + i] = 0.0f;
      for (int i = lane; i < 3 * T; i += 32) d_rgb[r * 3 * T + i] = 0.0f;
      dbg[0] += g0; dbg[1] += g1; dbg[2] += g2;
      continue;
    }
    for (int i = lane; i < 3 * T; i += 32) s_rgb[i] = __ldg(rgb + r * 3 * T + i);
    float total = stage_ray(ts, dens, r, T, __ldg(t_min_in + r), __ldg(t_max_in + r), lane, c0, c1,
                            s_ts, s_delta, s_a, s_acc);
    __syncwarp();
    const float p_esc = comp_expf(-total);
    const float g_bg = bg0 * g0 + bg1 * g1 + bg2 * g2;
    // forward sweep over the chunk: p_k, p_k*g_k; stash p in s_ts (ts no longer needed)
    float local = 0.0f;
    for (int i = c0; i < c1; ++i) {
      float p = comp_expf(-s_acc[i]) * (1.0f - comp_expf(-s_a[i]));
      float gk = s_rgb[i * 3] * g0 + s_rgb[i * 3 + 1] * g1 + s_rgb[i * 3 + 2] * g2;
      s_ts[i] = p;
      local += p * gk;
    }
    float suffix = warp_excl_suffix(local, lane) + p_esc * g_bg;  // sum over later lanes + bg term
    for (int i = c1 - 1; i >= c0; --i) {
      float p = s_ts[i];
      float gk = s_rgb[i * 3] * g0 + s_rgb[i * 3 + 1] * g1 + s_rgb[i * 3 + 2] * g2;
      float t_next = comp_expf(-(s_acc[i] + s_a[i]));  // T_{k+1}
      float dla = t_next * gk - suffix;
      suffix += p * gk;
      s_a[i] = dla * s_delta[i];  // d_dens
      s_rgb[i * 3] = p * g0;      // d_rgb
      s_rgb[i * 3 + 1] = p * g1;
      s_rgb[i * 3 + 2] = p * g2;
    }
    __syncwarp();
    for (int i = lane; i < T; i += 32) d_dens[r * T + i] = s_a[i];
    for (int i = lane; i < 3 * T; i += 32) d_rgb[r * 3 * T + i] = s_rgb[i];
    dbg[0] += p_esc * g0; dbg[1] += p_esc * g1; dbg[2] += p_esc * g2;
    __syncwarp();
  }
  if (lane == 0 && (dbg[0] != 0.f || dbg[1] != 0.f || dbg[2] != 0.f)) {
    atomicAdd(d_background + 0, dbg[0]);
    atomicAdd(d_background + 1, dbg[1]);
    atomicAdd(d_background + 2, dbg[2]);
  }
}

template <int C>
__global__ void __launch_bounds__(kCompWarps * 32)
composite_bwd_pf_kernel(const float* __restrict__ ts, const float* __restrict__ t_min_in,
                        const float* __restrict__ t_max_in, const uint8_t* __restrict__ mask_in,
                        const float* __restrict__ dens, const float* __restrict__ rgb,
                        const float* __restrict__ background, const float* __restrict__ d_outputs,
                        int64_t n, int T, float* __restrict__ d_dens, float* __restrict__ d_rgb,
                        float* __restrict__ d_background) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* s_ts = smem + size_t(wib) * 7 * T;
  float* s_delta = s_ts + T;
  float* s_a = s_delta + T;
  float* s_acc = s_a + T;
  float* s_rgb = s_acc + T;  // 3T
  const int CH = (T + 31) / 32;
  const int c0 = min(lane * CH, T), c1 = min(c0 + CH, T);
  const float bg0 = __ldg(background), bg1 = __ldg(background + 1), bg2 = __ldg(background + 2);
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float dbg[3] = {0.f, 0.f, 0.f};  // lane 0 accumulates this warp's d_background
  RayRegs<C> q;
  float gq[3];
  if (warp < n) {
    ray_prefetch<C>(q, ts, dens, rgb, t_min_in, t_max_in, mask_in, warp, T, lane);
#pragma unroll
    for (int k = 0; k < 3; ++k) gq[k] = __ldg(d_outputs + warp * 3 + k);
  }
  for (int64_t r = warp; r < n; r += nwarps) {
    const int mask = q.mask;
    const float t_min = q.tmin, t_max = q.tmax;
    const float g0 = gq[0], g1 = gq[1], g2 = gq[2];
    if (mask) ray_to_smem<C>(q, T, lane, s_ts, s_a, s_rgb);
    if (r + nwarps < n) {
      ray_prefetch<C>(q, ts, dens, rgb, t_min_in, t_max_in, mask_in, r + nwarps, T, lane);
#pragma unroll
      for (int k = 0; k < 3; ++k) gq[k] = __ldg(d_outputs + (r + nwarps) * 3 + k);
    }
    if (!mask) {
      for (int i = lane; i < T; i += 32) d_dens[r * T + i] = 0.0f;
      for (int i = lane; i < 3 * T; i += 32) d_rgb[r * 3 * T + i] = 0.0f;
      dbg[0] += g0; dbg[1] += g1; dbg[2] += g2;
      continue;
    }
    __syncwarp();
    float total = stage_ray_smem(T, t_min, t_max, lane, c0, c1, s_ts, s_delta, s_a, s_acc);
    __syncwarp();
    const float p_esc = comp_expf(-total);
    const float g_bg = bg0 * g0 + bg1 * g1 + bg2 * g2;
    float local = 0.0f;
    for (int i = c0; i < c1; ++i) {
      float p = comp_expf(-s_acc[i]) * (1.0f - comp_expf(-s_a[i]));
      float gk = s_rgb[i * 3] * g0 + s_rgb[i * 3 + 1] * g1 + s_rgb[i * 3 + 2] * g2;
      s_ts[i] = p;
      local += p * gk;
    }
    float suffix = warp_excl_suffix(local, lane) + p_esc * g_bg;  // sum over later lanes + bg term
    for (int i = c1 - 1; i >= c0; --i) {
      float p = s_ts[i];
      float gk = s_rgb[i * 3] * g0 + s_rgb[i * 3 + 1] * g1 + s_rgb[i * 3 + 2] * g2;
      float t_next = comp_expf(-(s_acc[i] + s_a[i]));  // T_{k+1}
      float dla = t_next * gk - suffix;
      suffix += p * gk;
      s_a[i] = dla * s_delta[i];  // d_dens
      s_rgb[i * 3] = p * g0;      // d_rgb
      s_rgb[i * 3 + 1] = p * g1;
      s_rgb[i * 3 + 2] = p * g2;
    }
    __syncwarp();
    for (int i = lane; i < T; i += 32) d_dens[r * T + i] = s_a[i];
    for (int i = lane; i < 3 * T; i += 32) d_rgb[r * 3 * T + i] = s_rgb[i];
    dbg[0] += p_esc * g0; dbg[1] += p_esc * g1; dbg[2] += p_esc * g2;
    __syncwarp();
  }
  if (lane == 0 && (dbg[0] != 0.f || dbg[1] != 0.f || dbg[2] != 0.f)) {
    atomicAdd(d_background + 0, dbg[0]);
    atomicAdd(d_background + 1, dbg[1]);
    atomicAdd(d_background + 2, dbg[2]);
  }
}

// train.py:140-142: per-level MSE and its gradient.
__global__ void __launch_bounds__(256)
mse_loss_kernel(const float* __restrict__ outputs, const float* __restrict__ targets,
                int64_t target_stride, int64_t n, float inv_count, float* __restrict__ loss_sum,
                float* __restrict__ d_outputs) {
  float acc = 0.0f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n * 3;
       i += int64_t(gridDim.x) * blockDim.x) {
    int64_t r = i / 3;
    int c = int(i - r * 3);
    float diff = outputs[i] - __ldg(targets + r * target_stride + c);
    acc += diff * diff;
    if (d_outputs) d_outputs[i] = 2.0f * diff * inv_count;
  }
  acc = warp_sum(acc);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += s[w];
    atomicAdd(loss_sum, t);
  }
}

static int comp_launch_dims(int64_t n, int T, int64_t& blocks, size_t& smem) {
  smem = size_t(kCompWarps) * 7 * T * sizeof(float);
  blocks = ceil_div(n, kCompWarps);
  int64_t cap = int64_t(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  return 0;
}

}  // namespace lnrf

extern "C" {

int lnrf_composite_fwd(const float* rays, const float* ts, const float* t_min, const float* t_max,
                       const uint8_t* mask, const float* dens, const float* rgb,
                       const float* background, int64_t n, int32_t T, float* outputs, float* alphas,
                       float* coords, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && T > 0, LNRF_E_INVALID, "lnrf_composite_fwd: n=%lld T=%d", (long long)n, T);
  LNRF_REQUIRE(T <= 1024, LNRF_E_UNSUPPORTED, "lnrf_composite_fwd: T=%d > 1024", T);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(rays && ts && t_min && t_max && mask && dens && rgb && background && outputs,
               LNRF_E_INVALID, "lnrf_composite_fwd: null pointer");
  int64_t blocks;
  size_t smem;
  lnrf::comp_launch_dims(n, T, blocks, smem);
  if (smem > 48 * 1024)  // the attribute is per device: set it per call (cheap, idempotent), no process state
    LNRF_CUDA(cudaFuncSetAttribute(lnrf::composite_fwd_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (T <= 64)
    lnrf::composite_fwd_pf_kernel<2><<<(unsigned)blocks, lnrf::kCompWarps * 32, smem, lnrf::as_stream(stream)>>>(
        rays, ts, t_min, t_max, mask, dens, rgb, background, n, T, outputs, alphas, coords);
  else if (T <= 192)
    lnrf::composite_fwd_pf_kernel<6><<<(unsigned)blocks, lnrf::kCompWarps * 32, smem, lnrf::as_stream(stream)>>>(
        rays, ts, t_min, t_max, mask, dens, rgb, background, n, T, outputs, alphas, coords);
  else
    lnrf::composite_fwd_kernel<<<(unsigned)blocks, lnrf::kCompWarps * 32, smem,
                                 lnrf::as_stream(stream)>>>(rays, ts, t_min, t_max, mask, dens, rgb,
                                                            background, n, T, outputs, alphas, coords);
  LNRF_LAUNCH_CHECK("composite_fwd_kernel");
  return LNRF_OK;
}

int lnrf_composite_bwd(const float* ts, const float* t_min, const float* t_max, const uint8_t* mask,
                       const float* dens, const float* rgb, const float* background,
                       const float* d_outputs, int64_t n, int32_t T, float* d_dens, float* d_rgb,
                       float* d_background, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && T > 0, LNRF_E_INVALID, "lnrf_composite_bwd: n=%lld T=%d", (long long)n, T);
  LNRF_REQUIRE(T <= 1024, LNRF_E_UNSUPPORTED, "lnrf_composite_bwd: T=%d > 1024", T);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(ts && t_min && t_max && mask && dens && rgb && background && d_outputs && d_dens &&
                   d_rgb && d_background,
               LNRF_E_INVALID, "lnrf_composite_bwd: null pointer");
  int64_t blocks;
  size_t smem;
  lnrf::comp_launch_dims(n, T, blocks, smem);
  if (smem > 48 * 1024)  // the attribute is per device: set it per call (cheap, idempotent), no process state
    LNRF_CUDA(cudaFuncSetAttribute(lnrf::composite_bwd_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (T <= 64)
    lnrf::composite_bwd_pf_kernel<2><<<(unsigned)blocks, lnrf::kCompWarps * 32, smem, lnrf::as_stream(stream)>>>(
        ts, t_min, t_max, mask, dens, rgb, background, d_outputs, n, T, d_dens, d_rgb, d_background);
  else if (T <= 192)
    lnrf::composite_bwd_pf_kernel<6><<<(unsigned)blocks, lnrf::kCompWarps * 32, smem, lnrf::as_stream(stream)>>>(
        ts, t_min, t_max, mask, dens, rgb, background, d_outputs, n, T, d_dens, d_rgb, d_background);
  else
    lnrf::composite_bwd_kernel<<<(unsigned)blocks, lnrf::kCompWarps * 32, smem,
                                 lnrf::as_stream(stream)>>>(ts, t_min, t_max, mask, dens, rgb,
                                                            background, d_outputs, n, T, d_dens, d_rgb,
                                                            d_background);
  LNRF_LAUNCH_CHECK("composite_bwd_kernel");
  return LNRF_OK;
}

int lnrf_mse_loss(const float* outputs, const float* targets, int64_t target_stride, int64_t n,
                  float inv_count, float* loss_sum, float* d_outputs, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && target_stride >= 3, LNRF_E_INVALID, "lnrf_mse_loss: n=%lld stride=%lld",
               (long long)n, (long long)target_stride);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(outputs && targets && loss_sum, LNRF_E_INVALID, "lnrf_mse_loss: null pointer");
  int64_t blocks = lnrf::ceil_div(n * 3, 256);
  if (blocks > 1024) blocks = 1024;
  lnrf::mse_loss_kernel<<<(unsigned)blocks, 256, 0, lnrf::as_stream(stream)>>>(
      outputs, targets, target_stride, n, inv_count, loss_sum, d_outputs);
  LNRF_LAUNCH_CHECK("mse_loss_kernel");
  return LNRF_OK;
}

}  // extern "C"
