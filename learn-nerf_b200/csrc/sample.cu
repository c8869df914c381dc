// K1 (bbox clip + stratified coarse sampling) and K4 (inverse-CDF fine sampling).
//
// Both are bit-exact twins of oracle/render_np.py: every reference op is one
// individually rounded fp32 op (__fadd_rn/__fmul_rn/__fdiv_rn forbid FMA
// contraction), cumulative sums run strictly left to right, exp is lnrf_expf.
// A warp owns one ray; loads/stores of the [n,T] arrays are lane-contiguous.
#include <math_constants.h>

#include "lnrf_common.cuh"
#include "lnrf_math.cuh"

namespace lnrf {

// np.minimum / np.maximum propagate NaN; fminf/fmaxf do not.
__device__ __forceinline__ float np_min(float a, float b) {
  return (a != a || b != b) ? CUDART_NAN_F : fminf(a, b);
}
__device__ __forceinline__ float np_max(float a, float b) {
  return (a != a || b != b) ? CUDART_NAN_F : fmaxf(a, b);
}

struct BBox {
  float lo[3];
  float hi[3];
};

// ray_t_range, render.py:346-389
__device__ __forceinline__ void ray_t_range(const float* __restrict__ ray, const BBox& bb,
                                            float min_t_range, float epsilon, float& t_min,
                                            float& t_max, bool& mask) {
  float lo_max = 0.0f, hi_min = 0.0f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float o = __ldg(ray + a), d = __ldg(ray + 3 + a);
    float denom = __fadd_rn(d, epsilon);                    // :369
    float t0 = __fdiv_rn(__fsub_rn(bb.lo[a], o), denom);    // :368-369
    float t1 = __fdiv_rn(__fsub_rn(bb.hi[a], o), denom);
    float lo = np_min(t0, t1), hi = np_max(t0, t1);         // :372-378
    lo_max = (a == 0) ? lo : np_max(lo_max, lo);
    hi_min = (a == 0) ? hi : np_min(hi_min, hi);
  }
  float min_t = np_max(0.0f, lo_max);                        // :381
  float max_t = hi_min;                                      // :382
  float max_t_clipped = np_max(max_t, __fadd_rn(min_t, min_t_range));  // :383
  mask = min_t < max_t;                                      // :386
  t_min = mask ? min_t : 0.0f;                               // :387
  t_max = mask ? max_t_clipped : min_t_range;
}

// One thread per 4 consecutive samples of a ray (T % 4 == 0) or per sample: flat, fully coalesced
// 128-bit loads/stores of u / ts.  The slab test is recomputed by every thread of a ray (six
// divisions, far cheaper than staging it); the thread of sample 0 writes the ray's bounds.
template <int V>
__global__ void __launch_bounds__(256)
sample_coarse_kernel(const float* __restrict__ rays, int64_t n, BBox bb, float min_t_range,
                     float epsilon, const float* __restrict__ u, int T, float* __restrict__ t_min_out,
                     float* __restrict__ t_max_out, uint8_t* __restrict__ mask_out,
                     float* __restrict__ ts_out) {
  const int per_ray = T / V;
  const int64_t items = n * per_ray;
  for (int64_t it = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; it < items;
       it += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = it / per_ray;
    const int i0 = int(it - r * per_ray) * V;
    float t_min, t_max;
    bool mask;
    ray_t_range(rays + r * 6, bb, min_t_range, epsilon, t_min, t_max, mask);
    if (i0 == 0) {
      t_min_out[r] = t_min;
      t_max_out[r] = t_max;
      mask_out[r] = mask ? 1 : 0;
    }
    // stratified_sampling, render.py:138-143
    const float bin = __fdiv_rn(__fsub_rn(t_max, t_min), float(T));
    float uv[V], out[V];
    if (V == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(u + r * T + i0));
      uv[0] = q.x; uv[1 % V] = q.y; uv[2 % V] = q.z; uv[3 % V] = q.w;
    } else {
      uv[0] = __ldg(u + r * T + i0);
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float start = __fadd_rn(__fmul_rn(float(i0 + k), bin), t_min);
      out[k] = __fadd_rn(start, __fmul_rn(uv[k], bin));
    }
    if (V == 4) *reinterpret_cast<float4*>(ts_out + r * T + i0) = make_float4(out[0], out[1 % V], out[2 % V], out[3 % V]);
    else ts_out[r * T + i0] = out[0];
  }
}

// stratified_sampling (render.py:138-143) on precomputed bounds, for the standalone
// RaySamples.stratified_sampling API; one thread per sample.
__global__ void __launch_bounds__(256)
stratified_kernel(const float* __restrict__ t_min, const float* __restrict__ t_max,
                  const float* __restrict__ u, int64_t n, int T, float* __restrict__ ts_out) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n * T;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / T;
    const int k = int(i - r * T);
    const float lo = __ldg(t_min + r);
    const float bin = __fdiv_rn(__fsub_rn(__ldg(t_max + r), lo), float(T));
    ts_out[i] = __fadd_rn(__fadd_rn(__fmul_rn(float(k), bin), lo), __fmul_rn(__ldg(u + i), bin));
  }
}

// ------------------------------------------------------------------ K4
// dynamic smem per warp (floats): ts[Tc] ddt[Tc] accprev[Tc] w[Tc] xs[Tc+1] ys[Tc+1] sort[P]
__host__ __device__ inline int fine_smem_floats(int Tc, int P) { return 4 * Tc + 2 * (Tc + 1) + P; }

__global__ void __launch_bounds__(256)
sample_fine_kernel(const float* __restrict__ ts_c, const float* __restrict__ dens_c,
                   const float* __restrict__ t_min_in, const float* __restrict__ t_max_in,
                   const float* __restrict__ u, int64_t n, int Tc, int Tf, int P, float eps,
                   float* __restrict__ ts_out, int32_t* __restrict__ idx_out,
                   float* __restrict__ new_ts_out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float* s_ts = smem + size_t(wib) * fine_smem_floats(Tc, P);
  float* s_ddt = s_ts + Tc;
  float* s_acc = s_ddt + Tc;
  float* s_w = s_acc + Tc;
  float* s_xs = s_w + Tc;
  float* s_ys = s_xs + Tc + 1;
  float* s_sort = s_ys + Tc + 1;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int T2 = Tc + Tf;
  const float ubin = __fdiv_rn(__fsub_rn(1.0f, 0.0f), float(Tf));  // render.py:244-250 -> :138

  for (int64_t r = warp; r < n; r += nwarps) {
    const float t_min = __ldg(t_min_in + r), t_max = __ldg(t_max_in + r);
    for (int i = lane; i < Tc; i += 32) s_ts[i] = __ldg(ts_c + r * Tc + i);
    __syncwarp();
    // starts/ends/deltas (render.py:259-268) and density_dt (:271)
    for (int i = lane; i < Tc; i += 32) {
      float t = s_ts[i];
      float start = (i == 0) ? t_min : __fdiv_rn(__fadd_rn(t, s_ts[i - 1]), 2.0f);
      float end = (i == Tc - 1) ? t_max : __fdiv_rn(__fadd_rn(s_ts[i + 1], t), 2.0f);
      float delta = __fsub_rn(end, start);
      s_ddt[i] = __fmul_rn(__ldg(dens_c + r * Tc + i), delta);
      s_ys[i + 1] = end;  // ys = [t_min, ends()]  (:238-241)
      s_sort[i] = t;
    }
    if (lane == 0) s_ys[0] = t_min;
    __syncwarp();
    // cumsum(density_dt) strictly sequential (:275); acc_prev = [0, acc][:Tc]
    if (lane == 0) {
      float acc = 0.0f;
      for (int i = 0; i < Tc; ++i) {
        s_acc[i] = acc;
        acc = __fadd_rn(acc, s_ddt[i]);
      }
    }
    __syncwarp();
    // w = surv * term + eps  (:279-287, :232)
    for (int i = lane; i < Tc; i += 32) {
      float surv = lnrf_expf(-s_acc[i]);
      float term = __fsub_rn(1.0f, lnrf_expf(-s_ddt[i]));
      s_w[i] = __fadd_rn(__fmul_rn(surv, term), eps);
    }
    __syncwarp();
    // xs = [0, cumsum(w)] (:235-236), sequential
    if (lane == 0) {
      float acc = 0.0f;
      s_xs[0] = 0.0f;
      for (int i = 0; i < Tc; ++i) {
        acc = __fadd_rn(acc, s_w[i]);
        s_xs[i + 1] = acc;
      }
    }
    __syncwarp();
    const float total = s_xs[Tc];
    __syncwarp();
    for (int i = lane; i <= Tc; i += 32) s_xs[i] = __fdiv_rn(s_xs[i], total);  // :237
    __syncwarp();
    // inverse CDF at stratified points: vmap(jnp.interp) (:251)
    const float xp_first = s_xs[0], xp_last = s_xs[Tc];
    const float fp_first = s_ys[0], fp_last = s_ys[Tc];
    for (int j = lane; j < Tf; j += 32) {
      float x = __fadd_rn(__fmul_rn(float(j), ubin), __fmul_rn(__ldg(u + r * Tf + j), ubin));
      // searchsorted(xs, x, side='right') over Tc+1 entries
      int lo = 0, hi = Tc + 1;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (s_xs[mid] <= x) lo = mid + 1; else hi = mid;
      }
      int i = min(max(lo, 1), Tc);
      float df = __fsub_rn(s_ys[i], s_ys[i - 1]);
      float dx = __fsub_rn(s_xs[i], s_xs[i - 1]);
      float delta = __fsub_rn(x, s_xs[i - 1]);
      bool dx0 = fabsf(dx) <= 1.4210854715202004e-14f;  // np.spacing(finfo(f32).eps)
      float ratio = __fdiv_rn(delta, dx0 ? 1.0f : dx);
      float f = dx0 ? s_ys[i - 1] : __fadd_rn(s_ys[i - 1], __fmul_rn(ratio, df));
      if (x < xp_first) f = fp_first;
      if (x > xp_last) f = fp_last;
      s_sort[Tc + j] = f;
      if (idx_out) idx_out[r * Tf + j] = i;
      if (new_ts_out) new_ts_out[r * Tf + j] = f;
    }
    for (int i = T2 + lane; i < P; i += 32) s_sort[i] = CUDART_INF_F;
    __syncwarp();
    // jnp.sort(concat[ts, new_ts]) (:253-255): bitonic network over P >= T2 slots
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < P; i += 32) {
          int l = i ^ j;
          if (l > i) {
            float a = s_sort[i], b = s_sort[l];
            bool up = (i & k) == 0;
            if (up ? (a > b) : (a < b)) {
              s_sort[i] = b;
              s_sort[l] = a;
            }
          }
        }
        __syncwarp();
      }
    }
    for (int i = lane; i < T2; i += 32) ts_out[r * T2 + i] = s_sort[i];
    __syncwarp();
  }
}


// ------------------------------------------------------------------ K4, default sizes (64 + 128)
// Same arithmetic as sample_fine_kernel (bit-exact), organised for throughput: every lane keeps
// its two coarse samples in registers, the two strictly sequential cumulative sums run as
// warp-synchronous shuffle loops (every lane carries the same running sum, no shared-memory
// round trips), and the final jnp.sort of [coarse | new] is a rank-based two-way merge (each
// element's position = its index + its rank in the other list) whenever both lists are already
// sorted -- which is checked per ray; otherwise (an fp32 rounding inversion) the bitonic network
// of the generic kernel runs.  A sorted array is unique, so both give the same bits.
constexpr int kF64Smem = 66 + 66 + 64 + 128 + 256;  // xs, ys, a, b, out (floats per warp)

__global__ void __launch_bounds__(256)
sample_fine64_kernel(const float* __restrict__ ts_c, const float* __restrict__ dens_c,
                     const float* __restrict__ t_min_in, const float* __restrict__ t_max_in,
                     const float* __restrict__ u, int64_t n, float eps, float* __restrict__ ts_out,
                     int32_t* __restrict__ idx_out, float* __restrict__ new_ts_out) {
  constexpr int Tc = 64, Tf = 128, T2 = 192;
  __shared__ float smem[8 * kF64Smem];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* s_xs = smem + wib * kF64Smem;
  float* s_ys = s_xs + 66;
  float* s_a = s_ys + 66;
  float* s_b = s_a + 64;
  float* s_out = s_b + 128;
  const unsigned full = 0xffffffffu;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const float ubin = __fdiv_rn(__fsub_rn(1.0f, 0.0f), float(Tf));
  for (int64_t r = warp; r < n; r += nwarps) {
    const float t_min = __ldg(t_min_in + r), t_max = __ldg(t_max_in + r);
    const float ts0 = __ldg(ts_c + r * Tc + lane), ts1 = __ldg(ts_c + r * Tc + 32 + lane);
    const float de0 = __ldg(dens_c + r * Tc + lane), de1 = __ldg(dens_c + r * Tc + 32 + lane);
    float uq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) uq[q] = __ldg(u + r * Tf + q * 32 + lane);
    // neighbours of the lane's two samples i = lane and i = 32 + lane
    float prev0 = __shfl_up_sync(full, ts0, 1), next0 = __shfl_down_sync(full, ts0, 1);
    float prev1 = __shfl_up_sync(full, ts1, 1), next1 = __shfl_down_sync(full, ts1, 1);
    const float ts1_first = __shfl_sync(full, ts1, 0), ts0_last = __shfl_sync(full, ts0, 31);
    if (lane == 31) next0 = ts1_first;
    if (lane == 0) prev1 = ts0_last;
    // starts/ends/deltas (render.py:259-268), density_dt (:271)
    const float start0 = lane == 0 ? t_min : __fdiv_rn(__fadd_rn(ts0, prev0), 2.0f);
    const float end0 = __fdiv_rn(__fadd_rn(next0, ts0), 2.0f);
    const float start1 = __fdiv_rn(__fadd_rn(ts1, prev1), 2.0f);
    const float end1 = lane == 31 ? t_max : __fdiv_rn(__fadd_rn(next1, ts1), 2.0f);
    const float ddt0 = __fmul_rn(de0, __fsub_rn(end0, start0));
    const float ddt1 = __fmul_rn(de1, __fsub_rn(end1, start1));
    // cumsum(density_dt), strictly sequential (:275): exclusive prefixes of the lane's samples
    float acc = 0.0f, accp0 = 0.0f, accp1 = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float v = __shfl_sync(full, ddt0, i);
      if (lane == i) accp0 = acc;
      acc = __fadd_rn(acc, v);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float v = __shfl_sync(full, ddt1, i);
      if (lane == i) accp1 = acc;
      acc = __fadd_rn(acc, v);
    }
    // w = surv * term + eps (:279-287, :232)
    const float w0 = __fadd_rn(__fmul_rn(lnrf_expf(-accp0), __fsub_rn(1.0f, lnrf_expf(-ddt0))), eps);
    const float w1 = __fadd_rn(__fmul_rn(lnrf_expf(-accp1), __fsub_rn(1.0f, lnrf_expf(-ddt1))), eps);
    // xs = [0, cumsum(w)] (:235-236), sequential; then / total (:237)
    float xa0 = 0.0f, xa1 = 0.0f;
    acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      acc = __fadd_rn(acc, __shfl_sync(full, w0, i));
      if (lane == i) xa0 = acc;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      acc = __fadd_rn(acc, __shfl_sync(full, w1, i));
      if (lane == i) xa1 = acc;
    }
    const float total = acc;
    __syncwarp();  // the previous ray's readers of the shared arrays are done
    if (lane == 0) {
      s_xs[0] = __fdiv_rn(0.0f, total);
      s_ys[0] = t_min;  // ys = [t_min, ends()] (:238-241)
    }
    s_xs[1 + lane] = __fdiv_rn(xa0, total);
    s_xs[33 + lane] = __fdiv_rn(xa1, total);
    s_ys[1 + lane] = end0;
    s_ys[33 + lane] = end1;
    s_a[lane] = ts0;
    s_a[32 + lane] = ts1;
    __syncwarp();
    // inverse CDF at stratified points: vmap(jnp.interp) (:251)
    const float xp_first = s_xs[0], xp_last = s_xs[Tc];
    const float fp_first = s_ys[0], fp_last = s_ys[Tc];
    float fq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = q * 32 + lane;
      const float x = __fadd_rn(__fmul_rn(float(j), ubin), __fmul_rn(uq[q], ubin));
      int lo = 0, hi = Tc + 1;  // searchsorted(xs, x, side='right') over Tc + 1 entries
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_xs[mid] <= x) lo = mid + 1; else hi = mid;
      }
      const int i = min(max(lo, 1), Tc);
      const float df = __fsub_rn(s_ys[i], s_ys[i - 1]);
      const float dx = __fsub_rn(s_xs[i], s_xs[i - 1]);
      const float delta = __fsub_rn(x, s_xs[i - 1]);
      const bool dx0 = fabsf(dx) <= 1.4210854715202004e-14f;  // np.spacing(finfo(f32).eps)
      const float ratio = __fdiv_rn(delta, dx0 ? 1.0f : dx);
      float f = dx0 ? s_ys[i - 1] : __fadd_rn(s_ys[i - 1], __fmul_rn(ratio, df));
      if (x < xp_first) f = fp_first;
      if (x > xp_last) f = fp_last;
      fq[q] = f;
      s_b[j] = f;
      if (idx_out) idx_out[r * Tf + j] = i;
      if (new_ts_out) new_ts_out[r * Tf + j] = f;
    }
    // ---- jnp.sort(concat[ts, new_ts]) (:253-255)
    bool ok = (ts0 <= next0) && (lane == 31 || ts1 <= next1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float nx = __shfl_down_sync(full, fq[q], 1);
      const float wrap = __shfl_sync(full, fq[(q + 1) & 3], 0);
      if (lane == 31) nx = wrap;
      if (!(q == 3 && lane == 31)) ok = ok && (fq[q] <= nx);
    }
    __syncwarp();
    if (__all_sync(full, ok)) {
      // two-way merge by ranks: a[i] goes to i + #{b < a[i]}, b[j] goes to j + #{a <= b[j]}
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float v = h ? ts1 : ts0;
        int lo = 0, hi = Tf;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (s_b[mid] < v) lo = mid + 1; else hi = mid;
        }
        s_out[h * 32 + lane + lo] = v;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float v = fq[q];
        int lo = 0, hi = Tc;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (s_a[mid] <= v) lo = mid + 1; else hi = mid;
        }
        s_out[q * 32 + lane + lo] = v;
      }
      __syncwarp();
    } else {
      s_out[lane] = ts0;
      s_out[32 + lane] = ts1;
#pragma unroll
      for (int q = 0; q < 4; ++q) s_out[64 + q * 32 + lane] = fq[q];
      s_out[192 + lane] = CUDART_INF_F;
      s_out[224 + lane] = CUDART_INF_F;
      __syncwarp();
      for (int k = 2; k <= 256; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = lane; i < 256; i += 32) {
            const int l = i ^ j;
            if (l > i) {
              const float x = s_out[i], y = s_out[l];
              const bool up = (i & k) == 0;
              if (up ? (x > y) : (x < y)) {
                s_out[i] = y;
                s_out[l] = x;
              }
            }
          }
          __syncwarp();
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) ts_out[r * T2 + k * 32 + lane] = s_out[k * 32 + lane];
  }
}


// ------------------------------------------------------------------ RaySamples helpers
// starts / ends / deltas (render.py:259-268) and termination_probs (render.py:270-287) as
// stand-alone calls for the RaySamples method surface; the same individually rounded arithmetic
// as the fine-sampling kernels above (strictly sequential cumsum, lnrf_expf): bit-exact with
// oracle.render_np.  One warp per ray, T <= 1024.
__global__ void __launch_bounds__(256)
ray_intervals_kernel(const float* __restrict__ ts, const float* __restrict__ t_min_in,
                     const float* __restrict__ t_max_in, int64_t n, int T, float* __restrict__ starts,
                     float* __restrict__ ends, float* __restrict__ deltas) {
  const int64_t total = n * T;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / T;
    const int k = int(i - r * T);
    const float t = __ldg(ts + i);
    const float start = (k == 0) ? __ldg(t_min_in + r) : __fdiv_rn(__fadd_rn(t, __ldg(ts + i - 1)), 2.0f);
    const float end = (k == T - 1) ? __ldg(t_max_in + r) : __fdiv_rn(__fadd_rn(__ldg(ts + i + 1), t), 2.0f);
    if (starts) starts[i] = start;
    if (ends) ends[i] = end;
    if (deltas) deltas[i] = __fsub_rn(end, start);
  }
}

__global__ void __launch_bounds__(256)
termination_probs_kernel(const float* __restrict__ ts, const float* __restrict__ t_min_in,
                         const float* __restrict__ t_max_in, const float* __restrict__ dens, int64_t n, int T,
                         float* __restrict__ probs) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* s_ts = smem + size_t(wib) * (3 * T + 1);
  float* s_ddt = s_ts + T;
  float* s_acc = s_ddt + T;  // T + 1 entries: acc_prev incl. the total
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += nwarps) {
    const float t_min = __ldg(t_min_in + r), t_max = __ldg(t_max_in + r);
    for (int i = lane; i < T; i += 32) s_ts[i] = __ldg(ts + r * T + i);
    __syncwarp();
    for (int i = lane; i < T; i += 32) {
      const float t = s_ts[i];
      const float start = (i == 0) ? t_min : __fdiv_rn(__fadd_rn(t, s_ts[i - 1]), 2.0f);
      const float end = (i == T - 1) ? t_max : __fdiv_rn(__fadd_rn(s_ts[i + 1], t), 2.0f);
      s_ddt[i] = __fmul_rn(__ldg(dens + r * T + i), __fsub_rn(end, start));  // :271
    }
    __syncwarp();
    if (lane == 0) {  // jnp.cumsum, strictly left to right (:275)
      float acc = 0.0f;
      for (int i = 0; i < T; ++i) {
        s_acc[i] = acc;
        acc = __fadd_rn(acc, s_ddt[i]);
      }
      s_acc[T] = acc;
    }
    __syncwarp();
    for (int i = lane; i <= T; i += 32) {  // prob_survive * prob_terminate (:279-287)
      const float surv = lnrf_expf(-s_acc[i]);
      const float term = (i == T) ? 1.0f : __fsub_rn(1.0f, lnrf_expf(-s_ddt[i]));
      probs[r * (T + 1) + i] = __fmul_rn(surv, term);
    }
    __syncwarp();
  }
}

}  // namespace lnrf

extern "C" {

int lnrf_ray_intervals(const float* ts, const float* t_min, const float* t_max, int64_t n, int32_t T,
                       float* starts, float* ends, float* deltas, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && T > 0, LNRF_E_INVALID, "lnrf_ray_intervals: n=%lld T=%d", (long long)n, T);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(ts && t_min && t_max, LNRF_E_INVALID, "lnrf_ray_intervals: null pointer");
  int64_t blocks = lnrf::ceil_div(n * T, 256);
  const int64_t cap = int64_t(lnrf::sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  lnrf::ray_intervals_kernel<<<(unsigned)blocks, 256, 0, lnrf::as_stream(stream)>>>(ts, t_min, t_max, n, T, starts,
                                                                                    ends, deltas);
  LNRF_LAUNCH_CHECK("ray_intervals_kernel");
  return LNRF_OK;
}

int lnrf_termination_probs(const float* ts, const float* t_min, const float* t_max, const float* dens, int64_t n,
                           int32_t T, float* probs, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && T > 0, LNRF_E_INVALID, "lnrf_termination_probs: n=%lld T=%d", (long long)n, T);
  LNRF_REQUIRE(T <= 1024, LNRF_E_UNSUPPORTED, "lnrf_termination_probs: T=%d exceeds 1024", T);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(ts && t_min && t_max && dens && probs, LNRF_E_INVALID, "lnrf_termination_probs: null pointer");
  const int warps = 8;
  const size_t smem = size_t(warps) * (3 * T + 1) * sizeof(float);  // <= 98 KB at T = 1024
  if (smem > 48 * 1024)
    LNRF_CUDA(cudaFuncSetAttribute(lnrf::termination_probs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   96 * 1024 + 1024));
  int64_t blocks = lnrf::ceil_div(n, warps);
  const int64_t cap = int64_t(lnrf::sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  lnrf::termination_probs_kernel<<<(unsigned)blocks, warps * 32, smem, lnrf::as_stream(stream)>>>(ts, t_min, t_max,
                                                                                               dens, n, T, probs);
  LNRF_LAUNCH_CHECK("termination_probs_kernel");
  return LNRF_OK;
}

int lnrf_sample_coarse(const float* rays, int64_t n, const float* bbox_min_host,
                       const float* bbox_max_host, float min_t_range, float epsilon,
                       const float* u, int32_t T, float* t_min, float* t_max, uint8_t* mask,
                       float* ts, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && T > 0, LNRF_E_INVALID, "lnrf_sample_coarse: n=%lld T=%d", (long long)n, T);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(rays && bbox_min_host && bbox_max_host && u && t_min && t_max && mask && ts,
               LNRF_E_INVALID, "lnrf_sample_coarse: null pointer");
  lnrf::BBox bb;
  for (int a = 0; a < 3; ++a) {
    bb.lo[a] = bbox_min_host[a];
    bb.hi[a] = bbox_max_host[a];
  }
  const int threads = 256;
  const bool vec = (T % 4 == 0) && ((uintptr_t)u % 16 == 0) && ((uintptr_t)ts % 16 == 0);
  int64_t blocks = lnrf::ceil_div(n * (vec ? T / 4 : T), threads);
  int64_t cap = int64_t(lnrf::sm_count()) * 32;
  if (blocks > cap) blocks = cap;
  if (vec)
    lnrf::sample_coarse_kernel<4><<<(unsigned)blocks, threads, 0, lnrf::as_stream(stream)>>>(
        rays, n, bb, min_t_range, epsilon, u, T, t_min, t_max, mask, ts);
  else
    lnrf::sample_coarse_kernel<1><<<(unsigned)blocks, threads, 0, lnrf::as_stream(stream)>>>(
        rays, n, bb, min_t_range, epsilon, u, T, t_min, t_max, mask, ts);
  LNRF_LAUNCH_CHECK("sample_coarse_kernel");
  return LNRF_OK;
}

int lnrf_stratified(const float* t_min, const float* t_max, const float* u, int64_t n, int32_t T,
                    float* ts, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && T > 0, LNRF_E_INVALID, "lnrf_stratified: n=%lld T=%d", (long long)n, T);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(t_min && t_max && u && ts, LNRF_E_INVALID, "lnrf_stratified: null pointer");
  int64_t blocks = lnrf::ceil_div(n * T, 256);
  int64_t cap = int64_t(lnrf::sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  lnrf::stratified_kernel<<<(unsigned)blocks, 256, 0, lnrf::as_stream(stream)>>>(t_min, t_max, u, n,
                                                                                T, ts);
  LNRF_LAUNCH_CHECK("stratified_kernel");
  return LNRF_OK;
}

int lnrf_sample_fine(const float* ts_c, const float* dens_c, const float* t_min,
                     const float* t_max, const float* u, int64_t n, int32_t Tc, int32_t Tf,
                     float eps, float* ts_out, int32_t* idx_out, float* new_ts_out,
                     lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && Tc > 0 && Tf > 0, LNRF_E_INVALID, "lnrf_sample_fine: n=%lld Tc=%d Tf=%d",
               (long long)n, Tc, Tf);
  LNRF_REQUIRE(Tc <= 256 && Tc + Tf <= 1024, LNRF_E_UNSUPPORTED,
               "lnrf_sample_fine: Tc=%d Tf=%d exceeds Tc<=256, Tc+Tf<=1024", Tc, Tf);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(ts_c && dens_c && t_min && t_max && u && ts_out, LNRF_E_INVALID,
               "lnrf_sample_fine: null pointer");
  if (Tc == 64 && Tf == 128) {  // the reference's default sample counts: the fast kernel
    int64_t blocks = lnrf::ceil_div(n, 8);
    const int64_t cap = int64_t(lnrf::sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    lnrf::sample_fine64_kernel<<<(unsigned)blocks, 256, 0, lnrf::as_stream(stream)>>>(
        ts_c, dens_c, t_min, t_max, u, n, eps, ts_out, idx_out, new_ts_out);
    LNRF_LAUNCH_CHECK("sample_fine64_kernel");
    return LNRF_OK;
  }
  int P = 2;
  while (P < Tc + Tf) P <<= 1;
  const int warps = 8, threads = warps * 32;
  size_t smem = size_t(warps) * lnrf::fine_smem_floats(Tc, P) * sizeof(float);
  if (smem > 48 * 1024)  // per call: the attribute is per device, and setting it is cheap and idempotent
    LNRF_CUDA(cudaFuncSetAttribute(lnrf::sample_fine_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = lnrf::ceil_div(n, warps);
  int64_t cap = int64_t(lnrf::sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  lnrf::sample_fine_kernel<<<(unsigned)blocks, threads, smem, lnrf::as_stream(stream)>>>(
      ts_c, dens_c, t_min, t_max, u, n, Tc, Tf, P, eps, ts_out, idx_out, new_ts_out);
  LNRF_LAUNCH_CHECK("sample_fine_kernel");
  return LNRF_OK;
}

}  // extern "C"
