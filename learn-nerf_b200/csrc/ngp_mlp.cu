// InstantNGPModel heads (learn_nerf/instant_ngp.py:37,46-53) on a precomputed encoding, fp32:
//   Dense_0: 2L -> 64 relu;  Dense_1: 64 -> 16 (col 0 -> exp -> density);
//   [d_emb(24) | out(16)] -> Dense_2: 40 -> 64 relu;  Dense_3: 64 -> 64 relu;  Dense_4: 64 -> 3 tanh
//
// The five layers are far too small for tiled GEMMs (9,920 MAC per sample), so the forward and
// the dX chain of the backward are each ONE fused kernel: a thread owns a sample and keeps the
// 64-wide activations in registers, all weights (40 KB) sit in shared memory and are read as
// warp-uniform 128-bit broadcasts.  Only what the weight gradients need (layer inputs and
// per-layer dL/dpre-activation) goes to HBM; dW = act^T g runs on the split-K FFMA GEMM.
#include "embed.cuh"
#include "lnrf_common.cuh"
#include "lnrf_math.cuh"
#include "sgemm.cuh"

namespace lnrf {

constexpr int kNgpMaxLevels = 16;
constexpr int kNgpHidden = 64, kNgpDensity = 16, kNgpDE = 24;
constexpr int kNgpIn2 = kNgpDE + kNgpDensity;  // 40
constexpr int kNgpThreads = 128;

struct NgpLayout {
  int in[5], out[5];
  int64_t w[5], b[5], total;
};
static NgpLayout ngp_layout(int L) {
  NgpLayout n{};
  const int ins[5] = {2 * L, kNgpHidden, kNgpIn2, kNgpHidden, kNgpHidden};
  const int outs[5] = {kNgpHidden, kNgpDensity, kNgpHidden, kNgpHidden, 3};
  int64_t off = 0;
  for (int i = 0; i < 5; ++i) {
    n.in[i] = ins[i];
    n.out[i] = outs[i];
    n.w[i] = off;
    off = align_up(off + int64_t(ins[i]) * outs[i], 4);
    n.b[i] = off;
    off = align_up(off + outs[i], 4);
  }
  n.total = off;
  return n;
}

struct NgpWs {
  float *h0, *in2, *h2, *h3;     // layer inputs kept for dW: [m,64], [m,40] = [d_emb | out], [m,64], [m,64]
  float *g0, *go1, *g2, *g3;     // dL/d pre-activation of Dense_0..3: [m,64], [m,16], [m,64], [m,64]
  int64_t bytes;
};
static NgpWs carve_ngp(void* base, int64_t m) {
  NgpWs w{};
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t floats) {
    float* r = reinterpret_cast<float*>(p + off);
    off += align_up(floats * 4, 256);
    return r;
  };
  w.h0 = take(m * kNgpHidden);
  w.in2 = take(m * kNgpIn2);
  w.h2 = take(m * kNgpHidden);
  w.h3 = take(m * kNgpHidden);
  w.g0 = take(m * kNgpHidden);
  w.go1 = take(m * kNgpDensity);
  w.g2 = take(m * kNgpHidden);
  w.g3 = take(m * kNgpHidden);
  w.bytes = off;
  return w;
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// acc[0..N) += x * row[0..N)  (row in shared memory, warp-uniform address -> broadcast)
template <int N>
__device__ __forceinline__ void axpy_row(float x, const float* __restrict__ row, float (&acc)[N]) {
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    const float4 w = lds4(row + j);
    acc[j] = fmaf(x, w.x, acc[j]);
    acc[j + 1] = fmaf(x, w.y, acc[j + 1]);
    acc[j + 2] = fmaf(x, w.z, acc[j + 2]);
    acc[j + 3] = fmaf(x, w.w, acc[j + 3]);
  }
}
// sum_j row[j] * g[j]
template <int N>
__device__ __forceinline__ float dot_row(const float* __restrict__ row, const float (&g)[N]) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    const float4 w = lds4(row + j);
    a0 = fmaf(w.x, g[j], a0);
    a1 = fmaf(w.y, g[j + 1], a1);
    a2 = fmaf(w.z, g[j + 2], a2);
    a3 = fmaf(w.w, g[j + 3], a3);
  }
  return (a0 + a1) + (a2 + a3);
}
template <int N>
__device__ __forceinline__ void store_row(float* __restrict__ dst, const float (&v)[N]) {
#pragma unroll
  for (int j = 0; j < N; j += 4) reinterpret_cast<float4*>(dst)[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}

struct NgpFwdArgs {
  const float* P;
  NgpLayout nl;
  int E;  // 2L
  const float* enc;
  const float* d;
  const float* rays;
  int T;
  int64_t m;
  NgpWs ws;
  float* dens;
  float* rgb;
};

template <bool SAVE>
__global__ void __launch_bounds__(kNgpThreads)
ngp_mlp_fwd_kernel(const __grid_constant__ NgpFwdArgs a) {
  extern __shared__ __align__(16) float sw[];
  for (int i = threadIdx.x; i < int(a.nl.total); i += kNgpThreads) sw[i] = __ldg(a.P + i);
  __syncthreads();
  const float* W0 = sw + a.nl.w[0]; const float* B0 = sw + a.nl.b[0];
  const float* W1 = sw + a.nl.w[1]; const float* B1 = sw + a.nl.b[1];
  const float* W2 = sw + a.nl.w[2]; const float* B2 = sw + a.nl.b[2];
  const float* W3 = sw + a.nl.w[3]; const float* B3 = sw + a.nl.b[3];
  const float* W4 = sw + a.nl.w[4]; const float* B4 = sw + a.nl.b[4];
  for (int64_t s = int64_t(blockIdx.x) * kNgpThreads + threadIdx.x; s < a.m; s += int64_t(gridDim.x) * kNgpThreads) {
    // ---- Dense_0 (2L -> 64) + relu                                         instant_ngp.py:46-47
    float h[kNgpHidden];
#pragma unroll
    for (int j = 0; j < kNgpHidden; ++j) h[j] = B0[j];
    const float4* erow = reinterpret_cast<const float4*>(a.enc + s * a.E);
    for (int i4 = 0; i4 < a.E / 4; ++i4) {
      const float4 x = __ldg(erow + i4);
      axpy_row<kNgpHidden>(x.x, W0 + (i4 * 4 + 0) * kNgpHidden, h);
      axpy_row<kNgpHidden>(x.y, W0 + (i4 * 4 + 1) * kNgpHidden, h);
      axpy_row<kNgpHidden>(x.z, W0 + (i4 * 4 + 2) * kNgpHidden, h);
      axpy_row<kNgpHidden>(x.w, W0 + (i4 * 4 + 3) * kNgpHidden, h);
    }
#pragma unroll
    for (int j = 0; j < kNgpHidden; ++j) h[j] = fmaxf(h[j], 0.0f);
    if (SAVE) store_row<kNgpHidden>(a.ws.h0 + s * kNgpHidden, h);
    // ---- Dense_1 (64 -> 16); density = exp(out[0])                         :48-49
    float in2[kNgpIn2];
    {
      float o[kNgpDensity];
#pragma unroll
      for (int j = 0; j < kNgpDensity; ++j) o[j] = B1[j];
#pragma unroll
      for (int i = 0; i < kNgpHidden; ++i) axpy_row<kNgpDensity>(h[i], W1 + i * kNgpDensity, o);
      a.dens[s] = expf(o[0]);
#pragma unroll
      for (int j = 0; j < kNgpDensity; ++j) in2[kNgpDE + j] = o[j];
    }
    // ---- d_emb = sinusoidal_emb(d, 4)                                      :37
    {
      float dv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) dv[k] = a.d ? __ldg(a.d + s * 3 + k) : __ldg(a.rays + (s / a.T) * 6 + 3 + k);
#pragma unroll
      for (int dim = 0; dim < 3; ++dim)
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          float sn, cs;
          sincosf(dv[dim] * float(1 << f), &sn, &cs);
          in2[dim * 8 + f] = sn;
          in2[dim * 8 + 4 + f] = cs;
        }
    }
    if (SAVE) store_row<kNgpIn2>(a.ws.in2 + s * kNgpIn2, in2);
    // ---- Dense_2 (40 -> 64) + relu                                         :50-52
#pragma unroll
    for (int j = 0; j < kNgpHidden; ++j) h[j] = B2[j];
#pragma unroll
    for (int i = 0; i < kNgpIn2; ++i) axpy_row<kNgpHidden>(in2[i], W2 + i * kNgpHidden, h);
#pragma unroll
    for (int j = 0; j < kNgpHidden; ++j) h[j] = fmaxf(h[j], 0.0f);
    if (SAVE) store_row<kNgpHidden>(a.ws.h2 + s * kNgpHidden, h);
    // ---- Dense_3 (64 -> 64) + relu
    float h3[kNgpHidden];
#pragma unroll
    for (int j = 0; j < kNgpHidden; ++j) h3[j] = B3[j];
#pragma unroll
    for (int i = 0; i < kNgpHidden; ++i) axpy_row<kNgpHidden>(h[i], W3 + i * kNgpHidden, h3);
#pragma unroll
    for (int j = 0; j < kNgpHidden; ++j) h3[j] = fmaxf(h3[j], 0.0f);
    if (SAVE) store_row<kNgpHidden>(a.ws.h3 + s * kNgpHidden, h3);
    // ---- Dense_4 (64 -> 3) + tanh                                          :53
    float o0 = B4[0], o1 = B4[1], o2 = B4[2];
#pragma unroll
    for (int i = 0; i < kNgpHidden; ++i) {
      o0 = fmaf(h3[i], W4[i * 3 + 0], o0);
      o1 = fmaf(h3[i], W4[i * 3 + 1], o1);
      o2 = fmaf(h3[i], W4[i * 3 + 2], o2);
    }
    a.rgb[s * 3 + 0] = tanhf(o0);
    a.rgb[s * 3 + 1] = tanhf(o1);
    a.rgb[s * 3 + 2] = tanhf(o2);
  }
}

struct NgpBwdArgs {
  const float* P;
  NgpLayout nl;
  int E;
  int64_t m;
  NgpWs ws;
  const float* dens;
  const float* rgb;
  const float* d_dens;
  const float* d_rgb;
  float* d_enc;
};

// dX chain: g3, g2, g_out1, g0 (written for the dW GEMMs) and d_enc.
__global__ void __launch_bounds__(kNgpThreads)
ngp_mlp_bwd_kernel(const __grid_constant__ NgpBwdArgs a) {
  extern __shared__ __align__(16) float sw[];
  for (int i = threadIdx.x; i < int(a.nl.total); i += kNgpThreads) sw[i] = __ldg(a.P + i);
  __syncthreads();
  const float* W0 = sw + a.nl.w[0];
  const float* W1 = sw + a.nl.w[1];
  const float* W2 = sw + a.nl.w[2];
  const float* W3 = sw + a.nl.w[3];
  const float* W4 = sw + a.nl.w[4];
  for (int64_t s = int64_t(blockIdx.x) * kNgpThreads + threadIdx.x; s < a.m; s += int64_t(gridDim.x) * kNgpThreads) {
    float dp[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float y = __ldg(a.rgb + s * 3 + j);
      dp[j] = __ldg(a.d_rgb + s * 3 + j) * (1.0f - y * y);  // tanh'
    }
    // g3 = (dp @ W4^T) * [h3 > 0]
    float g3[kNgpHidden];
    {
      const float4* hrow = reinterpret_cast<const float4*>(a.ws.h3 + s * kNgpHidden);
#pragma unroll
      for (int i4 = 0; i4 < kNgpHidden / 4; ++i4) {
        const float4 hv = __ldg(hrow + i4);
        const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i4 * 4 + k;
          const float t = dp[0] * W4[i * 3 + 0] + dp[1] * W4[i * 3 + 1] + dp[2] * W4[i * 3 + 2];
          g3[i] = hh[k] > 0.0f ? t : 0.0f;
        }
      }
    }
    store_row<kNgpHidden>(a.ws.g3 + s * kNgpHidden, g3);
    // g2 = (g3 @ W3^T) * [h2 > 0]
    float g2[kNgpHidden];
    {
      const float4* hrow = reinterpret_cast<const float4*>(a.ws.h2 + s * kNgpHidden);
#pragma unroll
      for (int i4 = 0; i4 < kNgpHidden / 4; ++i4) {
        const float4 hv = __ldg(hrow + i4);
        const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i4 * 4 + k;
          const float t = dot_row<kNgpHidden>(W3 + i * kNgpHidden, g3);
          g2[i] = hh[k] > 0.0f ? t : 0.0f;
        }
      }
    }
    store_row<kNgpHidden>(a.ws.g2 + s * kNgpHidden, g2);
    // g_out1 = g2 @ W2[24:40]^T, plus dL/d out[0] += d_dens * density  (density = exp(out[0]))
    float go1[kNgpDensity];
#pragma unroll
    for (int i = 0; i < kNgpDensity; ++i) go1[i] = dot_row<kNgpHidden>(W2 + (kNgpDE + i) * kNgpHidden, g2);
    go1[0] += __ldg(a.d_dens + s) * __ldg(a.dens + s);
    store_row<kNgpDensity>(a.ws.go1 + s * kNgpDensity, go1);
    // g0 = (g_out1 @ W1^T) * [h0 > 0]
    float g0[kNgpHidden];
    {
      const float4* hrow = reinterpret_cast<const float4*>(a.ws.h0 + s * kNgpHidden);
#pragma unroll
      for (int i4 = 0; i4 < kNgpHidden / 4; ++i4) {
        const float4 hv = __ldg(hrow + i4);
        const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i4 * 4 + k;
          const float t = dot_row<kNgpDensity>(W1 + i * kNgpDensity, go1);
          g0[i] = hh[k] > 0.0f ? t : 0.0f;
        }
      }
    }
    store_row<kNgpHidden>(a.ws.g0 + s * kNgpHidden, g0);
    // d_enc = g0 @ W0^T
    float4* drow = reinterpret_cast<float4*>(a.d_enc + s * a.E);
    for (int i4 = 0; i4 < a.E / 4; ++i4) {
      float4 o;
      o.x = dot_row<kNgpHidden>(W0 + (i4 * 4 + 0) * kNgpHidden, g0);
      o.y = dot_row<kNgpHidden>(W0 + (i4 * 4 + 1) * kNgpHidden, g0);
      o.z = dot_row<kNgpHidden>(W0 + (i4 * 4 + 2) * kNgpHidden, g0);
      o.w = dot_row<kNgpHidden>(W0 + (i4 * 4 + 3) * kNgpHidden, g0);
      drow[i4] = o;
    }
  }
}

// dW4 += h3^T dp, db4 += sum dp with dp = d_rgb (1 - rgb^2): warp per sample, lane owns 2 inputs.
__global__ void __launch_bounds__(256)
ngp_dw4_kernel(const float* __restrict__ h3, const float* __restrict__ rgb, const float* __restrict__ d_rgb,
               int64_t m, float* __restrict__ dw4, float* __restrict__ db4) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float gw[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, gb[3] = {0.f, 0.f, 0.f};
  for (int64_t s = warp; s < m; s += nwarps) {
    float dp[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float y = __ldg(rgb + s * 3 + j);
      dp[j] = __ldg(d_rgb + s * 3 + j) * (1.0f - y * y);
      gb[j] += dp[j];
    }
    const float2 a = __ldg(reinterpret_cast<const float2*>(h3 + s * kNgpHidden) + lane);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      gw[0][j] = fmaf(a.x, dp[j], gw[0][j]);
      gw[1][j] = fmaf(a.y, dp[j], gw[1][j]);
    }
  }
  __shared__ float s_gw[8][kNgpHidden * 3];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) s_gw[wib][(lane * 2 + k) * 3 + j] = gw[k][j];
  __syncthreads();
  for (int i = threadIdx.x; i < kNgpHidden * 3; i += blockDim.x) {
    float t = 0.0f;
    for (int ww = 0; ww < 8; ++ww) t += s_gw[ww][i];
    atomicAdd(dw4 + i, t);
  }
  if (lane == 0) {
    atomicAdd(db4 + 0, gb[0]);
    atomicAdd(db4 + 1, gb[1]);
    atomicAdd(db4 + 2, gb[2]);
  }
}

static int ngp_grid(int64_t m) {
  int64_t blocks = ceil_div(m, kNgpThreads);
  const int64_t cap = int64_t(sm_count()) * 3;
  return int(blocks < cap ? blocks : cap);
}

}  // namespace lnrf

extern "C" {

int64_t lnrf_ngp_mlp_param_count(int32_t L) { return lnrf::ngp_layout(L).total; }

int lnrf_ngp_mlp_param_offsets(int32_t L, int64_t* out_host) {
  LNRF_REQUIRE(out_host && L >= 1 && L <= lnrf::kNgpMaxLevels, LNRF_E_INVALID, "lnrf_ngp_mlp_param_offsets: bad args");
  const lnrf::NgpLayout n = lnrf::ngp_layout(L);
  for (int i = 0; i < 5; ++i) {
    out_host[2 * i] = n.w[i];
    out_host[2 * i + 1] = n.b[i];
  }
  return LNRF_OK;
}

int lnrf_ngp_mlp_workspace_bytes(int64_t m, int32_t L, int64_t* bytes_out_host) {
  (void)L;
  LNRF_REQUIRE(m >= 0 && bytes_out_host, LNRF_E_INVALID, "lnrf_ngp_mlp_workspace_bytes: bad args");
  *bytes_out_host = lnrf::carve_ngp(nullptr, m).bytes;
  return LNRF_OK;
}

int lnrf_ngp_mlp_fwd(const float* params, int32_t L, const float* enc, const float* d,
                     const float* rays, int64_t n, int32_t T, int32_t save_for_backward, void* workspace,
                     int64_t workspace_bytes, float* dens, float* rgb, lnrf_stream_t stream) {
  using namespace lnrf;
  LNRF_REQUIRE(n >= 0 && T >= 1 && L >= 1 && L <= kNgpMaxLevels, LNRF_E_INVALID,
               "lnrf_ngp_mlp_fwd: n=%lld T=%d L=%d", (long long)n, T, L);
  LNRF_REQUIRE((2 * L) % 4 == 0, LNRF_E_UNSUPPORTED, "lnrf_ngp_mlp_fwd: 2L=%d must be a multiple of 4", 2 * L);
  const int64_t m = n * T;
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(params && enc && dens && rgb && ((d && !rays) || (!d && rays)), LNRF_E_INVALID,
               "lnrf_ngp_mlp_fwd: null pointer / pass either d or rays");
  const bool save = save_for_backward != 0;
  NgpWs w{};
  if (save) {
    LNRF_REQUIRE(workspace && workspace_bytes >= carve_ngp(nullptr, m).bytes, LNRF_E_WORKSPACE,
                 "lnrf_ngp_mlp_fwd: workspace %lld < %lld bytes", (long long)workspace_bytes,
                 (long long)carve_ngp(nullptr, m).bytes);
    w = carve_ngp(workspace, m);
  }
  const NgpLayout nl = ngp_layout(L);
  NgpFwdArgs a{params, nl, 2 * L, enc, d, rays, T, m, w, dens, rgb};
  const size_t smem = size_t(nl.total) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    LNRF_CUDA(cudaFuncSetAttribute(ngp_mlp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    LNRF_CUDA(cudaFuncSetAttribute(ngp_mlp_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    LNRF_CUDA(cudaFuncSetAttribute(ngp_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    configured = true;
  }
  if (save) ngp_mlp_fwd_kernel<true><<<ngp_grid(m), kNgpThreads, smem, as_stream(stream)>>>(a);
  else ngp_mlp_fwd_kernel<false><<<ngp_grid(m), kNgpThreads, smem, as_stream(stream)>>>(a);
  LNRF_LAUNCH_CHECK("ngp_mlp_fwd_kernel");
  return LNRF_OK;
}

int lnrf_ngp_mlp_bwd(const float* params, int32_t L, const float* enc, int64_t m, void* workspace,
                     int64_t workspace_bytes, const float* dens, const float* rgb, const float* d_dens,
                     const float* d_rgb, float* d_params, float* d_enc, lnrf_stream_t stream) {
  using namespace lnrf;
  LNRF_REQUIRE(m >= 0 && L >= 1 && L <= kNgpMaxLevels && (2 * L) % 4 == 0, LNRF_E_INVALID,
               "lnrf_ngp_mlp_bwd: m=%lld L=%d", (long long)m, L);
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(params && enc && workspace && dens && rgb && d_dens && d_rgb && d_params && d_enc,
               LNRF_E_INVALID, "lnrf_ngp_mlp_bwd: null pointer");
  LNRF_REQUIRE(workspace_bytes >= carve_ngp(nullptr, m).bytes, LNRF_E_WORKSPACE,
               "lnrf_ngp_mlp_bwd: workspace too small");
  const NgpLayout nl = ngp_layout(L);
  const NgpWs w = carve_ngp(workspace, m);
  cudaStream_t st = as_stream(stream);
  float* G = d_params;
  NgpBwdArgs a{params, nl, 2 * L, m, w, dens, rgb, d_dens, d_rgb, d_enc};
  ngp_mlp_bwd_kernel<<<ngp_grid(m), kNgpThreads, size_t(nl.total) * sizeof(float), st>>>(a);
  LNRF_LAUNCH_CHECK("ngp_mlp_bwd_kernel");
  // weight / bias gradients: dW_l = input_l^T g_l (split-K FFMA GEMM), db_l = column sums
  int rc;
  ngp_dw4_kernel<<<ew_blocks(m, 8 * 16), 256, 0, st>>>(w.h3, rgb, d_rgb, m, G + nl.w[4], G + nl.b[4]);
  LNRF_LAUNCH_CHECK("ngp_dw4_kernel");
  if ((rc = gemm_tn_small(st, kNgpHidden, kNgpHidden, w.h2, kNgpHidden, w.g3, kNgpHidden, m, G + nl.w[3], kNgpHidden,
                          G + nl.b[3]))) return rc;
  if ((rc = gemm_tn_small(st, kNgpIn2, kNgpHidden, w.in2, kNgpIn2, w.g2, kNgpHidden, m, G + nl.w[2], kNgpHidden,
                          G + nl.b[2]))) return rc;
  if ((rc = gemm_tn_small(st, kNgpHidden, kNgpDensity, w.h0, kNgpHidden, w.go1, kNgpDensity, m, G + nl.w[1],
                          kNgpDensity, G + nl.b[1]))) return rc;
  if ((rc = gemm_tn_small(st, 2 * L, kNgpHidden, enc, 2 * L, w.g0, kNgpHidden, m, G + nl.w[0], kNgpHidden,
                          G + nl.b[0]))) return rc;
  return LNRF_OK;
}

}  // extern "C"
