// K2/K6, fp32 SIMT path: NeRFModel forward and backward with FFMA GEMMs.
// This is the 1e-5-accurate path (SURVEY §7 hard part 1); the tensor-core path is
// mlp_tc.cu.  Reference: learn_nerf/model.py:42-77 (forward); the backward is what
// jax.grad derives at train.py:90.
#include "embed.cuh"
#include "lnrf_common.cuh"
#include "lnrf_math.cuh"
#include "nerf_layout.cuh"
#include "sgemm.cuh"

namespace lnrf {

// ---------------------------------------------------------------- heads
// density = softplus(z8 . w9 + b9)  (model.py:57).  Warp per sample.
__global__ void __launch_bounds__(256)
density_head_fwd_kernel(const float* __restrict__ z8, const float* __restrict__ w9,
                        const float* __restrict__ b9, int64_t m, float* __restrict__ dens) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const float4 wa = __ldg(reinterpret_cast<const float4*>(w9) + lane);
  const float4 wb = __ldg(reinterpret_cast<const float4*>(w9) + 32 + lane);
  const float bias = __ldg(b9);
  for (int64_t s = warp; s < m; s += nwarps) {
    const float4* row = reinterpret_cast<const float4*>(z8 + s * kH);
    float4 a = __ldg(row + lane), b = __ldg(row + 32 + lane);
    float acc = a.x * wa.x + a.y * wa.y + a.z * wa.z + a.w * wa.w + b.x * wb.x + b.y * wb.y +
                b.z * wb.z + b.w * wb.w;
    acc = warp_sum(acc);
    if (lane == 0) dens[s] = softplus_f(acc + bias);
  }
}

// rgb = tanh(c . W11 + b11)  (model.py:60).  Warp per sample; lane owns 4 inputs.
__global__ void __launch_bounds__(256)
rgb_head_fwd_kernel(const float* __restrict__ c, const float* __restrict__ w11,
                    const float* __restrict__ b11, int64_t m, float* __restrict__ rgb) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float w[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) w[k][j] = __ldg(w11 + (lane * 4 + k) * 3 + j);
  const float bb = lane < 3 ? __ldg(b11 + lane) : 0.0f;
  for (int64_t s = warp; s < m; s += nwarps) {
    float4 a = __ldg(reinterpret_cast<const float4*>(c + s * kHC) + lane);
    const float av[4] = {a.x, a.y, a.z, a.w};
    float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < 3; ++j) o[j] = fmaf(av[k], w[k][j], o[j]);
#pragma unroll
    for (int j = 0; j < 3; ++j) o[j] = warp_sum(o[j]);
    if (lane < 3) rgb[s * 3 + lane] = tanhf((lane == 0 ? o[0] : (lane == 1 ? o[1] : o[2])) + bb);
  }
}

// Backward of the rgb head: dpre = d_rgb * (1 - rgb^2); dW11 += c^T dpre; db11 += sum dpre;
// dc = (dpre @ W11^T) * (c > 0).
__global__ void __launch_bounds__(256)
rgb_head_bwd_kernel(const float* __restrict__ c, const float* __restrict__ rgb,
                    const float* __restrict__ d_rgb, const float* __restrict__ w11, int64_t m,
                    float* __restrict__ dc, float* __restrict__ dw11, float* __restrict__ db11) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float w[4][3], gw[4][3], gb[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      w[k][j] = __ldg(w11 + (lane * 4 + k) * 3 + j);
      gw[k][j] = 0.0f;
    }
  for (int64_t s = warp; s < m; s += nwarps) {
    float dp[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float y = __ldg(rgb + s * 3 + j);
      dp[j] = __ldg(d_rgb + s * 3 + j) * (1.0f - y * y);
      gb[j] += dp[j];
    }
    float4 a = __ldg(reinterpret_cast<const float4*>(c + s * kHC) + lane);
    const float av[4] = {a.x, a.y, a.z, a.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t = dp[0] * w[k][0] + dp[1] * w[k][1] + dp[2] * w[k][2];
      o[k] = av[k] > 0.0f ? t : 0.0f;
#pragma unroll
      for (int j = 0; j < 3; ++j) gw[k][j] = fmaf(av[k], dp[j], gw[k][j]);
    }
    reinterpret_cast<float4*>(dc + s * kHC)[lane] = make_float4(o[0], o[1], o[2], o[3]);
  }
  // block-level reduction of the weight gradient, then one atomic per element per block
  __shared__ float s_gw[8][kHC * 3];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) s_gw[wib][(lane * 4 + k) * 3 + j] = gw[k][j];
  __syncthreads();
  for (int i = threadIdx.x; i < kHC * 3; i += blockDim.x) {
    float t = 0.0f;
    for (int ww = 0; ww < 8; ++ww) t += s_gw[ww][i];
    atomicAdd(dw11 + i, t);
  }
  if (lane == 0) {  // every lane holds the same gb (all lanes read the same dp)
    atomicAdd(db11 + 0, gb[0]);
    atomicAdd(db11 + 1, gb[1]);
    atomicAdd(db11 + 2, gb[2]);
  }
}

// Backward of the density head: spre = d_dens * sigmoid(pre) with sigmoid(pre) = 1 - exp(-dens);
// dW9 += z8^T spre; db9 += sum spre.  (d z8 += spre (x) w9 is folded into the next GEMM.)
__global__ void __launch_bounds__(256)
density_head_bwd_kernel(const float* __restrict__ z8, const float* __restrict__ dens,
                        const float* __restrict__ d_dens, int64_t m, float* __restrict__ spre,
                        float* __restrict__ dw9, float* __restrict__ db9) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float ga[4] = {0.f, 0.f, 0.f, 0.f}, gb4[4] = {0.f, 0.f, 0.f, 0.f}, gbias = 0.0f;
  for (int64_t s = warp; s < m; s += nwarps) {
    const float sp = __ldg(d_dens + s) * (-expm1f(-__ldg(dens + s)));
    if (lane == 0) spre[s] = sp;
    gbias += sp;
    const float4* row = reinterpret_cast<const float4*>(z8 + s * kH);
    float4 a = __ldg(row + lane), b = __ldg(row + 32 + lane);
    ga[0] = fmaf(a.x, sp, ga[0]); ga[1] = fmaf(a.y, sp, ga[1]);
    ga[2] = fmaf(a.z, sp, ga[2]); ga[3] = fmaf(a.w, sp, ga[3]);
    gb4[0] = fmaf(b.x, sp, gb4[0]); gb4[1] = fmaf(b.y, sp, gb4[1]);
    gb4[2] = fmaf(b.z, sp, gb4[2]); gb4[3] = fmaf(b.w, sp, gb4[3]);
  }
  __shared__ float s_g[8][kH];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s_g[wib][lane * 4 + k] = ga[k];
    s_g[wib][128 + lane * 4 + k] = gb4[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kH; i += blockDim.x) {
    float t = 0.0f;
    for (int ww = 0; ww < 8; ++ww) t += s_g[ww][i];
    atomicAdd(dw9 + i, t);
  }
  if (lane == 0) atomicAdd(db9, gbias);
}

// ---------------------------------------------------------------- workspace
struct Fp32Ws {
  float* xe;     // [m,60]
  float* de;     // [m,24]
  float* h[9];   // h0..h7 post-ReLU, h[8] = z8 pre-activation; aliased ping-pong when !save
  float* c;      // [m,128]
  float* spre;   // [m]
  float* gA;     // [m,256] backward ping-pong
  float* gB;
  float* dc;     // [m,128]
  int64_t bytes;
};

static Fp32Ws carve_fp32(void* base, int64_t m, bool save) {
  Fp32Ws w{};
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t floats) {
    float* r = reinterpret_cast<float*>(p + off);
    off += align_up(floats * 4, 256);
    return r;
  };
  w.xe = take(m * kXE);
  w.de = take(m * kDE);
  if (save) {
    for (int i = 0; i < 9; ++i) w.h[i] = take(m * kH);
  } else {
    float* a = take(m * kH);
    float* b = take(m * kH);
    for (int i = 0; i < 9; ++i) w.h[i] = (i & 1) ? b : a;
  }
  w.c = take(m * kHC);
  if (save) {
    w.spre = take(m);
    w.gA = take(m * kH);
    w.gB = take(m * kH);
    w.dc = take(m * kHC);
  }
  w.bytes = off;
  return w;
}

int64_t fp32_workspace_bytes(int64_t m, bool save) { return carve_fp32(nullptr, m, save).bytes; }

int nerf_fwd_fp32(const float* P, const float* x, const float* d, const float* rays, const float* ts,
                  int64_t m, int T, bool save, void* ws_base, int64_t ws_bytes, float* dens,
                  float* rgb, cudaStream_t st) {
  LNRF_REQUIRE(ws_bytes >= fp32_workspace_bytes(m, save), LNRF_E_WORKSPACE,
               "lnrf_nerf_mlp_fwd(fp32): workspace %lld < %lld bytes", (long long)ws_bytes,
               (long long)fp32_workspace_bytes(m, save));
  Fp32Ws w = carve_fp32(ws_base, m, save);
  embed_kernel<kXFreqs><<<ew_blocks(m * 3 * kXFreqs, 256), 256, 0, st>>>(x, rays, ts, T, 0, m, w.xe);
  LNRF_LAUNCH_CHECK("embed_kernel<x>");
  embed_kernel<kDFreqs><<<ew_blocks(m * 3 * kDFreqs, 256), 256, 0, st>>>(d, rays, ts, T, 1, m, w.de);
  LNRF_LAUNCH_CHECK("embed_kernel<d>");
  int rc;
  // input stack, model.py:50-51
  rc = gemm_nn<EPI_BIAS_RELU>(st, m, kH, w.xe, kXE, kXE, nullptr, 0, 0, P + kNerf.w[0], kH, w.h[0], kH,
                              P + kNerf.b[0]);
  if (rc) return rc;
  for (int l = 1; l <= 4; ++l) {
    rc = gemm_nn<EPI_BIAS_RELU>(st, m, kH, w.h[l - 1], kH, kH, nullptr, 0, 0, P + kNerf.w[l], kH,
                                w.h[l], kH, P + kNerf.b[l]);
    if (rc) return rc;
  }
  // skip concat [z | x_emb], model.py:52; Dense_5..7 outputs are consumed through ReLU (:53-56)
  rc = gemm_nn<EPI_BIAS_RELU>(st, m, kH, w.h[4], kH, kH, w.xe, kXE, kXE, P + kNerf.w[5], kH, w.h[5],
                              kH, P + kNerf.b[5]);
  if (rc) return rc;
  for (int l = 6; l <= 7; ++l) {
    rc = gemm_nn<EPI_BIAS_RELU>(st, m, kH, w.h[l - 1], kH, kH, nullptr, 0, 0, P + kNerf.w[l], kH,
                                w.h[l], kH, P + kNerf.b[l]);
    if (rc) return rc;
  }
  // Dense_8 output z is used raw by both heads (:57-58)
  rc = gemm_nn<EPI_BIAS>(st, m, kH, w.h[7], kH, kH, nullptr, 0, 0, P + kNerf.w[8], kH, w.h[8], kH,
                         P + kNerf.b[8]);
  if (rc) return rc;
  density_head_fwd_kernel<<<ew_blocks(m, 8), 256, 0, st>>>(w.h[8], P + kNerf.w[9], P + kNerf.b[9], m,
                                                           dens);
  LNRF_LAUNCH_CHECK("density_head_fwd_kernel");
  rc = gemm_nn<EPI_BIAS_RELU>(st, m, kHC, w.h[8], kH, kH, w.de, kDE, kDE, P + kNerf.w[10], kHC, w.c,
                              kHC, P + kNerf.b[10]);
  if (rc) return rc;
  rgb_head_fwd_kernel<<<ew_blocks(m, 8), 256, 0, st>>>(w.c, P + kNerf.w[11], P + kNerf.b[11], m, rgb);
  LNRF_LAUNCH_CHECK("rgb_head_fwd_kernel");
  return LNRF_OK;
}

int nerf_bwd_fp32(const float* P, int64_t m, void* ws_base, int64_t ws_bytes, const float* dens,
                  const float* rgb, const float* d_dens, const float* d_rgb, float* G,
                  cudaStream_t st) {
  LNRF_REQUIRE(ws_bytes >= fp32_workspace_bytes(m, true), LNRF_E_WORKSPACE,
               "lnrf_nerf_mlp_bwd(fp32): workspace %lld < %lld bytes", (long long)ws_bytes,
               (long long)fp32_workspace_bytes(m, true));
  Fp32Ws w = carve_fp32(ws_base, m, true);
  const unsigned rb = ew_blocks(m, 8 * 16);  // fewer, longer-lived blocks: less atomic traffic
  rgb_head_bwd_kernel<<<rb, 256, 0, st>>>(w.c, rgb, d_rgb, P + kNerf.w[11], m, w.dc, G + kNerf.w[11],
                                          G + kNerf.b[11]);
  LNRF_LAUNCH_CHECK("rgb_head_bwd_kernel");
  density_head_bwd_kernel<<<rb, 256, 0, st>>>(w.h[8], dens, d_dens, m, w.spre, G + kNerf.w[9],
                                              G + kNerf.b[9]);
  LNRF_LAUNCH_CHECK("density_head_bwd_kernel");
  int rc;
  const unsigned cb = ew_blocks(m, 512);  // >= 1k blocks at training sizes
  // colour layer Dense_10: input [z8 | d_emb]
  rc = gemm_tn_acc(st, kH, kHC, w.h[8], kH, w.dc, kHC, m, G + kNerf.w[10], kHC);
  if (rc) return rc;
  rc = gemm_tn_acc(st, kDE, kHC, w.de, kDE, w.dc, kHC, m, G + kNerf.w[10] + int64_t(kH) * kHC, kHC);
  if (rc) return rc;
  colsum_kernel<><<<cb, 256, 0, st>>>(w.dc, m, kHC, G + kNerf.b[10]);
  LNRF_LAUNCH_CHECK("colsum_kernel");
  // g8 = dc @ W10[:256]^T + spre (x) w9
  float* g = w.gA;
  float* gn = w.gB;
  rc = gemm_nt<EPI_RANK1>(st, m, kH, w.dc, kHC, kHC, P + kNerf.w[10], kHC, g, kH, nullptr, 0, w.spre,
                          P + kNerf.w[9]);
  if (rc) return rc;
  for (int l = 8; l >= 1; --l) {
    // dW_l = in_l^T g_l ; db_l = colsum(g_l) ; g_{l-1} = (g_l @ W_l[:256]^T) * (h_{l-1} > 0)
    rc = gemm_tn_acc(st, kH, kH, w.h[l - 1], kH, g, kH, m, G + kNerf.w[l], kH);
    if (rc) return rc;
    if (l == 5) {
      rc = gemm_tn_acc(st, kXE, kH, w.xe, kXE, g, kH, m, G + kNerf.w[5] + int64_t(kH) * kH, kH);
      if (rc) return rc;
    }
    colsum_kernel<><<<cb, 256, 0, st>>>(g, m, kH, G + kNerf.b[l]);
    LNRF_LAUNCH_CHECK("colsum_kernel");
    rc = gemm_nt<EPI_MASK>(st, m, kH, g, kH, kH, P + kNerf.w[l], kH, gn, kH, w.h[l - 1], kH);
    if (rc) return rc;
    float* t = g; g = gn; gn = t;
  }
  rc = gemm_tn_acc(st, kXE, kH, w.xe, kXE, g, kH, m, G + kNerf.w[0], kH);
  if (rc) return rc;
  colsum_kernel<><<<cb, 256, 0, st>>>(g, m, kH, G + kNerf.b[0]);
  LNRF_LAUNCH_CHECK("colsum_kernel");
  return LNRF_OK;
}

}  // namespace lnrf
