// K2/K6, fp32-accurate path: NeRFModel forward and backward as a chain of split-fp16 tcgen05 GEMMs
// (gemm_tc.cu: fp32 in / fp32 out, every product formed from fp16 hi/lo pairs on the tensor cores with
// fp32 accumulation) with fused bias / ReLU / mask epilogues and fused bias-gradient column sums.
// This is the 1e-5-accurate path (SURVEY §7 hard part 1); the bf16 path is mlp_tc_cta2_*.cu.  The FFMA
// GEMMs of sgemm.cuh this replaced stay reachable through LNRF_FP32_FFMA=1 (read once in lnrf_init, for
// A/B measurements).  Reference: learn_nerf/model.py:42-77 (forward); the backward is what jax.grad
// derives at train.py:90.
#include <stdlib.h>

#include "embed.cuh"
#include "gemm_tc.cuh"
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

__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Backward of the rgb head: dpre = d_rgb * (1 - rgb^2); dW11 += c^T dpre; db11 += sum dpre;
// dc = (dpre @ W11^T) * (c > 0).
__global__ void __launch_bounds__(256)
rgb_head_bwd_kernel(const float* __restrict__ c, const float* __restrict__ rgb,
                    const float* __restrict__ d_rgb, const float* __restrict__ w11, int64_t m,
                    float* __restrict__ dc, float* __restrict__ dw11, float* __restrict__ db11,
                    float* __restrict__ dc_amax) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float w[4][3], gw[4][3], gb[3] = {0.f, 0.f, 0.f};
  float amax = 0.0f;  // max |dc|: the scale of the first tensor-core operand of the backward chain
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
      amax = fmaxf(amax, fabsf(o[k]));
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
  amax = warp_max_f(amax);
  if (lane == 0 && dc_amax != nullptr && amax > 0.0f && amax < 3.0e38f)
    atomicMax(reinterpret_cast<unsigned int*>(dc_amax), __float_as_uint(amax));
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
  uint32_t* mask[8];  // ReLU bit masks of h0..h7 (tcg_mask_words each; training only)
  float* amax;   // [32] max |.| of the GEMM operands: 0..8 h0..h7, z8; 9 c; 10 dc; 11..19 g8..g0 (see gemm_tc.cuh)
  int64_t bytes;
};
constexpr int kAmaxC = 9, kAmaxDc = 10, kAmaxG = 11;  // g_l lives in slot kAmaxG + (8 - l)
static bool g_fp32_ffma = false;  // LNRF_FP32_FFMA=1 at lnrf_init: the FFMA GEMMs of sgemm.cuh (A/B measurements)
bool fp32_ffma() { return g_fp32_ffma; }
void init_mlp_fp32() {
  const char* e = getenv("LNRF_FP32_FFMA");
  g_fp32_ffma = e != nullptr && atoi(e) != 0;
}

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
  w.amax = take(32);
  if (save) {
    for (int i = 0; i < 8; ++i) w.mask[i] = reinterpret_cast<uint32_t*>(take(tcg_mask_words(m, kH)));
    w.spre = take(m);
    w.gA = take(m * kH);
    w.gB = take(m * kH);
    w.dc = take(m * kHC);
  }
  w.bytes = off;
  return w;
}

int64_t fp32_workspace_bytes(int64_t m, bool save) { return carve_fp32(nullptr, m, save).bytes; }

// ---- the three contractions, on the tensor cores (gemm_tc.cu) or, for A/B runs, on FFMA (sgemm.cuh)
template <int EPI>
static int dense_fwd(cudaStream_t st, int64_t m, int N, const float* A0, int lda0, int K0, const float* A1, int lda1,
                     int K1, const float* W, float* C, const float* bias, const float* a_amax, float* c_amax,
                     uint32_t* mask_out = nullptr) {
  if (g_fp32_ffma) return gemm_nn<EPI>(st, m, N, A0, lda0, K0, A1, lda1, K1, W, N, C, N, bias);
  return tcg_rows(st, EPI, false, m, N, A0, lda0, K0, A1, lda1, K1, W, N, C, N, bias, nullptr, 0, nullptr, nullptr,
                  a_amax, nullptr, c_amax, nullptr, mask_out);
}
// C[m,N] = epi(G[m,K] @ W[N rows, K cols]^T)
template <int EPI>
static int dense_dx(cudaStream_t st, int64_t m, int N, const float* G, int K, const float* W, float* C,
                    const float* aux, const float* r1s, const float* r1w, const float* a_amax, float* c_amax,
                    const uint32_t* mask_in = nullptr) {
  if (g_fp32_ffma) return gemm_nt<EPI>(st, m, N, G, K, K, W, K, C, N, aux, N, r1s, r1w);
  // ReLU masks travel as bits (32 B instead of 1 KB per sample and layer)
  return tcg_rows(st, EPI == EPI_MASK ? TCG_MASKBITS : EPI, true, m, N, G, K, K, nullptr, 0, 0, W, K, C, N, nullptr, aux, N,
                  r1s, r1w, a_amax, nullptr, c_amax, mask_in, nullptr);
}
// dW[M,N] += H[m,M]^T G[m,N]; db[N] += column sums of G (nullable)
static int dense_dw(cudaStream_t st, int M, int N, const float* H, const float* G, int64_t m, float* dW, float* db,
                    const float* h_amax, const float* g_amax) {
  if (g_fp32_ffma) {
    const int rc = gemm_tn_acc(st, M, N, H, M, G, N, m, dW, N);
    if (rc || db == nullptr) return rc;
    colsum_kernel<><<<ew_blocks(m, 512), 256, 0, st>>>(G, m, N, db);
    LNRF_LAUNCH_CHECK("colsum_kernel");
    return LNRF_OK;
  }
  return tcg_tn_acc(st, M, N, H, M, G, N, m, dW, N, db, h_amax, g_amax);
}

int nerf_fwd_fp32(const float* P, const float* x, const float* d, const float* rays, const float* ts,
                  int64_t m, int T, bool save, void* ws_base, int64_t ws_bytes, float* dens,
                  float* rgb, cudaStream_t st) {
  LNRF_REQUIRE(ws_bytes >= fp32_workspace_bytes(m, save), LNRF_E_WORKSPACE,
               "lnrf_nerf_mlp_fwd(fp32): workspace %lld < %lld bytes", (long long)ws_bytes,
               (long long)fp32_workspace_bytes(m, save));
  Fp32Ws w = carve_fp32(ws_base, m, save);
  LNRF_CUDA(cudaMemsetAsync(w.amax, 0, 32 * sizeof(float), st));
  embed_kernel<kXFreqs><<<ew_blocks(m * 3 * kXFreqs, 256), 256, 0, st>>>(x, rays, ts, T, 0, m, w.xe);
  LNRF_LAUNCH_CHECK("embed_kernel<x>");
  embed_kernel<kDFreqs><<<ew_blocks(m * 3 * kDFreqs, 256), 256, 0, st>>>(d, rays, ts, T, 1, m, w.de);
  LNRF_LAUNCH_CHECK("embed_kernel<d>");
  int rc;
  float* am = w.amax;  // am[l] = max |h_l| (encodings are O(1): no scale)
  // input stack, model.py:50-51
  if ((rc = dense_fwd<EPI_BIAS_RELU>(st, m, kH, w.xe, kXE, kXE, nullptr, 0, 0, P + kNerf.w[0], w.h[0], P + kNerf.b[0],
                                     nullptr, am + 0, save ? w.mask[0] : nullptr)))
    return rc;
  for (int l = 1; l <= 4; ++l)
    if ((rc = dense_fwd<EPI_BIAS_RELU>(st, m, kH, w.h[l - 1], kH, kH, nullptr, 0, 0, P + kNerf.w[l], w.h[l],
                                       P + kNerf.b[l], am + l - 1, am + l, save ? w.mask[l] : nullptr)))
      return rc;
  // skip concat [z | x_emb], model.py:52; Dense_5..7 outputs are consumed through ReLU (:53-56)
  if ((rc = dense_fwd<EPI_BIAS_RELU>(st, m, kH, w.h[4], kH, kH, w.xe, kXE, kXE, P + kNerf.w[5], w.h[5], P + kNerf.b[5],
                                     nullptr, am + 5, save ? w.mask[5] : nullptr)))
    return rc;
  for (int l = 6; l <= 7; ++l)
    if ((rc = dense_fwd<EPI_BIAS_RELU>(st, m, kH, w.h[l - 1], kH, kH, nullptr, 0, 0, P + kNerf.w[l], w.h[l],
                                       P + kNerf.b[l], am + l - 1, am + l, save ? w.mask[l] : nullptr)))
      return rc;
  // Dense_8 output z is used raw by both heads (:57-58)
  if ((rc = dense_fwd<EPI_BIAS>(st, m, kH, w.h[7], kH, kH, nullptr, 0, 0, P + kNerf.w[8], w.h[8], P + kNerf.b[8],
                                am + 7, am + 8)))
    return rc;
  density_head_fwd_kernel<<<ew_blocks(m, 8), 256, 0, st>>>(w.h[8], P + kNerf.w[9], P + kNerf.b[9], m,
                                                           dens);
  LNRF_LAUNCH_CHECK("density_head_fwd_kernel");
  if ((rc = dense_fwd<EPI_BIAS_RELU>(st, m, kHC, w.h[8], kH, kH, w.de, kDE, kDE, P + kNerf.w[10], w.c, P + kNerf.b[10],
                                     nullptr, am + kAmaxC)))
    return rc;
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
  float* am = w.amax;
  LNRF_CUDA(cudaMemsetAsync(am + kAmaxDc, 0, (32 - kAmaxDc) * sizeof(float), st));
  const unsigned rb = ew_blocks(m, 8 * 16);  // fewer, longer-lived blocks: less atomic traffic
  rgb_head_bwd_kernel<<<rb, 256, 0, st>>>(w.c, rgb, d_rgb, P + kNerf.w[11], m, w.dc, G + kNerf.w[11],
                                          G + kNerf.b[11], am + kAmaxDc);
  LNRF_LAUNCH_CHECK("rgb_head_bwd_kernel");
  density_head_bwd_kernel<<<rb, 256, 0, st>>>(w.h[8], dens, d_dens, m, w.spre, G + kNerf.w[9],
                                              G + kNerf.b[9]);
  LNRF_LAUNCH_CHECK("density_head_bwd_kernel");
  int rc;
  // colour layer Dense_10: input [z8 | d_emb]
  if ((rc = dense_dw(st, kH, kHC, w.h[8], w.dc, m, G + kNerf.w[10], G + kNerf.b[10], am + 8, am + kAmaxDc))) return rc;
  if ((rc = dense_dw(st, kDE, kHC, w.de, w.dc, m, G + kNerf.w[10] + int64_t(kH) * kHC, nullptr, nullptr, am + kAmaxDc)))
    return rc;
  // g8 = dc @ W10[:256]^T + spre (x) w9
  float* g = w.gA;
  float* gn = w.gB;
  if ((rc = dense_dx<EPI_RANK1>(st, m, kH, w.dc, kHC, P + kNerf.w[10], g, nullptr, w.spre, P + kNerf.w[9], am + kAmaxDc,
                                am + kAmaxG)))
    return rc;
  for (int l = 8; l >= 1; --l) {
    // dW_l = in_l^T g_l ; db_l = colsum(g_l) ; g_{l-1} = (g_l @ W_l[:256]^T) * (h_{l-1} > 0)
    const float* ga = am + kAmaxG + (8 - l);
    if ((rc = dense_dw(st, kH, kH, w.h[l - 1], g, m, G + kNerf.w[l], G + kNerf.b[l], am + l - 1, ga))) return rc;
    if (l == 5 && (rc = dense_dw(st, kXE, kH, w.xe, g, m, G + kNerf.w[5] + int64_t(kH) * kH, nullptr, nullptr, ga)))
      return rc;
    if ((rc = dense_dx<EPI_MASK>(st, m, kH, g, kH, P + kNerf.w[l], gn, w.h[l - 1], nullptr, nullptr, ga,
                                 am + kAmaxG + (9 - l), w.mask[l - 1])))
      return rc;
    float* t = g; g = gn; gn = t;
  }
  return dense_dw(st, kXE, kH, w.xe, g, m, G + kNerf.w[0], G + kNerf.b[0], nullptr, am + kAmaxG + 8);
}

}  // namespace lnrf
