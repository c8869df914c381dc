// K9: Ref-NeRF (learn_nerf/ref_nerf.py:34-107, sh_degree = 4) forward and backward, fp32 path.
//
// RefNERFBase.__call__ differentiates the spatial MLP w.r.t. its INPUT inside the forward
// (real_normal = normalize(d(-out[:,0])/dx), ref_nerf.py:38-43) and training differentiates
// through that.  With ReLU networks both passes are chains of GEMMs over the samples:
//
//   forward    h_l = relu(h_{l-1} W_l + b_l)                                   (spatial_block)
//   normals    Gn_7 = -W_8[:,0] * [h_7>0];  Gn_{l-1} = (Gn_l W_l^T) * [h_{l-1}>0]   (VJP of -z8[0])
//              d x_emb = Gn_0 W_0^T + Gn_5 W_5[256:]^T;  n_raw = (d emb/dx)^T d x_emb
//   backward   standard chain G_l from dL/dz8, dW_l += h_{l-1}^T G_l, plus the second-order
//              term through n_raw: with u = dL/dn_raw, the tangent pass
//              T_emb = (d emb/dx) u,  T_l = (T_{l-1} W_l) * [h_l>0]  gives
//              dW_l += T_{l-1}^T Gn_l  and  dW_8[:,0] -= sum T_7      (biases get nothing).
//
// All contractions run on the fp32 FFMA GEMM of sgemm.cuh; the per-sample head arithmetic
// (activations, reflection, integrated directional encoding, sRGB, aux losses) is elementwise.
#include "embed.cuh"
#include "lnrf_common.cuh"
#include "lnrf_math.cuh"
#include "gemm_tc.cuh"
#include "nerf_layout.cuh"
#include "sgemm.cuh"

namespace lnrf {

// ---- the contractions: split-fp16 tcgen05 GEMMs (gemm_tc.cu), or the FFMA GEMMs of sgemm.cuh when
// LNRF_FP32_FFMA=1 was set at lnrf_init (A/B measurements).  `*_amax` are device floats holding max |operand|
// (see gemm_tc.cuh); nullptr = the operand is O(1).
bool fp32_ffma();  // mlp_fp32.cu
template <int EPI>
static int rg_nn(cudaStream_t st, int64_t m, int N, const float* A0, int lda0, int K0, const float* A1, int lda1, int K1,
                 const float* W, float* C, const float* bias, const float* aux, const float* a_amax,
                 const float* a1_amax, float* c_amax, const uint32_t* mask_in = nullptr, uint32_t* mask_out = nullptr) {
  if (fp32_ffma() || !tcg_supported(N, K0, K1)) return gemm_nn<EPI>(st, m, N, A0, lda0, K0, A1, lda1, K1, W, N, C, N, bias, aux, N);
  // ReLU masks travel as bits when the caller has them (32 B instead of 1 KB per sample and layer)
  return tcg_rows(st, EPI == EPI_MASK && mask_in ? TCG_MASKBITS : EPI, false, m, N, A0, lda0, K0, A1, lda1, K1, W, N, C, N,
                  bias, aux, N, nullptr, nullptr, a_amax, a1_amax, c_amax, mask_in, mask_out);
}
// C[m,N] = epi(Gr[m,K] @ W[N rows, K cols]^T)
template <int EPI>
static int rg_nt(cudaStream_t st, int64_t m, int N, const float* Gr, int K, const float* W, float* C, const float* aux,
                 const float* a_amax, float* c_amax, const uint32_t* mask_in = nullptr) {
  if (fp32_ffma() || !tcg_supported(N, K, 0)) return gemm_nt<EPI>(st, m, N, Gr, K, K, W, K, C, N, aux, N);
  return tcg_rows(st, EPI == EPI_MASK && mask_in ? TCG_MASKBITS : EPI, true, m, N, Gr, K, K, nullptr, 0, 0, W, K, C, N,
                  nullptr, aux, N, nullptr, nullptr, a_amax, nullptr, c_amax, mask_in, nullptr);
}
// dW[M,N] += H[m,M]^T Gr[m,N]; db[N] += column sums of Gr (nullable)
static int rg_tn(cudaStream_t st, int M, int N, const float* H, const float* Gr, int64_t m, float* dW, float* db,
                 const float* h_amax, const float* g_amax) {
  if (fp32_ffma()) {
    const int rc = gemm_tn_acc(st, M, N, H, M, Gr, N, m, dW, N);
    if (rc || db == nullptr) return rc;
    colsum_kernel<><<<ew_blocks(m, 512), 256, 0, st>>>(Gr, m, N, db);
    LNRF_LAUNCH_CHECK("colsum_kernel");
    return LNRF_OK;
  }
  return tcg_tn_acc(st, M, N, H, M, Gr, N, m, dW, N, db, h_amax, g_amax);
}
// amax slots of the Ref-NeRF workspace (floats at RefWs::amax): forward 0..18, backward 19..
constexpr int kRaH = 0, kRaGn = 9, kRaC = 17, kRaW8 = 18, kRaFwdEnd = 19;
constexpr int kRaGc = 19, kRaG8pre = 20, kRaG = 21 /* g8 .. g0: 9 */, kRaTemb = 30, kRaT = 31 /* T0 .. T7 */, kRaEnd = 40;

constexpr int kRefLayers = 11;
constexpr int kRefEnc = 16;            // sum(HARMONIC_COUNTS[:4])
constexpr int kRefDirIn = kH + kRefEnc + 1;  // 273 (ref_nerf.py:63)
constexpr int kRefE = 20;              // [IDE(16) | n.(-d) | 3 zero pads]: second K segment of Dense_9
constexpr int kRefDirPad = kH + kRefE; // 276 kernel rows in the flat buffer (rows 273..275 stay zero)

struct RefLayout {
  int in[kRefLayers], out[kRefLayers];
  int64_t w[kRefLayers], b[kRefLayers], total;
};
constexpr RefLayout make_ref_layout() {
  RefLayout L{};
  const int ins[kRefLayers] = {kXE, kH, kH, kH, kH, kH + kXE, kH, kH, kH, kRefDirIn, kHC};
  const int rows[kRefLayers] = {kXE, kH, kH, kH, kH, kH + kXE, kH, kH, kH, kRefDirPad, kHC};
  const int outs[kRefLayers] = {kH, kH, kH, kH, kH, kH, kH, kH, kH, kHC, 3};
  int64_t off = 0;
  for (int i = 0; i < kRefLayers; ++i) {
    L.in[i] = ins[i];
    L.out[i] = outs[i];
    L.w[i] = off;
    off += int64_t(rows[i]) * outs[i];
    off = (off + 3) / 4 * 4;
    L.b[i] = off;
    off += outs[i];
    off = (off + 3) / 4 * 4;
  }
  L.total = off;
  return L;
}
constexpr RefLayout kRef = make_ref_layout();

// ---------------------------------------------------------------- spherical harmonics, degree 4
// ref_nerf.py:146-195 (tiny-cuda-nn constants); level of term k: 0 | 1,1,1 | 2 x5 | 3 x7.
__device__ __forceinline__ void sh16(float x, float y, float z, float* o) {
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = -0.48860251190291987f * y;
  o[2] = 0.48860251190291987f * z;
  o[3] = -0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy;
  o[5] = -1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz;
  o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
  o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
  o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
  o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}
// g += J^T w  for the 16 terms above
__device__ __forceinline__ void sh16_bwd(float x, float y, float z, const float* w, float* g) {
  const float c1 = 0.48860251190291987f, c2 = 1.0925484305920792f, c3 = 0.94617469575755997f;
  const float c5 = 0.54627421529603959f, c6 = 0.59004358992664352f, c7 = 2.8906114426405538f;
  const float c8 = 0.45704579946446572f, c9 = 0.3731763325901154f, c10 = 1.4453057213202769f;
  const float x2 = x * x, y2 = y * y, z2 = z * z;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  gy += -c1 * w[1];
  gz += c1 * w[2];
  gx += -c1 * w[3];
  gx += c2 * y * w[4];  gy += c2 * x * w[4];
  gy += -c2 * z * w[5]; gz += -c2 * y * w[5];
  gz += 2.0f * c3 * z * w[6];
  gx += -c2 * z * w[7]; gz += -c2 * x * w[7];
  gx += 2.0f * c5 * x * w[8]; gy += -2.0f * c5 * y * w[8];
  gx += c6 * (-6.0f * x * y) * w[9]; gy += c6 * (-3.0f * x2 + 3.0f * y2) * w[9];
  gx += c7 * y * z * w[10]; gy += c7 * x * z * w[10]; gz += c7 * x * y * w[10];
  gy += c8 * (1.0f - 5.0f * z2) * w[11]; gz += c8 * (-10.0f * y * z) * w[11];
  gz += c9 * (15.0f * z2 - 3.0f) * w[12];
  gx += c8 * (1.0f - 5.0f * z2) * w[13]; gz += c8 * (-10.0f * x * z) * w[13];
  gx += c10 * 2.0f * x * z * w[14]; gy += -c10 * 2.0f * y * z * w[14]; gz += c10 * (x2 - y2) * w[14];
  gx += c6 * (-3.0f * x2 + 3.0f * y2) * w[15]; gy += c6 * 6.0f * x * y * w[15];
  g[0] += gx; g[1] += gy; g[2] += gz;
}
__device__ __forceinline__ int sh_level(int k) { return k == 0 ? 0 : (k < 4 ? 1 : (k < 9 ? 2 : 3)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void load_dir(const float* __restrict__ d, const float* __restrict__ rays, int T,
                                         int64_t s, float dv[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) dv[a] = d ? __ldg(d + s * 3 + a) : __ldg(rays + (s / T) * 6 + 3 + a);
}
__device__ __forceinline__ void load_pos(const float* __restrict__ x, const float* __restrict__ rays,
                                         const float* __restrict__ ts, int T, int64_t s, float p[3]) {
  if (x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = __ldg(x + s * 3 + a);
  } else {
    const int64_t r = s / T;
    const float t = __ldg(ts + s);
#pragma unroll
    for (int a = 0; a < 3; ++a)
      p[a] = __fadd_rn(__ldg(rays + r * 6 + a), __fmul_rn(__ldg(rays + r * 6 + 3 + a), t));  // render.py:153
  }
}

// ---------------------------------------------------------------- normal chain helpers
// Gn_7[s, j] = -W_8[j, 0] * [h_7[s, j] > 0]    (seed of the VJP of -z8[:, 0])
__global__ void __launch_bounds__(256)
ref_seed_kernel(const float* __restrict__ h7, const float* __restrict__ w8, int64_t m, float* __restrict__ gn7,
                int width = kH, int ldw = kH) {  // width = hidden units (power of two), ldw = columns of the last kernel
  const int64_t total = m * width;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int j = int(i & (width - 1));
    gn7[i] = __ldg(h7 + i) > 0.0f ? -__ldg(w8 + int64_t(j) * ldw) : 0.0f;
  }
}

// n_raw = (d emb / d x)^T (dxe0 + dxe5): emb[dim*20 + f] = sin(2^f x), emb[dim*20 + 10 + f] = cos(2^f x)
// (model.py:65-77).  One thread per (sample, dim).  Output [m,4] (3 used).
__global__ void __launch_bounds__(256)
ref_nraw_kernel(const float* __restrict__ x, const float* __restrict__ rays, const float* __restrict__ ts, int T,
                int64_t m, const float* __restrict__ dxe0, const float* __restrict__ dxe5,
                float* __restrict__ nraw) {
  const int64_t total = m * 3;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t s = i / 3;
    const int dim = int(i - s * 3);
    float p[3];
    load_pos(x, rays, ts, T, s, p);
    float acc = 0.0f;
#pragma unroll
    for (int f = 0; f < kXFreqs; ++f) {
      const float c = float(1 << f);
      float sn, cs;
      sincosf(p[dim] * c, &sn, &cs);
      const int64_t o = s * kXE + dim * 2 * kXFreqs + f;
      const float gs = __ldg(dxe0 + o) + __ldg(dxe5 + o);
      const float gc = __ldg(dxe0 + o + kXFreqs) + __ldg(dxe5 + o + kXFreqs);
      acc += c * (cs * gs - sn * gc);
    }
    nraw[s * 4 + dim] = acc;
  }
}

// T_emb = (d emb / d x) u : the forward-mode tangent of the embedding along u[m,4].
__global__ void __launch_bounds__(256)
ref_temb_kernel(const float* __restrict__ x, const float* __restrict__ rays, const float* __restrict__ ts, int T,
                int64_t m, const float* __restrict__ u, float* __restrict__ temb) {
  const int64_t total = m * 3 * kXFreqs;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t s = i / (3 * kXFreqs);
    const int rem = int(i - s * 3 * kXFreqs);
    const int dim = rem / kXFreqs, f = rem - dim * kXFreqs;
    float p[3];
    load_pos(x, rays, ts, T, s, p);
    const float c = float(1 << f);
    float sn, cs;
    sincosf(p[dim] * c, &sn, &cs);
    const float ud = __ldg(u + s * 4 + dim);
    temb[s * kXE + dim * 2 * kXFreqs + f] = c * cs * ud;
    temb[s * kXE + dim * 2 * kXFreqs + kXFreqs + f] = -c * sn * ud;
  }
}

// dW_8[:, 0] -= sum_s T_7[s, :]   (the tangent network's last layer is column 0 of Dense_8)
__global__ void __launch_bounds__(256)
ref_w8col_kernel(const float* __restrict__ t7, int64_t m, float* __restrict__ dw8, int ldw = kH) {
  const int col = threadIdx.x;  // one thread per input unit (blockDim.x = hidden width)
  const int width = blockDim.x;
  const int64_t rows_per_block = ceil_div(m, gridDim.x);
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, m);
  float acc = 0.0f;
  for (int64_t r = r0; r < r1; ++r) acc += __ldg(t7 + r * width + col);
  atomicAdd(dw8 + int64_t(col) * ldw, -acc);
}

// ---------------------------------------------------------------- per-sample head
struct RefHead {  // everything RefNERFBase.__call__ derives from z8[:, :9], d and n_raw (ref_nerf.py:45-75)
  float density, dif[3], spec, rough, n[3], ninv, nn2, dn, refl[3], sh[16], att[4], rn[3], rninv, rn2;
};
__device__ __forceinline__ RefHead ref_head(const float z[9], const float d[3], const float nr[3]) {
  RefHead h;
  h.density = expf(z[0]);                                     // :48
#pragma unroll
  for (int i = 0; i < 3; ++i) h.dif[i] = sigmoid_f(z[1 + i] - 1.0986122886681098f);  // :52 (log 3)
  h.spec = sigmoid_f(z[4]);                                   // :54
  h.rough = softplus_f(z[5]);                                 // :55
  h.nn2 = z[6] * z[6] + z[7] * z[7] + z[8] * z[8];
  h.ninv = 1.0f / sqrtf(h.nn2 + 1e-10f);                      // :56, :314-317
#pragma unroll
  for (int i = 0; i < 3; ++i) h.n[i] = z[6 + i] * h.ninv;
  h.dn = d[0] * h.n[0] + d[1] * h.n[1] + d[2] * h.n[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) h.refl[i] = d[i] - 2.0f * h.n[i] * h.dn;  // :58
  sh16(h.refl[0], h.refl[1], h.refl[2], h.sh);
#pragma unroll
  for (int l = 0; l < 4; ++l) h.att[l] = expf(-h.rough * float(l * (l + 1)) * 0.5f);  // :141
  h.rn2 = nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2];
  h.rninv = 1.0f / sqrtf(h.rn2 + 1e-10f);                     // :43
#pragma unroll
  for (int i = 0; i < 3; ++i) h.rn[i] = nr[i] * h.rninv;
  return h;
}

// forward part 1: density, the directional block's extra inputs E = [IDE(16) | n.(-d) | 0 0 0], aux losses
__global__ void __launch_bounds__(256)
ref_head_fwd1_kernel(const float* __restrict__ z8, const float* __restrict__ d, const float* __restrict__ rays,
                     int T, const float* __restrict__ nraw, int64_t m, float* __restrict__ dens,
                     float* __restrict__ E, float* __restrict__ aux_mse, float* __restrict__ aux_neg, int ldz = kH) {
  for (int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < m; s += int64_t(gridDim.x) * blockDim.x) {
    float z[9], dv[3], nr[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) z[i] = __ldg(z8 + s * ldz + i);
    load_dir(d, rays, T, s, dv);
#pragma unroll
    for (int i = 0; i < 3; ++i) nr[i] = __ldg(nraw + s * 4 + i);
    const RefHead h = ref_head(z, dv, nr);
    dens[s] = h.density;
    float* e = E + s * kRefE;
#pragma unroll
    for (int k = 0; k < 16; ++k) e[k] = h.sh[k] * h.att[sh_level(k)];  // :142-143
    e[16] = -h.dn;                                                      // :62
    e[17] = 0.f; e[18] = 0.f; e[19] = 0.f;
    float mse = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) mse += (h.n[i] - h.rn[i]) * (h.n[i] - h.rn[i]);
    aux_mse[s] = mse;                                                   // :73
    const float pos = fmaxf(0.0f, h.dn);
    aux_neg[s] = pos * pos;                                             // :74
  }
}

// sRGB gamma (:110-118) and its derivative on the clipped colour
__device__ __forceinline__ float srgb_f(float c) {
  return c <= 0.0031308f ? 12.92f * c : 1.055f * powf(fmaxf(1e-5f, c), 1.0f / 2.4f) - 0.055f;
}
__device__ __forceinline__ float srgb_df(float c) {
  return c <= 0.0031308f ? 12.92f : (1.055f / 2.4f) * powf(c, 1.0f / 2.4f - 1.0f);
}

// forward part 2: o = directional_block output [m,4] -> rgb (:65-71)
__global__ void __launch_bounds__(256)
ref_head_fwd2_kernel(const float* __restrict__ z8, const float* __restrict__ o, int64_t m, float* __restrict__ rgb,
                     int ldz = kH) {
  for (int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < m; s += int64_t(gridDim.x) * blockDim.x) {
    const float spec = sigmoid_f(__ldg(z8 + s * ldz + 4));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float dif = sigmoid_f(__ldg(z8 + s * ldz + 1 + i) - 1.0986122886681098f);
      const float lin = sigmoid_f(__ldg(o + s * 4 + i)) * spec + dif;
      const float cl = fminf(fmaxf(lin, 0.0f), 1.0f);  // _leaky_clip forward value (:320-326)
      rgb[s * 3 + i] = srgb_f(cl) * 2.0f - 1.0f;
    }
  }
}

// Dense_10 (128 -> 3) / the 64 -> 3 layer of the hash-grid variant: o = c @ W + b, warp per sample
// (o is [m,4], 3 used); a lane owns PER = width / 32 consecutive hidden units.
__device__ __forceinline__ void load_per(const float* p, float (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load_per(const float* p, float (&v)[2]) {
  const float2 a = __ldg(reinterpret_cast<const float2*>(p));
  v[0] = a.x; v[1] = a.y;
}
__device__ __forceinline__ void store_per(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store_per(float* p, const float (&v)[2]) {
  *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}
template <int PER>
__global__ void __launch_bounds__(256)
ref_out_fwd_kernel(const float* __restrict__ c, const float* __restrict__ w10, const float* __restrict__ b10,
                   int64_t m, float* __restrict__ o) {
  constexpr int kW = PER * 32;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float w[PER][3];
#pragma unroll
  for (int k = 0; k < PER; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) w[k][j] = __ldg(w10 + (lane * PER + k) * 3 + j);
  for (int64_t s = warp; s < m; s += nwarps) {
    float av[PER];
    load_per(c + s * kW + lane * PER, av);
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < PER; ++k)
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[j] = fmaf(av[k], w[k][j], acc[j]);
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[j] = warp_sum(acc[j]);
    if (lane < 3) o[s * 4 + lane] = (lane == 0 ? acc[0] : (lane == 1 ? acc[1] : acc[2])) + __ldg(b10 + lane);
  }
}

// backward part 1: d o (gradient w.r.t. the directional block's output) from d rgb; then
// Dense_10 backward: gc = (d_o @ W10^T) * [c > 0], dW10 += c^T d_o, db10 += sum d_o.
template <int PER>
__global__ void __launch_bounds__(256)
ref_out_bwd_kernel(const float* __restrict__ z8, const float* __restrict__ o, const float* __restrict__ c,
                   const float* __restrict__ d_rgb, const float* __restrict__ w10, int64_t m,
                   float* __restrict__ d_o, float* __restrict__ gc, float* __restrict__ dw10,
                   float* __restrict__ db10, int ldz) {
  constexpr int kW = PER * 32;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float w[PER][3], gw[PER][3], gb[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < PER; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      w[k][j] = __ldg(w10 + (lane * PER + k) * 3 + j);
      gw[k][j] = 0.0f;
    }
  // A warp takes 32 samples at a time: each lane first forms d_o of ITS sample (the sigmoids and the sRGB
  // derivative once per sample, not once per lane), then the warp streams the 32 feature rows with d_o
  // broadcast by shuffles.
  for (int64_t base = warp * 32; base < m; base += nwarps * 32) {
    const int64_t s = base + lane;
    float dov[3] = {0.f, 0.f, 0.f};
    if (s < m) {
      const float spec = sigmoid_f(__ldg(z8 + s * ldz + 4));
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float dif = sigmoid_f(__ldg(z8 + s * ldz + 1 + i) - 1.0986122886681098f);
        const float sc = sigmoid_f(__ldg(o + s * 4 + i));
        const float lin = sc * spec + dif;
        const float cl = fminf(fmaxf(lin, 0.0f), 1.0f);
        const float dlin = 2.0f * __ldg(d_rgb + s * 3 + i) * srgb_df(cl);  // straight-through clip
        dov[i] = dlin * spec * sc * (1.0f - sc);
        gb[i] += dov[i];
        d_o[s * 4 + i] = dov[i];
      }
    }
    const int cnt = m - base < 32 ? int(m - base) : 32;
#pragma unroll 4
    for (int t = 0; t < cnt; ++t) {
      const float d0 = __shfl_sync(0xffffffffu, dov[0], t), d1 = __shfl_sync(0xffffffffu, dov[1], t),
                  d2 = __shfl_sync(0xffffffffu, dov[2], t);
      const int64_t st = base + t;
      float av[PER];
      load_per(c + st * kW + lane * PER, av);
      float g4[PER];
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const float tt = d0 * w[k][0] + d1 * w[k][1] + d2 * w[k][2];
        g4[k] = av[k] > 0.0f ? tt : 0.0f;
        gw[k][0] = fmaf(av[k], d0, gw[k][0]);
        gw[k][1] = fmaf(av[k], d1, gw[k][1]);
        gw[k][2] = fmaf(av[k], d2, gw[k][2]);
      }
      store_per(gc + st * kW + lane * PER, g4);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) gb[i] = warp_sum(gb[i]);
  __shared__ float s_gw[8][kW * 3];
#pragma unroll
  for (int k = 0; k < PER; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) s_gw[wib][(lane * PER + k) * 3 + j] = gw[k][j];
  __syncthreads();
  for (int i = threadIdx.x; i < kW * 3; i += blockDim.x) {
    float t = 0.0f;
    for (int ww = 0; ww < 8; ++ww) t += s_gw[ww][i];
    atomicAdd(dw10 + i, t);
  }
  if (lane == 0) {
    atomicAdd(db10 + 0, gb[0]);
    atomicAdd(db10 + 1, gb[1]);
    atomicAdd(db10 + 2, gb[2]);
  }
}

// backward part 2: everything between z8[:, :9] / n_raw and (density, E, aux losses, colour mix).
// Adds dL/dz8[:, :9] into g8 (which already holds gc @ W9[:256]^T) and writes u = dL/dn_raw [m,4].
__global__ void __launch_bounds__(256)
ref_head_bwd_kernel(const float* __restrict__ z8, const float* __restrict__ d, const float* __restrict__ rays,
                    int T, const float* __restrict__ nraw, const float* __restrict__ o,
                    const float* __restrict__ d_dens, const float* __restrict__ d_rgb,
                    const float* __restrict__ d_mse, const float* __restrict__ d_neg,
                    const float* __restrict__ dE, int64_t m, float* __restrict__ g8, float* __restrict__ u,
                    int ldz = kH) {
  for (int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < m; s += int64_t(gridDim.x) * blockDim.x) {
    float z[9], dv[3], nr[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) z[i] = __ldg(z8 + s * ldz + i);
    load_dir(d, rays, T, s, dv);
#pragma unroll
    for (int i = 0; i < 3; ++i) nr[i] = __ldg(nraw + s * 4 + i);
    const RefHead h = ref_head(z, dv, nr);
    float dz[9];
    dz[0] = __ldg(d_dens + s) * h.density;
    // colour mix: lin = sigmoid(o) * spec + dif
    float dspec = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float sc = sigmoid_f(__ldg(o + s * 4 + i));
      const float lin = sc * h.spec + h.dif[i];
      const float cl = fminf(fmaxf(lin, 0.0f), 1.0f);
      const float dlin = 2.0f * __ldg(d_rgb + s * 3 + i) * srgb_df(cl);
      dspec += dlin * sc;
      dz[1 + i] = dlin * h.dif[i] * (1.0f - h.dif[i]);
    }
    dz[4] = dspec * h.spec * (1.0f - h.spec);
    // integrated directional encoding: E_k = sh_k(refl) * att_level(k)
    const float* de = dE + s * kRefE;
    float drough = 0.0f, dsh[16], drefl[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int l = sh_level(k);
      const float g = __ldg(de + k);
      drough += g * h.sh[k] * h.att[l] * (-0.5f * float(l * (l + 1)));
      dsh[k] = g * h.att[l];
    }
    dz[5] = drough * sigmoid_f(z[5]);  // softplus'
    sh16_bwd(h.refl[0], h.refl[1], h.refl[2], dsh, drefl);
    // refl = d - 2 n (d.n);  ndot = -(d.n);  neg_normal = max(0, d.n)^2;  normal_mse = |n - rn|^2
    float dn_vec[3], ddn = -__ldg(de + 16);
    const float n_dot_drefl = h.n[0] * drefl[0] + h.n[1] * drefl[1] + h.n[2] * drefl[2];
    ddn += -2.0f * n_dot_drefl + __ldg(d_neg + s) * 2.0f * fmaxf(0.0f, h.dn);
    const float gm = __ldg(d_mse + s);
    float drn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float diff = h.n[i] - h.rn[i];
      dn_vec[i] = -2.0f * h.dn * drefl[i] + ddn * dv[i] + 2.0f * gm * diff;
      drn[i] = -2.0f * gm * diff;
    }
    // n = v / sqrt(|v|^2 + eps): dv_j = dn_j * inv - v_j (dn . v) inv^3
    {
      const float dot = dn_vec[0] * z[6] + dn_vec[1] * z[7] + dn_vec[2] * z[8];
      const float inv3 = h.ninv * h.ninv * h.ninv;
#pragma unroll
      for (int i = 0; i < 3; ++i) dz[6 + i] = dn_vec[i] * h.ninv - z[6 + i] * dot * inv3;
    }
    {
      const float dot = drn[0] * nr[0] + drn[1] * nr[1] + drn[2] * nr[2];
      const float inv3 = h.rninv * h.rninv * h.rninv;
#pragma unroll
      for (int i = 0; i < 3; ++i) u[s * 4 + i] = drn[i] * h.rninv - nr[i] * dot * inv3;
      u[s * 4 + 3] = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) g8[s * ldz + i] += dz[i];
  }
}

// ---------------------------------------------------------------- workspace
struct RefWs {
  float* xe;      // [m,60]
  float* h[9];    // h0..h7 post-ReLU, h[8] = z8
  float* gn[8];   // normal-chain gradients Gn_0..Gn_7 (aliased ping-pong when !save)
  float* dxe0;    // [m,60]  Gn_0 W_0^T
  float* dxe5;    // [m,60]  Gn_5 W_5[256:]^T
  float* nraw;    // [m,4]
  float* E;       // [m,20]
  float* c;       // [m,128]
  float* o;       // [m,4]
  // backward only
  float *gA, *gB, *tA, *tB, *temb, *d_o, *gc, *dE, *u;
  float* amax;    // [64] max |operand| slots of the tensor-core GEMMs (kRa*)
  uint32_t* mask[8];  // ReLU bit masks of h0..h7 (written by the forward GEMM epilogues)
  int64_t bytes;
};
static RefWs carve_ref(void* base, int64_t m, bool save) {
  RefWs w{};
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t floats) {
    float* r = reinterpret_cast<float*>(p + off);
    off += align_up(floats * 4, 256);
    return r;
  };
  w.xe = take(m * kXE);
  for (int i = 0; i < 9; ++i) w.h[i] = take(m * kH);  // the normal chain needs every ReLU mask
  if (save) {
    for (int i = 0; i < 8; ++i) w.gn[i] = take(m * kH);
  } else {
    float* a = take(m * kH);
    float* b = take(m * kH);
    for (int i = 0; i < 8; ++i) w.gn[i] = (i & 1) ? b : a;
  }
  w.dxe0 = take(m * kXE);
  w.dxe5 = take(m * kXE);
  w.nraw = take(m * 4);
  w.E = take(m * kRefE);
  w.c = take(m * kHC);
  w.o = take(m * 4);
  w.amax = take(64);
  for (int i = 0; i < 8; ++i) w.mask[i] = reinterpret_cast<uint32_t*>(take(tcg_mask_words(m, kH)));
  if (save) {
    w.gA = take(m * kH);
    w.gB = take(m * kH);
    w.tA = take(m * kH);
    w.tB = take(m * kH);
    w.temb = take(m * kXE);
    w.d_o = take(m * 4);
    w.gc = take(m * kHC);
    w.dE = take(m * kRefE);
    w.u = take(m * 4);
  }
  w.bytes = off;
  return w;
}

// ================================================================ InstantNGPRefNERFModel
// instant_ngp.py:57-89 on RefNERFBase (ref_nerf.py:34-77): spatial_block = smooth multiresolution
// hash grid (E = 2L features) -> Dense_0 (E -> 64) ReLU -> Dense_1 (64 -> 16); the 16 outputs are
// split exactly like z8[:, :9] above; directional_block = Dense_2 (16 + 16 + 1 = 33 -> 64) ReLU ->
// Dense_3 (64 -> 64) ReLU -> Dense_4 (64 -> 3).  real_normal needs d(-out[:,0])/dx THROUGH the hash
// grid: vec = Gn_0 W_0^T with Gn_0 = -W_1[:,0] * [h_0 > 0], then n_raw = J^T vec with J = d enc / d x
// (hashgrid.cu).  The backward adds, along u = dL/dn_raw: T_enc = J u, dW_0 += T_enc^T Gn_0,
// dW_1[:,0] -= sum (T_enc W_0) * [h_0 > 0], and the table gradient (u . d w_c/dx) vec.
constexpr int kNrHidden = 64, kNrOut = 16;
constexpr int kNrDirRows = kNrOut + kRefE;  // 36 kernel rows in the flat buffer (33 used, 3 zero pads)
struct NgpRefLayout {
  int in[5], rows[5], out[5];
  int64_t w[5], b[5], total;
};
static NgpRefLayout ngpref_layout(int L) {
  NgpRefLayout n{};
  const int ins[5] = {2 * L, kNrHidden, kNrOut + kRefEnc + 1, kNrHidden, kNrHidden};
  const int rows[5] = {2 * L, kNrHidden, kNrDirRows, kNrHidden, kNrHidden};
  const int outs[5] = {kNrHidden, kNrOut, kNrHidden, kNrHidden, 3};
  int64_t off = 0;
  for (int i = 0; i < 5; ++i) {
    n.in[i] = ins[i];
    n.rows[i] = rows[i];
    n.out[i] = outs[i];
    n.w[i] = off;
    off = align_up(off + int64_t(rows[i]) * outs[i], 4);
    n.b[i] = off;
    off = align_up(off + outs[i], 4);
  }
  n.total = off;
  return n;
}

struct NgpRefWs {
  float *enc, *h0, *z, *gn0, *vec, *nraw, *E, *c1, *c2, *o;       // forward
  float *d_o, *gc2, *gc1, *g, *dE, *u, *g0, *d_enc, *tenc, *t0;   // backward only
  float* amax;                // [32] max |operand| slots of the tensor-core GEMMs (kNa*)
  uint32_t *mask_h0, *mask_c1;  // ReLU bit masks written by the forward GEMM epilogues
  int64_t bytes;
};
// amax slots: forward 0..7, backward 8..
constexpr int kNaGn0 = 0, kNaFwdEnd = 8, kNaGc2 = 8, kNaGc1 = 9, kNaG = 10, kNaG0 = 11, kNaTenc = 12;
static NgpRefWs carve_ngpref(void* base, int64_t m, int E, bool save) {
  NgpRefWs w{};
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t floats) {
    float* r = reinterpret_cast<float*>(p + off);
    off += align_up(floats * 4, 256);
    return r;
  };
  w.enc = take(m * E);
  w.h0 = take(m * kNrHidden);
  w.z = take(m * kNrOut);
  w.gn0 = take(m * kNrHidden);
  w.vec = take(m * E);
  w.nraw = take(m * 4);
  w.E = take(m * kRefE);
  w.c1 = take(m * kNrHidden);
  w.c2 = take(m * kNrHidden);
  w.o = take(m * 4);
  w.amax = take(32);
  w.mask_h0 = reinterpret_cast<uint32_t*>(take(tcg_mask_words(m, kNrHidden)));
  w.mask_c1 = reinterpret_cast<uint32_t*>(take(tcg_mask_words(m, kNrHidden)));
  if (save) {
    w.d_o = take(m * 4);
    w.gc2 = take(m * kNrHidden);
    w.gc1 = take(m * kNrHidden);
    w.g = take(m * kNrOut);
    w.dE = take(m * kRefE);
    w.u = take(m * 4);
    w.g0 = take(m * kNrHidden);
    w.d_enc = take(m * E);
    w.tenc = take(m * E);
    w.t0 = take(m * kNrHidden);
  }
  w.bytes = off;
  return w;
}

// hashgrid.cu: 0 = encode, 1 = scatter d_enc, 2 = J^T vec, 3 = J u + second-order table scatter
int hashgrid_launch(int which, const float* tables, const int64_t* level_offsets, const int32_t* grid_sizes,
                    const int32_t* table_sizes, int L, const float* bmin, const float* bmax, int smooth,
                    const float* x, const float* rays, const float* ts, int T, int64_t m, const float* in0,
                    const float* in1, float* out0, float* out1, cudaStream_t st, const float* in2 = nullptr);

}  // namespace lnrf

extern "C" {

int64_t lnrf_refnerf_param_count(void) {
  int64_t n = 0;
  for (int i = 0; i < lnrf::kRefLayers; ++i) n += int64_t(lnrf::kRef.in[i]) * lnrf::kRef.out[i] + lnrf::kRef.out[i];
  return n;
}
int64_t lnrf_refnerf_param_floats(void) { return lnrf::kRef.total; }
int lnrf_refnerf_param_offsets(int64_t* out_host) {
  LNRF_REQUIRE(out_host, LNRF_E_INVALID, "lnrf_refnerf_param_offsets: null pointer");
  for (int i = 0; i < lnrf::kRefLayers; ++i) {
    out_host[2 * i] = lnrf::kRef.w[i];
    out_host[2 * i + 1] = lnrf::kRef.b[i];
  }
  return LNRF_OK;
}
int lnrf_refnerf_workspace_bytes(int64_t m, int32_t save_for_backward, int64_t* bytes_out_host) {
  LNRF_REQUIRE(m >= 0 && bytes_out_host, LNRF_E_INVALID, "lnrf_refnerf_workspace_bytes: bad args");
  *bytes_out_host = lnrf::carve_ref(nullptr, m, save_for_backward != 0).bytes;
  return LNRF_OK;
}

int lnrf_refnerf_fwd(const float* params, const float* x, const float* d, const float* rays, const float* ts,
                     int64_t n, int32_t T, int32_t save_for_backward, void* workspace, int64_t workspace_bytes,
                     float* dens, float* rgb, float* aux_normal_mse, float* aux_neg_normal,
                     lnrf_stream_t stream) {
  using namespace lnrf;
  LNRF_REQUIRE(n >= 0 && T >= 1, LNRF_E_INVALID, "lnrf_refnerf_fwd: n=%lld T=%d", (long long)n, T);
  const int64_t m = n * T;
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(params && workspace && dens && rgb && aux_normal_mse && aux_neg_normal, LNRF_E_INVALID,
               "lnrf_refnerf_fwd: null pointer");
  LNRF_REQUIRE((x && d && !rays && !ts) || (!x && !d && rays && ts), LNRF_E_INVALID,
               "lnrf_refnerf_fwd: pass either (x,d) or (rays,ts)");
  LNRF_REQUIRE(m < (int64_t(1) << 31), LNRF_E_UNSUPPORTED, "lnrf_refnerf_fwd: %lld samples per call; chunk the batch",
               (long long)m);
  const bool save = save_for_backward != 0;
  LNRF_REQUIRE(workspace_bytes >= carve_ref(nullptr, m, save).bytes, LNRF_E_WORKSPACE,
               "lnrf_refnerf_fwd: workspace %lld < %lld bytes", (long long)workspace_bytes,
               (long long)carve_ref(nullptr, m, save).bytes);
  const RefWs w = carve_ref(workspace, m, save);
  cudaStream_t st = as_stream(stream);
  const float* P = params;
  int rc;
  float* am = w.amax;
  LNRF_CUDA(cudaMemsetAsync(am, 0, kRaFwdEnd * sizeof(float), st));
  // ---- spatial_block (ref_nerf.py:92-103)
  embed_kernel<kXFreqs><<<ew_blocks(m * 3 * kXFreqs, 256), 256, 0, st>>>(x, rays, ts, T, 0, m, w.xe);
  LNRF_LAUNCH_CHECK("embed_kernel<x>");
  if ((rc = rg_nn<EPI_BIAS_RELU>(st, m, kH, w.xe, kXE, kXE, nullptr, 0, 0, P + kRef.w[0], w.h[0], P + kRef.b[0], nullptr,
                                 nullptr, nullptr, am + kRaH + 0, nullptr, w.mask[0]))) return rc;
  for (int l = 1; l <= 4; ++l)
    if ((rc = rg_nn<EPI_BIAS_RELU>(st, m, kH, w.h[l - 1], kH, kH, nullptr, 0, 0, P + kRef.w[l], w.h[l], P + kRef.b[l],
                                   nullptr, am + kRaH + l - 1, nullptr, am + kRaH + l, nullptr, w.mask[l]))) return rc;
  if ((rc = rg_nn<EPI_BIAS_RELU>(st, m, kH, w.h[4], kH, kH, w.xe, kXE, kXE, P + kRef.w[5], w.h[5], P + kRef.b[5], nullptr,
                                 am + kRaH + 4, nullptr, am + kRaH + 5, nullptr, w.mask[5]))) return rc;
  for (int l = 6; l <= 7; ++l)
    if ((rc = rg_nn<EPI_BIAS_RELU>(st, m, kH, w.h[l - 1], kH, kH, nullptr, 0, 0, P + kRef.w[l], w.h[l], P + kRef.b[l],
                                   nullptr, am + kRaH + l - 1, nullptr, am + kRaH + l, nullptr, w.mask[l]))) return rc;
  if ((rc = rg_nn<EPI_BIAS>(st, m, kH, w.h[7], kH, kH, nullptr, 0, 0, P + kRef.w[8], w.h[8], P + kRef.b[8], nullptr,
                            am + kRaH + 7, nullptr, am + kRaH + 8))) return rc;
  // ---- real_normal: VJP of -z8[:, 0] w.r.t. x (ref_nerf.py:38-43)
  ref_seed_kernel<<<ew_blocks(m * kH, 256), 256, 0, st>>>(w.h[7], P + kRef.w[8], m, w.gn[7]);
  LNRF_LAUNCH_CHECK("ref_seed_kernel");
  // |Gn_7| <= max |W_8|: a bound is all the scale needs
  if ((rc = tcg_amax(st, P + kRef.w[8], int64_t(kH) * kH, am + kRaGn + 7))) return rc;
  for (int l = 7; l >= 1; --l) {
    if (l == 5)  // the skip input [z | x_emb]: rows 256.. of Dense_5 feed x_emb directly
      if ((rc = rg_nt<EPI_STORE>(st, m, kXE, w.gn[5], kH, P + kRef.w[5] + int64_t(kH) * kH, w.dxe5, nullptr,
                                 am + kRaGn + 5, nullptr))) return rc;
    if ((rc = rg_nt<EPI_MASK>(st, m, kH, w.gn[l], kH, P + kRef.w[l], w.gn[l - 1], w.h[l - 1], am + kRaGn + l,
                              am + kRaGn + l - 1, w.mask[l - 1]))) return rc;
  }
  if ((rc = rg_nt<EPI_STORE>(st, m, kXE, w.gn[0], kH, P + kRef.w[0], w.dxe0, nullptr, am + kRaGn + 0, nullptr))) return rc;
  ref_nraw_kernel<<<ew_blocks(m * 3, 256), 256, 0, st>>>(x, rays, ts, T, m, w.dxe0, w.dxe5, w.nraw);
  LNRF_LAUNCH_CHECK("ref_nraw_kernel");
  // ---- heads (ref_nerf.py:45-75)
  ref_head_fwd1_kernel<<<ew_blocks(m, 256), 256, 0, st>>>(w.h[8], d, rays, T, w.nraw, m, dens, w.E,
                                                          aux_normal_mse, aux_neg_normal);
  LNRF_LAUNCH_CHECK("ref_head_fwd1_kernel");
  if ((rc = rg_nn<EPI_BIAS_RELU>(st, m, kHC, w.h[8], kH, kH, w.E, kRefE, kRefE, P + kRef.w[9], w.c, P + kRef.b[9], nullptr,
                                 am + kRaH + 8, nullptr, am + kRaC))) return rc;  // directional_block :105-106
  ref_out_fwd_kernel<4><<<ew_blocks(m, 8), 256, 0, st>>>(w.c, P + kRef.w[10], P + kRef.b[10], m, w.o);  // :107
  LNRF_LAUNCH_CHECK("ref_out_fwd_kernel");
  ref_head_fwd2_kernel<<<ew_blocks(m, 256), 256, 0, st>>>(w.h[8], w.o, m, rgb);
  LNRF_LAUNCH_CHECK("ref_head_fwd2_kernel");
  return LNRF_OK;
}

int lnrf_refnerf_bwd(const float* params, const float* x, const float* d, const float* rays, const float* ts,
                     int64_t n, int32_t T, void* workspace, int64_t workspace_bytes, const float* d_dens,
                     const float* d_rgb, const float* d_aux_normal_mse, const float* d_aux_neg_normal,
                     float* d_params, lnrf_stream_t stream) {
  using namespace lnrf;
  LNRF_REQUIRE(n >= 0 && T >= 1, LNRF_E_INVALID, "lnrf_refnerf_bwd: n=%lld T=%d", (long long)n, T);
  const int64_t m = n * T;
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(params && workspace && d_dens && d_rgb && d_aux_normal_mse && d_aux_neg_normal && d_params,
               LNRF_E_INVALID, "lnrf_refnerf_bwd: null pointer");
  LNRF_REQUIRE((x && d && !rays && !ts) || (!x && !d && rays && ts), LNRF_E_INVALID,
               "lnrf_refnerf_bwd: pass either (x,d) or (rays,ts)");
  LNRF_REQUIRE(workspace_bytes >= carve_ref(nullptr, m, true).bytes, LNRF_E_WORKSPACE,
               "lnrf_refnerf_bwd: workspace too small");
  const RefWs w = carve_ref(workspace, m, true);
  cudaStream_t st = as_stream(stream);
  const float* P = params;
  float* G = d_params;
  const unsigned cb = ew_blocks(m, 512);  // >= 1k blocks at training sizes
  int rc;
  // ---- directional block
  ref_out_bwd_kernel<4><<<ew_blocks(m, 8 * 128), 256, 0, st>>>(w.h[8], w.o, w.c, d_rgb, P + kRef.w[10], m, w.d_o,
                                                               w.gc, G + kRef.w[10], G + kRef.b[10], kH);
  LNRF_LAUNCH_CHECK("ref_out_bwd_kernel");
  float* am = w.amax;
  LNRF_CUDA(cudaMemsetAsync(am + kRaFwdEnd, 0, (64 - kRaFwdEnd) * sizeof(float), st));
  if ((rc = tcg_amax(st, w.gc, m * kHC, am + kRaGc))) return rc;
  if ((rc = rg_tn(st, kH, kHC, w.h[8], w.gc, m, G + kRef.w[9], G + kRef.b[9], am + kRaH + 8, am + kRaGc))) return rc;
  if ((rc = rg_tn(st, kRefE, kHC, w.E, w.gc, m, G + kRef.w[9] + int64_t(kH) * kHC, nullptr, nullptr, am + kRaGc))) return rc;
  float* g = w.gA;
  float* gnext = w.gB;
  if ((rc = rg_nt<EPI_STORE>(st, m, kH, w.gc, kHC, P + kRef.w[9], g, nullptr, am + kRaGc, nullptr))) return rc;
  if ((rc = rg_nt<EPI_STORE>(st, m, kRefE, w.gc, kHC, P + kRef.w[9] + int64_t(kH) * kHC, w.dE, nullptr, am + kRaGc,
                             nullptr))) return rc;
  // ---- heads: adds dL/dz8[:, :9] into g, produces u = dL/dn_raw
  ref_head_bwd_kernel<<<ew_blocks(m, 256), 256, 0, st>>>(w.h[8], d, rays, T, w.nraw, w.o, d_dens, d_rgb,
                                                         d_aux_normal_mse, d_aux_neg_normal, w.dE, m, g, w.u);
  LNRF_LAUNCH_CHECK("ref_head_bwd_kernel");
  if ((rc = tcg_amax(st, g, m * kH, am + kRaG))) return rc;  // G_8 as the chain sees it (the heads changed nine columns)
  // ---- first-order chain through the spatial block (as NeRF's, from G_8 = g)
  for (int l = 8; l >= 1; --l) {
    const float* ga = am + kRaG + (8 - l);
    if ((rc = rg_tn(st, kH, kH, w.h[l - 1], g, m, G + kRef.w[l], G + kRef.b[l], am + kRaH + l - 1, ga))) return rc;
    if (l == 5)
      if ((rc = rg_tn(st, kXE, kH, w.xe, g, m, G + kRef.w[5] + int64_t(kH) * kH, nullptr, nullptr, ga))) return rc;
    if ((rc = rg_nt<EPI_MASK>(st, m, kH, g, kH, P + kRef.w[l], gnext, w.h[l - 1], ga, am + kRaG + (9 - l), w.mask[l - 1]))) return rc;
    float* t = g; g = gnext; gnext = t;
  }
  if ((rc = rg_tn(st, kXE, kH, w.xe, g, m, G + kRef.w[0], G + kRef.b[0], nullptr, am + kRaG + 8))) return rc;
  // ---- second-order term through real_normal: tangent pass along u
  ref_temb_kernel<<<ew_blocks(m * 3 * kXFreqs, 256), 256, 0, st>>>(x, rays, ts, T, m, w.u, w.temb);
  LNRF_LAUNCH_CHECK("ref_temb_kernel");
  if ((rc = tcg_amax(st, w.temb, m * kXE, am + kRaTemb))) return rc;
  if ((rc = rg_tn(st, kXE, kH, w.temb, w.gn[0], m, G + kRef.w[0], nullptr, am + kRaTemb, am + kRaGn + 0))) return rc;  // dW_0 += T_emb^T Gn_0
  float* tc = w.tA;
  float* tn = w.tB;
  if ((rc = rg_nn<EPI_MASK>(st, m, kH, w.temb, kXE, kXE, nullptr, 0, 0, P + kRef.w[0], tc, nullptr, w.h[0], am + kRaTemb,
                            nullptr, am + kRaT + 0, w.mask[0]))) return rc;  // T_0
  for (int l = 1; l <= 7; ++l) {
    const float* ta = am + kRaT + l - 1;
    if ((rc = rg_tn(st, kH, kH, tc, w.gn[l], m, G + kRef.w[l], nullptr, ta, am + kRaGn + l))) return rc;  // dW_l += T_{l-1}^T Gn_l
    if (l == 5) {
      if ((rc = rg_tn(st, kXE, kH, w.temb, w.gn[5], m, G + kRef.w[5] + int64_t(kH) * kH, nullptr, am + kRaTemb,
                      am + kRaGn + 5))) return rc;
      if ((rc = rg_nn<EPI_MASK>(st, m, kH, tc, kH, kH, w.temb, kXE, kXE, P + kRef.w[5], tn, nullptr, w.h[5], ta,
                                am + kRaTemb, am + kRaT + 5, w.mask[5]))) return rc;
    } else {
      if ((rc = rg_nn<EPI_MASK>(st, m, kH, tc, kH, kH, nullptr, 0, 0, P + kRef.w[l], tn, nullptr, w.h[l], ta, nullptr,
                                am + kRaT + l, w.mask[l]))) return rc;
    }
    float* t = tc; tc = tn; tn = t;
  }
  ref_w8col_kernel<<<cb, 256, 0, st>>>(tc, m, G + kRef.w[8]);  // dW_8[:, 0] -= sum T_7
  LNRF_LAUNCH_CHECK("ref_w8col_kernel");
  return LNRF_OK;
}

// ---------------------------------------------------------------- InstantNGPRefNERFModel entries
int64_t lnrf_ngpref_mlp_param_floats(int32_t L) { return lnrf::ngpref_layout(L).total; }
int lnrf_ngpref_param_offsets(int32_t L, int64_t* out_host) {
  LNRF_REQUIRE(out_host && L >= 1 && L <= 16, LNRF_E_INVALID, "lnrf_ngpref_param_offsets: bad args");
  const lnrf::NgpRefLayout n = lnrf::ngpref_layout(L);
  for (int i = 0; i < 5; ++i) {
    out_host[2 * i] = n.w[i];
    out_host[2 * i + 1] = n.b[i];
  }
  return LNRF_OK;
}
int lnrf_ngpref_workspace_bytes(int64_t m, int32_t L, int32_t save_for_backward, int64_t* bytes_out_host) {
  LNRF_REQUIRE(m >= 0 && L >= 1 && L <= 16 && bytes_out_host, LNRF_E_INVALID, "lnrf_ngpref_workspace_bytes: bad args");
  *bytes_out_host = lnrf::carve_ngpref(nullptr, m, 2 * L, save_for_backward != 0).bytes;
  return LNRF_OK;
}

int lnrf_ngpref_fwd(const float* params, const int64_t* level_offsets_host, const int32_t* grid_sizes_host,
                    const int32_t* table_sizes_host, int32_t L, const float* bbox_min_host,
                    const float* bbox_max_host, const float* x, const float* d, const float* rays, const float* ts,
                    int64_t n, int32_t T, int32_t save_for_backward, void* workspace, int64_t workspace_bytes,
                    float* dens, float* rgb, float* aux_normal_mse, float* aux_neg_normal, lnrf_stream_t stream) {
  using namespace lnrf;
  LNRF_REQUIRE(n >= 0 && T >= 1 && L >= 1 && L <= 16, LNRF_E_INVALID, "lnrf_ngpref_fwd: n=%lld T=%d L=%d", (long long)n, T, L);
  const int64_t m = n * T;
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(params && workspace && dens && rgb && aux_normal_mse && aux_neg_normal, LNRF_E_INVALID,
               "lnrf_ngpref_fwd: null pointer");
  LNRF_REQUIRE((x && d && !rays && !ts) || (!x && !d && rays && ts), LNRF_E_INVALID,
               "lnrf_ngpref_fwd: pass either (x,d) or (rays,ts)");
  LNRF_REQUIRE(m < (int64_t(1) << 31), LNRF_E_UNSUPPORTED, "lnrf_ngpref_fwd: %lld samples per call; chunk the batch",
               (long long)m);
  const int E = 2 * L;
  const bool save = save_for_backward != 0;
  LNRF_REQUIRE(workspace_bytes >= carve_ngpref(nullptr, m, E, save).bytes, LNRF_E_WORKSPACE,
               "lnrf_ngpref_fwd: workspace %lld < %lld bytes", (long long)workspace_bytes,
               (long long)carve_ngpref(nullptr, m, E, save).bytes);
  const NgpRefWs w = carve_ngpref(workspace, m, E, save);
  const NgpRefLayout nl = ngpref_layout(L);
  cudaStream_t st = as_stream(stream);
  const float* P = params;
  int rc;
#define LNRF_GRID(which, in0, in1, out0, out1)                                                                  \
  hashgrid_launch(which, P, level_offsets_host, grid_sizes_host, table_sizes_host, L, bbox_min_host, bbox_max_host, \
                  1, x, rays, ts, T, m, in0, in1, out0, out1, st)
  // ---- spatial_block (instant_ngp.py:69-82): smooth hash grid -> Dense_0 ReLU -> Dense_1
  if ((rc = LNRF_GRID(0, nullptr, nullptr, w.enc, nullptr))) return rc;
  float* am = w.amax;
  LNRF_CUDA(cudaMemsetAsync(am, 0, kNaFwdEnd * sizeof(float), st));
  if ((rc = rg_nn<EPI_BIAS_RELU>(st, m, kNrHidden, w.enc, E, E, nullptr, 0, 0, P + nl.w[0], w.h0, P + nl.b[0], nullptr,
                                 nullptr, nullptr, nullptr, nullptr, w.mask_h0))) return rc;
  if ((rc = rg_nn<EPI_BIAS>(st, m, kNrOut, w.h0, kNrHidden, kNrHidden, nullptr, 0, 0, P + nl.w[1], w.z, P + nl.b[1],
                            nullptr, nullptr, nullptr, nullptr))) return rc;
  // ---- real_normal (ref_nerf.py:38-43): Gn_0 = -W_1[:,0] [h_0 > 0]; vec = Gn_0 W_0^T; n_raw = J^T vec
  ref_seed_kernel<<<ew_blocks(m * kNrHidden, 256), 256, 0, st>>>(w.h0, P + nl.w[1], m, w.gn0, kNrHidden, kNrOut);
  LNRF_LAUNCH_CHECK("ref_seed_kernel");
  if ((rc = tcg_amax(st, P + nl.w[1], int64_t(kNrHidden) * kNrOut, am + kNaGn0))) return rc;  // |Gn_0| <= max |W_1|
  if ((rc = rg_nt<EPI_STORE>(st, m, E, w.gn0, kNrHidden, P + nl.w[0], w.vec, nullptr, am + kNaGn0, nullptr))) return rc;
  if ((rc = LNRF_GRID(2, w.vec, nullptr, w.nraw, nullptr))) return rc;
  // ---- heads (ref_nerf.py:45-75)
  ref_head_fwd1_kernel<<<ew_blocks(m, 256), 256, 0, st>>>(w.z, d, rays, T, w.nraw, m, dens, w.E, aux_normal_mse,
                                                          aux_neg_normal, kNrOut);
  LNRF_LAUNCH_CHECK("ref_head_fwd1_kernel");
  // ---- directional_block (instant_ngp.py:84-89): [spatial_out | IDE | n.(-d)] -> 64 -> 64 -> 3
  if ((rc = rg_nn<EPI_BIAS_RELU>(st, m, kNrHidden, w.z, kNrOut, kNrOut, w.E, kRefE, kRefE, P + nl.w[2], w.c1, P + nl.b[2],
                                 nullptr, nullptr, nullptr, nullptr, nullptr, w.mask_c1))) return rc;
  if ((rc = rg_nn<EPI_BIAS_RELU>(st, m, kNrHidden, w.c1, kNrHidden, kNrHidden, nullptr, 0, 0, P + nl.w[3], w.c2,
                                 P + nl.b[3], nullptr, nullptr, nullptr, nullptr))) return rc;
  ref_out_fwd_kernel<2><<<ew_blocks(m, 8), 256, 0, st>>>(w.c2, P + nl.w[4], P + nl.b[4], m, w.o);
  LNRF_LAUNCH_CHECK("ref_out_fwd_kernel");
  ref_head_fwd2_kernel<<<ew_blocks(m, 256), 256, 0, st>>>(w.z, w.o, m, rgb, kNrOut);
  LNRF_LAUNCH_CHECK("ref_head_fwd2_kernel");
  return LNRF_OK;
}

int lnrf_ngpref_bwd(const float* params, const int64_t* level_offsets_host, const int32_t* grid_sizes_host,
                    const int32_t* table_sizes_host, int32_t L, const float* bbox_min_host,
                    const float* bbox_max_host, const float* x, const float* d, const float* rays, const float* ts,
                    int64_t n, int32_t T, void* workspace, int64_t workspace_bytes, const float* d_dens,
                    const float* d_rgb, const float* d_aux_normal_mse, const float* d_aux_neg_normal,
                    float* d_params, lnrf_stream_t stream) {
  using namespace lnrf;
  LNRF_REQUIRE(n >= 0 && T >= 1 && L >= 1 && L <= 16, LNRF_E_INVALID, "lnrf_ngpref_bwd: n=%lld T=%d L=%d", (long long)n, T, L);
  const int64_t m = n * T;
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(params && workspace && d_dens && d_rgb && d_aux_normal_mse && d_aux_neg_normal && d_params,
               LNRF_E_INVALID, "lnrf_ngpref_bwd: null pointer");
  LNRF_REQUIRE((x && d && !rays && !ts) || (!x && !d && rays && ts), LNRF_E_INVALID,
               "lnrf_ngpref_bwd: pass either (x,d) or (rays,ts)");
  LNRF_REQUIRE((uintptr_t)d_params % 16 == 0, LNRF_E_INVALID, "lnrf_ngpref_bwd: d_params not 16-byte aligned");
  const int E = 2 * L;
  LNRF_REQUIRE(workspace_bytes >= carve_ngpref(nullptr, m, E, true).bytes, LNRF_E_WORKSPACE,
               "lnrf_ngpref_bwd: workspace too small");
  const NgpRefWs w = carve_ngpref(workspace, m, E, true);
  const NgpRefLayout nl = ngpref_layout(L);
  cudaStream_t st = as_stream(stream);
  const float* P = params;
  float* G = d_params;
  int rc;
  // ---- directional block: Dense_4, Dense_3, Dense_2
  ref_out_bwd_kernel<2><<<ew_blocks(m, 8 * 128), 256, 0, st>>>(w.z, w.o, w.c2, d_rgb, P + nl.w[4], m, w.d_o, w.gc2,
                                                               G + nl.w[4], G + nl.b[4], kNrOut);
  LNRF_LAUNCH_CHECK("ref_out_bwd_kernel");
  float* am = w.amax;
  LNRF_CUDA(cudaMemsetAsync(am + kNaFwdEnd, 0, (32 - kNaFwdEnd) * sizeof(float), st));
  if ((rc = tcg_amax(st, w.gc2, m * kNrHidden, am + kNaGc2))) return rc;
  if ((rc = rg_tn(st, kNrHidden, kNrHidden, w.c1, w.gc2, m, G + nl.w[3], G + nl.b[3], nullptr, am + kNaGc2))) return rc;
  if ((rc = rg_nt<EPI_MASK>(st, m, kNrHidden, w.gc2, kNrHidden, P + nl.w[3], w.gc1, w.c1, am + kNaGc2, am + kNaGc1,
                            w.mask_c1))) return rc;
  if ((rc = rg_tn(st, kNrOut, kNrHidden, w.z, w.gc1, m, G + nl.w[2], G + nl.b[2], nullptr, am + kNaGc1))) return rc;
  if ((rc = rg_tn(st, kRefE, kNrHidden, w.E, w.gc1, m, G + nl.w[2] + int64_t(kNrOut) * kNrHidden, nullptr, nullptr,
                  am + kNaGc1))) return rc;
  if ((rc = rg_nt<EPI_STORE>(st, m, kNrOut, w.gc1, kNrHidden, P + nl.w[2], w.g, nullptr, am + kNaGc1, nullptr))) return rc;
  if ((rc = rg_nt<EPI_STORE>(st, m, kRefE, w.gc1, kNrHidden, P + nl.w[2] + int64_t(kNrOut) * kNrHidden, w.dE, nullptr,
                             am + kNaGc1, nullptr))) return rc;
  // ---- heads: adds dL/dspatial_out[:, :9] into g, produces u = dL/dn_raw
  ref_head_bwd_kernel<<<ew_blocks(m, 256), 256, 0, st>>>(w.z, d, rays, T, w.nraw, w.o, d_dens, d_rgb, d_aux_normal_mse,
                                                         d_aux_neg_normal, w.dE, m, w.g, w.u, kNrOut);
  LNRF_LAUNCH_CHECK("ref_head_bwd_kernel");
  // ---- first-order chain through the spatial block and the hash grid
  if ((rc = tcg_amax(st, w.g, m * kNrOut, am + kNaG))) return rc;  // after the heads changed nine columns
  if ((rc = rg_tn(st, kNrHidden, kNrOut, w.h0, w.g, m, G + nl.w[1], G + nl.b[1], nullptr, am + kNaG))) return rc;
  if ((rc = rg_nt<EPI_MASK>(st, m, kNrHidden, w.g, kNrOut, P + nl.w[1], w.g0, w.h0, am + kNaG, am + kNaG0,
                            w.mask_h0))) return rc;
  if ((rc = rg_tn(st, E, kNrHidden, w.enc, w.g0, m, G + nl.w[0], G + nl.b[0], nullptr, am + kNaG0))) return rc;
  if ((rc = rg_nt<EPI_STORE>(st, m, E, w.g0, kNrHidden, P + nl.w[0], w.d_enc, nullptr, am + kNaG0, nullptr))) return rc;
#define LNRF_GRID_B(which, in0, in1, out0, out1)                                                                \
  hashgrid_launch(which, P, level_offsets_host, grid_sizes_host, table_sizes_host, L, bbox_min_host, bbox_max_host, \
                  1, x, rays, ts, T, m, in0, in1, out0, out1, st)
  // ---- table gradients: first-order scatter of d_enc and the second-order term through real_normal (tangent
  //      pass along u) in ONE pass over the tables; T_enc = J u
  if ((rc = hashgrid_launch(3, P, level_offsets_host, grid_sizes_host, table_sizes_host, L, bbox_min_host, bbox_max_host, 1,
                            x, rays, ts, T, m, w.vec, w.u, w.tenc, G, st, w.d_enc))) return rc;
  if ((rc = tcg_amax(st, w.tenc, m * E, am + kNaTenc))) return rc;
  if ((rc = rg_tn(st, E, kNrHidden, w.tenc, w.gn0, m, G + nl.w[0], nullptr, am + kNaTenc, am + kNaGn0))) return rc;
  if ((rc = rg_nn<EPI_MASK>(st, m, kNrHidden, w.tenc, E, E, nullptr, 0, 0, P + nl.w[0], w.t0, nullptr, w.h0,
                            am + kNaTenc, nullptr, nullptr, w.mask_h0))) return rc;                // T_0
  ref_w8col_kernel<<<ew_blocks(m, 512), kNrHidden, 0, st>>>(w.t0, m, G + nl.w[1], kNrOut);         // dW_1[:,0] -= sum T_0
  LNRF_LAUNCH_CHECK("ref_w8col_kernel");
  return LNRF_OK;
}

}  // extern "C"
