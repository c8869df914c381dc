// InstantNGPModel heads (learn_nerf/instant_ngp.py:37,46-53) on a precomputed encoding, fp32:
//   Dense_0: 2L -> 64 relu;  Dense_1: 64 -> 16 (col 0 -> exp -> density);
//   [d_emb(24) | out(16)] -> Dense_2: 40 -> 64 relu;  Dense_3: 64 -> 64 relu;  Dense_4: 64 -> 3 tanh
//
// Two implementations:
//  * train path (forward with save_for_backward + backward): the layers as split-fp16 tcgen05 GEMMs of gemm_tc.cu
//    (ngp_fwd_engine / ngp_bwd_engine below: fp32-accurate, ReLU masks as bits, operand ranges through amax slots)
//    with small kernels for the direction embedding, Dense_4 (64 -> 3) and the density exp;
//  * workspace-free forward (render), and everything under LNRF_FP32_FFMA=1 (A/B measurements): the five layers are
//    tiny (9,920 MAC per sample), so the forward and the dX chain of the backward are each ONE fused kernel: a block
//    owns a tile of 128 samples, all weights (40 KB) and the current activations sit in shared memory, and every
//    layer is an 8x8 register-tiled FFMA GEMM over that tile (64 FFMA per four 128-bit shared loads); dW = act^T g
//    on the 64x64 split-K FFMA kernel of sgemm.cuh.
#include "embed.cuh"
#include "gemm_tc.cuh"
#include "lnrf_common.cuh"
#include "lnrf_math.cuh"
#include "sgemm.cuh"

namespace lnrf {

bool fp32_ffma();  // mlp_fp32.cu

constexpr int kNgpMaxLevels = 16;
constexpr int kNgpHidden = 64, kNgpDensity = 16, kNgpDE = 24;
constexpr int kNgpIn2 = kNgpDE + kNgpDensity;  // 40
// A tile of 128 samples x 64 columns is covered by (64 / CW) x 16 threads, each owning an 8 (rows) x
// CW (cols) register tile.  Forward: CW = 8 on 128 threads (64 FFMA per four 128-bit shared loads);
// backward: CW = 4 on 256 threads (twice the warps hide the mask loads and the five dependent GEMMs:
// train step 31.3 -> 30.3 ms; the forward is 4 % slower that way and keeps CW = 8).
constexpr int kFwdCW = 8, kBwdCW = 4;
constexpr int kNgpFwdThreads = 16 * 64 / kFwdCW, kNgpBwdThreads = 16 * 64 / kBwdCW;

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
  float* amax;                   // operand ranges of the tensor-core GEMMs: [0..3] max|g0|, |go1|, |g2|, |g3| (backward),
                                 // [4..7] max|enc|, |h0|, |in2|, |h2|, [8] max|Dense_1 out| (forward with save_for_backward)
  uint32_t *mask0, *mask2;       // ReLU bit masks of Dense_0 / Dense_2 (tcg_rows layout; engine path only)
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
  w.amax = take(16);
  w.mask0 = reinterpret_cast<uint32_t*>(take(tcg_mask_words(m, kNgpHidden)));
  w.mask2 = reinterpret_cast<uint32_t*>(take(tcg_mask_words(m, kNgpHidden)));
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------- tile GEMM chain
// A block owns a tile of 128 samples.  Activations live in shared memory feature-major
// (Xt[k][row], 8-row groups XOR-swizzled by k/8 so that both the 8x8 register-tile loads and the
// transposed epilogue stores are bank-conflict free); thread (tx, ty) computes rows ty*8..+7 x
// columns tx*8..+7 of each layer: per k two 128-bit loads of activations + two of weights feed
// 64 FFMAs.
constexpr int kTM = 128;  // samples per tile
__device__ __forceinline__ int xoff(int k, int rowgrp) { return k * kTM + (((rowgrp ^ (k >> 3)) & 15) << 3); }

// kCW consecutive floats (kCW = 4 or 8) with 128-bit loads / stores
template <int kCW>
__device__ __forceinline__ void load_cols(const float* p, float (&v)[kCW]) {
#pragma unroll
  for (int q = 0; q < kCW / 4; ++q) {
    const float4 x = *reinterpret_cast<const float4*>(p + 4 * q);
    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
  }
}
template <int kCW>
__device__ __forceinline__ void store_cols(float* p, const float (&v)[kCW]) {
#pragma unroll
  for (int q = 0; q < kCW / 4; ++q)
    *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

template <int NCOLS, int kCW>  // NCOLS = 64 or 16: threads with tx * kCW >= NCOLS idle
__device__ __forceinline__ void tile_gemm(const float* __restrict__ Xt, int K, const float* __restrict__ W, int ldw,
                                          float (&acc)[8][kCW], int tx, int ty) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < kCW; ++j) acc[i][j] = 0.0f;
  if (tx * kCW >= NCOLS) return;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float* ap = Xt + xoff(k, ty);
    const float4 a0 = *reinterpret_cast<const float4*>(ap), a1 = *reinterpret_cast<const float4*>(ap + 4);
    float wv[kCW];
    load_cols<kCW>(W + k * ldw + tx * kCW, wv);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < kCW; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
  }
}
// registers (8 rows x 8 cols) -> feature-major smem at feature offset k0
template <int kCW>
__device__ __forceinline__ void tile_store_smem(float* __restrict__ Yt, int k0, const float (&v)[8][kCW], int tx, int ty) {
#pragma unroll
  for (int j = 0; j < kCW; ++j) {
    float* p = Yt + xoff(k0 + tx * kCW + j, ty);
    *reinterpret_cast<float4*>(p) = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4][j], v[5][j], v[6][j], v[7][j]);
  }
}
// registers -> row-major global [m, ld] at column offset c0
template <int kCW>
__device__ __forceinline__ void tile_store_global(float* __restrict__ dst, int ld, int c0, int64_t row0, int64_t m,
                                                  const float (&v)[8][kCW], int tx, int ty, int ncols) {
  if (tx * kCW >= ncols) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = row0 + ty * 8 + i;
    if (r >= m) continue;
    float* p = dst + r * ld + c0 + tx * kCW;
    if (tx * kCW + kCW <= ncols) {
      store_cols<kCW>(p, v[i]);
    } else {  // ragged last column group (E = 12 with 8 columns per thread: columns 8..11)
#pragma unroll
      for (int j = 0; j < kCW; ++j)
        if (tx * kCW + j < ncols) p[j] = v[i][j];
    }
  }
}
// v *= [h > 0] with h row-major in global [m, 64]
template <int kCW>
__device__ __forceinline__ void tile_mask(const float* __restrict__ h, int64_t row0, int64_t m, float (&v)[8][kCW],
                                          int tx, int ty) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = row0 + ty * 8 + i;
    float hv[kCW];
#pragma unroll
    for (int j = 0; j < kCW; ++j) hv[j] = 0.0f;
    if (r < m) load_cols<kCW>(h + r * kNgpHidden + tx * kCW, hv);
#pragma unroll
    for (int j = 0; j < kCW; ++j) v[i][j] = hv[j] > 0.0f ? v[i][j] : 0.0f;
  }
}

// max |acc| over a thread's register tile (rows past m hold zeros)
template <int kCW>
__device__ __forceinline__ float tile_absmax(const float (&v)[8][kCW], float mx) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < kCW; ++j) mx = fmaxf(mx, fabsf(v[i][j]));
  return mx;
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

constexpr int kNgpXFloats = kNgpHidden * kTM;  // one feature-major activation buffer

template <bool SAVE>
__global__ void __launch_bounds__(kNgpFwdThreads, 2)
ngp_mlp_fwd_kernel(const __grid_constant__ NgpFwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* sw = sm;                                   // all parameters, reference layout ([in][out])
  float* Xa = sm + align_up(a.nl.total, 4);
  float* Xb = Xa + kNgpXFloats;
  for (int i = threadIdx.x; i < int(a.nl.total); i += kNgpFwdThreads) sw[i] = __ldg(a.P + i);
  const float* W0 = sw + a.nl.w[0]; const float* B0 = sw + a.nl.b[0];
  const float* W1 = sw + a.nl.w[1]; const float* B1 = sw + a.nl.b[1];
  const float* W2 = sw + a.nl.w[2]; const float* B2 = sw + a.nl.b[2];
  const float* W3 = sw + a.nl.w[3]; const float* B3 = sw + a.nl.b[3];
  const float* W4 = sw + a.nl.w[4]; const float* B4 = sw + a.nl.b[4];
  constexpr int kCW = kFwdCW;
  const int t = threadIdx.x, tx = t % (64 / kCW), ty = t / (64 / kCW);
  const int64_t tiles = ceil_div(a.m, kTM);
  float acc[8][kCW];
  float mxe = 0.0f, mxh0 = 0.0f, mxo = 0.0f, mxh2 = 0.0f;  // SAVE: max|enc|, |h0|, |Dense_1 out|, |h2| seen by this thread
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTM;
    const bool owner = t < kTM;   // the first 128 threads each own one sample in the per-sample phases
    const int64_t s = row0 + t;
    const bool valid = owner && s < a.m;
    __syncthreads();  // previous tile's readers of Xa / Xb are done (also covers the weight load)
    // ---- inputs: encoding -> Xa[k < E]; d_emb = sinusoidal_emb(d, 4) (:37) -> registers
    if (owner) {
      const int rg = t >> 3, rl = t & 7;
      for (int k4 = 0; k4 < a.E / 4; ++k4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) x = __ldg(reinterpret_cast<const float4*>(a.enc + s * a.E) + k4);
        if (SAVE) mxe = fmaxf(fmaxf(mxe, fmaxf(fabsf(x.x), fabsf(x.y))), fmaxf(fabsf(x.z), fabsf(x.w)));
        Xa[xoff(k4 * 4 + 0, rg) + rl] = x.x;
        Xa[xoff(k4 * 4 + 1, rg) + rl] = x.y;
        Xa[xoff(k4 * 4 + 2, rg) + rl] = x.z;
        Xa[xoff(k4 * 4 + 3, rg) + rl] = x.w;
      }
    }
    float de[kNgpDE];
    if (owner) {
      float dv[3] = {0.f, 0.f, 0.f};
      if (valid) {
#pragma unroll
        for (int k = 0; k < 3; ++k) dv[k] = a.d ? __ldg(a.d + s * 3 + k) : __ldg(a.rays + (s / a.T) * 6 + 3 + k);
      }
#pragma unroll
      for (int dim = 0; dim < 3; ++dim)
#pragma unroll
        for (int f = 0; f < 4; ++f) sincosf(dv[dim] * float(1 << f), &de[dim * 8 + f], &de[dim * 8 + 4 + f]);
    }
    __syncthreads();
    // ---- Dense_0 (2L -> 64) + relu -> Xb                                   instant_ngp.py:46-47
    tile_gemm<64>(Xa, a.E, W0, kNgpHidden, acc, tx, ty);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < kCW; ++j) acc[i][j] = fmaxf(acc[i][j] + B0[tx * kCW + j], 0.0f);
    tile_store_smem(Xb, 0, acc, tx, ty);
    if (SAVE) {
      tile_store_global(a.ws.h0, kNgpHidden, 0, row0, a.m, acc, tx, ty, kNgpHidden);
      mxh0 = tile_absmax(acc, mxh0);
    }
    __syncthreads();
    // ---- Dense_1 (64 -> 16) -> Xa[24..39]; d_emb -> Xa[0..23]; density = exp(out[0])   :48-50
    tile_gemm<16>(Xb, kNgpHidden, W1, kNgpDensity, acc, tx, ty);
    if (tx * kCW < kNgpDensity) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < kCW; ++j) acc[i][j] += B1[tx * kCW + j];
      tile_store_smem(Xa, kNgpDE, acc, tx, ty);
      if (SAVE) {
        tile_store_global(a.ws.in2, kNgpIn2, kNgpDE, row0, a.m, acc, tx, ty, kNgpDensity);
        mxo = tile_absmax(acc, mxo);
      }
      if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t r = row0 + ty * 8 + i;
          if (r < a.m) a.dens[r] = expf(acc[i][0]);
        }
      }
    }
    if (owner) {
      const int rg = t >> 3, rl = t & 7;
#pragma unroll
      for (int k = 0; k < kNgpDE; ++k) Xa[xoff(k, rg) + rl] = de[k];
      if (SAVE && valid) {
        float4* dst = reinterpret_cast<float4*>(a.ws.in2 + s * kNgpIn2);
#pragma unroll
        for (int k4 = 0; k4 < kNgpDE / 4; ++k4) dst[k4] = make_float4(de[k4 * 4], de[k4 * 4 + 1], de[k4 * 4 + 2], de[k4 * 4 + 3]);
      }
    }
    __syncthreads();
    // ---- Dense_2 (40 -> 64) + relu -> Xb                                   :51-52
    tile_gemm<64>(Xa, kNgpIn2, W2, kNgpHidden, acc, tx, ty);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < kCW; ++j) acc[i][j] = fmaxf(acc[i][j] + B2[tx * kCW + j], 0.0f);
    tile_store_smem(Xb, 0, acc, tx, ty);  // readers of Xb (Dense_1) finished before the last barrier
    if (SAVE) {
      tile_store_global(a.ws.h2, kNgpHidden, 0, row0, a.m, acc, tx, ty, kNgpHidden);
      mxh2 = tile_absmax(acc, mxh2);
    }
    __syncthreads();
    // ---- Dense_3 (64 -> 64) + relu -> Xa
    tile_gemm<64>(Xb, kNgpHidden, W3, kNgpHidden, acc, tx, ty);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < kCW; ++j) acc[i][j] = fmaxf(acc[i][j] + B3[tx * kCW + j], 0.0f);
    tile_store_smem(Xa, 0, acc, tx, ty);  // readers of Xa (Dense_2) finished before the last barrier
    if (SAVE) tile_store_global(a.ws.h3, kNgpHidden, 0, row0, a.m, acc, tx, ty, kNgpHidden);
    __syncthreads();
    // ---- Dense_4 (64 -> 3) + tanh: thread per sample                        :53
    if (owner) {
      const int rg = t >> 3, rl = t & 7;
      float o0 = B4[0], o1 = B4[1], o2 = B4[2];
#pragma unroll 8
      for (int k = 0; k < kNgpHidden; ++k) {
        const float h = Xa[xoff(k, rg) + rl];
        o0 = fmaf(h, W4[k * 3 + 0], o0);
        o1 = fmaf(h, W4[k * 3 + 1], o1);
        o2 = fmaf(h, W4[k * 3 + 2], o2);
      }
      if (valid) {
        a.rgb[s * 3 + 0] = tanhf(o0);
        a.rgb[s * 3 + 1] = tanhf(o1);
        a.rgb[s * 3 + 2] = tanhf(o2);
      }
    }
  }
  if (SAVE) {  // operand ranges of the dW GEMMs (in2 = [d_emb | Dense_1 out]: |d_emb| <= 1)
    const float mx[4] = {mxe, mxh0, fmaxf(mxo, 1.0f), mxh2};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(mx[i] < 3.0e38f ? mx[i] : 0.0f));
      if ((t & 31) == 0 && w != 0u) atomicMax(reinterpret_cast<uint32_t*>(a.ws.amax) + 4 + i, w);
    }
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

// transposed weights in shared memory for the dX chain: Wt[j][i] = W[i][j]
struct NgpBwdSmem {
  int wt4, wt3, wt2o, wt1, wt0, x, total;  // float offsets
  int epad;
};
__host__ __device__ inline NgpBwdSmem ngp_bwd_smem(int E) {
  NgpBwdSmem s{};
  s.epad = (E + 7) / 8 * 8;
  int off = 0;
  s.wt4 = off; off += 4 * kNgpHidden;              // [4 (3 used)][64]
  s.wt3 = off; off += kNgpHidden * kNgpHidden;     // [64][64]
  s.wt2o = off; off += kNgpHidden * kNgpDensity;   // [64][16]  (rows 24..39 of Dense_2)
  s.wt1 = off; off += kNgpDensity * kNgpHidden;    // [16][64]
  s.wt0 = off; off += kNgpHidden * s.epad;         // [64][epad]
  s.x = off; off += 2 * kNgpXFloats;
  s.total = off;
  return s;
}

// dX chain: g3, g2, g_out1, g0 (written for the dW GEMMs) and d_enc.
__global__ void __launch_bounds__(kNgpBwdThreads, 2)
ngp_mlp_bwd_kernel(const __grid_constant__ NgpBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const NgpBwdSmem L = ngp_bwd_smem(a.E);
  float* Wt4 = sm + L.wt4; float* Wt3 = sm + L.wt3; float* Wt2o = sm + L.wt2o;
  float* Wt1 = sm + L.wt1; float* Wt0 = sm + L.wt0;
  float* Xa = sm + L.x;
  float* Xb = Xa + kNgpXFloats;
  const float* P = a.P;
  for (int i = threadIdx.x; i < 4 * kNgpHidden; i += kNgpBwdThreads) {  // Wt4[j][i] = W4[i][j]
    const int j = i / kNgpHidden, ii = i % kNgpHidden;
    Wt4[i] = j < 3 ? __ldg(P + a.nl.w[4] + ii * 3 + j) : 0.0f;
  }
  for (int i = threadIdx.x; i < kNgpHidden * kNgpHidden; i += kNgpBwdThreads) {  // Wt3[j][i] = W3[i][j]
    const int j = i / kNgpHidden, ii = i % kNgpHidden;
    Wt3[i] = __ldg(P + a.nl.w[3] + ii * kNgpHidden + j);
  }
  for (int i = threadIdx.x; i < kNgpHidden * kNgpDensity; i += kNgpBwdThreads) {  // Wt2o[j][i] = W2[24 + i][j]
    const int j = i / kNgpDensity, ii = i % kNgpDensity;
    Wt2o[i] = __ldg(P + a.nl.w[2] + (kNgpDE + ii) * kNgpHidden + j);
  }
  for (int i = threadIdx.x; i < kNgpDensity * kNgpHidden; i += kNgpBwdThreads) {  // Wt1[j][i] = W1[i][j]
    const int j = i / kNgpHidden, ii = i % kNgpHidden;
    Wt1[i] = __ldg(P + a.nl.w[1] + ii * kNgpDensity + j);
  }
  for (int i = threadIdx.x; i < kNgpHidden * L.epad; i += kNgpBwdThreads) {  // Wt0[j][i] = W0[i][j], zero padded
    const int j = i / L.epad, ii = i % L.epad;
    Wt0[i] = ii < a.E ? __ldg(P + a.nl.w[0] + ii * kNgpHidden + j) : 0.0f;
  }
  constexpr int kCW = kBwdCW;
  const int t = threadIdx.x, tx = t % (64 / kCW), ty = t / (64 / kCW);
  const int rg = t >> 3, rl = t & 7;
  const int64_t tiles = ceil_div(a.m, kTM);
  float acc[8][kCW];
  float mx0 = 0.0f, mx1 = 0.0f, mx2 = 0.0f, mx3 = 0.0f;  // max|g0|, |g_out1|, |g2|, |g3| of this thread's stores
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTM;
    const bool owner = t < kTM;  // the first 128 threads each own one sample in the per-sample phase
    const int64_t s = row0 + t;
    const bool valid = owner && s < a.m;
    __syncthreads();
    // ---- dp = d_rgb * tanh' -> Xa[0..3]
    if (owner) {
      float dp[4] = {0.f, 0.f, 0.f, 0.f};
      if (valid) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float y = __ldg(a.rgb + s * 3 + j);
          dp[j] = __ldg(a.d_rgb + s * 3 + j) * (1.0f - y * y);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) Xa[xoff(j, rg) + rl] = dp[j];
    }
    __syncthreads();
    // ---- g3 = (dp @ W4^T) * [h3 > 0] -> Xb, global
    tile_gemm<64>(Xa, 4, Wt4, kNgpHidden, acc, tx, ty);
    tile_mask(a.ws.h3, row0, a.m, acc, tx, ty);
    tile_store_smem(Xb, 0, acc, tx, ty);
    tile_store_global(a.ws.g3, kNgpHidden, 0, row0, a.m, acc, tx, ty, kNgpHidden);
    mx3 = tile_absmax(acc, mx3);
    __syncthreads();
    // ---- g2 = (g3 @ W3^T) * [h2 > 0] -> Xa, global
    tile_gemm<64>(Xb, kNgpHidden, Wt3, kNgpHidden, acc, tx, ty);
    tile_mask(a.ws.h2, row0, a.m, acc, tx, ty);
    tile_store_smem(Xa, 0, acc, tx, ty);
    tile_store_global(a.ws.g2, kNgpHidden, 0, row0, a.m, acc, tx, ty, kNgpHidden);
    mx2 = tile_absmax(acc, mx2);
    __syncthreads();
    // ---- g_out1 = g2 @ W2[24:40]^T, + d_dens * density on column 0 -> Xb[0..15], global
    tile_gemm<16>(Xa, kNgpHidden, Wt2o, kNgpDensity, acc, tx, ty);
    if (tx * kCW < kNgpDensity) {
      if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t r = row0 + ty * 8 + i;
          if (r < a.m) acc[i][0] += __ldg(a.d_dens + r) * __ldg(a.dens + r);
        }
      }
      tile_store_smem(Xb, 0, acc, tx, ty);
      tile_store_global(a.ws.go1, kNgpDensity, 0, row0, a.m, acc, tx, ty, kNgpDensity);
      mx1 = tile_absmax(acc, mx1);
    }
    __syncthreads();
    // ---- g0 = (g_out1 @ W1^T) * [h0 > 0] -> Xa, global
    tile_gemm<64>(Xb, kNgpDensity, Wt1, kNgpHidden, acc, tx, ty);
    tile_mask(a.ws.h0, row0, a.m, acc, tx, ty);
    tile_store_smem(Xa, 0, acc, tx, ty);
    tile_store_global(a.ws.g0, kNgpHidden, 0, row0, a.m, acc, tx, ty, kNgpHidden);
    mx0 = tile_absmax(acc, mx0);
    __syncthreads();
    // ---- d_enc = g0 @ W0^T -> global [m, E]
    if (tx * kCW < L.epad) {
      // epad <= 32: the first epad / 4 column groups are active
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < kCW; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
      for (int k = 0; k < kNgpHidden; ++k) {
        const float* ap = Xa + xoff(k, ty);
        const float4 a0 = *reinterpret_cast<const float4*>(ap), a1 = *reinterpret_cast<const float4*>(ap + 4);
        float wv[kCW];
        load_cols<kCW>(Wt0 + k * L.epad + tx * kCW, wv);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < kCW; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
      }
      tile_store_global(a.d_enc, a.E, 0, row0, a.m, acc, tx, ty, a.E);
    }
  }
  // operand ranges of the dW GEMMs (atomic max on the bit pattern: the values are non-negative)
  const float mx[4] = {mx0, mx1, mx2, mx3};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(mx[i] < 3.0e38f ? mx[i] : 0.0f));
    if ((t & 31) == 0 && w != 0u) atomicMax(reinterpret_cast<uint32_t*>(a.ws.amax) + i, w);
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
  // 32 samples at a time: every lane forms dp of its own sample, the warp then streams the 32 rows of h3
  for (int64_t base = warp * 32; base < m; base += nwarps * 32) {
    const int64_t s = base + lane;
    float dp[3] = {0.f, 0.f, 0.f};
    if (s < m) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float y = __ldg(rgb + s * 3 + j);
        dp[j] = __ldg(d_rgb + s * 3 + j) * (1.0f - y * y);
        gb[j] += dp[j];
      }
    }
    const int cnt = m - base < 32 ? int(m - base) : 32;
#pragma unroll 4
    for (int t = 0; t < cnt; ++t) {
      const float2 a = __ldg(reinterpret_cast<const float2*>(h3 + (base + t) * kNgpHidden) + lane);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float d = __shfl_sync(0xffffffffu, dp[j], t);
        gw[0][j] = fmaf(a.x, d, gw[0][j]);
        gw[1][j] = fmaf(a.y, d, gw[1][j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) gb[j] = warp_sum(gb[j]);
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


// ---------------------------------------------------------------- the train path on the tensor-core engine
// With save_for_backward the five layers run as split-fp16 tcgen05 GEMMs (gemm_tc.cu: fp32-accurate, ReLU masks
// as bits, operand ranges through the amax slots) with three small kernels around them; the fused FFMA kernels
// above stay for the forward without a workspace (render) and for LNRF_FP32_FFMA=1.

// after Dense_1: density = exp(out[0]), d_emb = sinusoidal_emb(d, 4) into in2[:, :24] (instant_ngp.py:37,49-51)
__global__ void __launch_bounds__(256)
ngp_mid_kernel(const float* __restrict__ d, const float* __restrict__ rays, int T, int64_t m, float* __restrict__ in2,
               float* __restrict__ dens, float* __restrict__ amax) {
  if (blockIdx.x == 0 && threadIdx.x == 0) amax[6] = fmaxf(amax[8], 1.0f);  // max|in2|: |d_emb| <= 1
  if (d == nullptr) {
    // directions come per RAY: the few rays a block of 256 consecutive samples touches get their 12 sincos
    // once (one per thread), every sample then copies its ray's 24 values
    constexpr int kMaxRays = 20;  // 256 / T + 2 rays per block; larger blocks of rays fall through to the loop below
    __shared__ __align__(16) float s_de[kMaxRays][kNgpDE];
    if (256 / T + 2 <= kMaxRays) {
      for (int64_t s0 = int64_t(blockIdx.x) * 256; s0 < m; s0 += int64_t(gridDim.x) * 256) {
        const int64_t ray0 = s0 / T;
        const int64_t s_last = s0 + 255 < m - 1 ? s0 + 255 : m - 1;
        const int nr = int(s_last / T - ray0) + 1;
        __syncthreads();
        if (int(threadIdx.x) < nr * 12) {
          const int r = threadIdx.x / 12, q = threadIdx.x % 12, dim = q >> 2, f = q & 3;
          const float dv = __ldg(rays + (ray0 + r) * 6 + 3 + dim);
          sincosf(dv * float(1 << f), &s_de[r][dim * 8 + f], &s_de[r][dim * 8 + 4 + f]);
        }
        __syncthreads();
        // the block's 256 rows x 6 sixteen-byte pieces, consecutive threads on consecutive pieces of a row (a warp
        // store covers 5.3 rows x 96 contiguous bytes instead of 32 rows x 16 bytes)
#pragma unroll
        for (int it = 0; it < kNgpDE / 4; ++it) {
          const int idx = it * 256 + threadIdx.x, row = idx / (kNgpDE / 4), q = idx - row * (kNgpDE / 4);
          const int64_t sr = s0 + row;
          if (sr < m)
            reinterpret_cast<float4*>(in2 + sr * kNgpIn2)[q] = reinterpret_cast<const float4*>(s_de[sr / T - ray0])[q];
        }
        const int64_t s = s0 + threadIdx.x;
        if (s < m) dens[s] = expf(in2[s * kNgpIn2 + kNgpDE]);
      }
      return;
    }
  }
  for (int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < m; s += int64_t(gridDim.x) * blockDim.x) {
    float dv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) dv[k] = d ? __ldg(d + s * 3 + k) : __ldg(rays + (s / T) * 6 + 3 + k);
    float de[kNgpDE];
#pragma unroll
    for (int dim = 0; dim < 3; ++dim)
#pragma unroll
      for (int f = 0; f < 4; ++f) sincosf(dv[dim] * float(1 << f), &de[dim * 8 + f], &de[dim * 8 + 4 + f]);
    float4* dst = reinterpret_cast<float4*>(in2 + s * kNgpIn2);
#pragma unroll
    for (int k4 = 0; k4 < kNgpDE / 4; ++k4) dst[k4] = make_float4(de[k4 * 4], de[k4 * 4 + 1], de[k4 * 4 + 2], de[k4 * 4 + 3]);
    dens[s] = expf(in2[s * kNgpIn2 + kNgpDE]);
  }
}

// Dense_4 (64 -> 3) + tanh (:53): a warp takes 32 samples, lane = sample.  The 32 x 64 tile of h3 comes in with
// coalesced 16-byte loads (two halves of 32 features) and is transposed through a padded shared-memory tile, so
// every lane then walks its own row against the broadcast weights: 5 instructions per feature and 32 samples
// (the warp-per-sample version spent 15 shuffles + 15 adds per sample on the three reductions).
__global__ void __launch_bounds__(256)
ngp_rgb_kernel(const float* __restrict__ h3, const float* __restrict__ w4, const float* __restrict__ b4, int64_t m,
               float* __restrict__ rgb) {
  __shared__ float s_tile[8][32 * 33];
  __shared__ __align__(16) float s_w[kNgpHidden * 4];
  for (int i = threadIdx.x; i < kNgpHidden * 4; i += blockDim.x) s_w[i] = (i & 3) < 3 ? __ldg(w4 + (i >> 2) * 3 + (i & 3)) : 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* tile = s_tile[wib];
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const float bb[3] = {__ldg(b4), __ldg(b4 + 1), __ldg(b4 + 2)};
  const int lr = lane >> 3, lq = lane & 7;  // loads: row 4 it + lr of the tile, floats 4 lq .. 4 lq + 3 of the half row
  for (int64_t base = warp * 32; base < m; base += nwarps * 32) {
    float o0 = bb[0], o1 = bb[1], o2 = bb[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 v[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int64_t row = base + 4 * it + lr;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < m) v[it] = __ldg(reinterpret_cast<const float4*>(h3 + row * kNgpHidden + half * 32) + lq);
      }
      __syncwarp();  // the previous half's readers are done
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        float* p = tile + (4 * it + lr) * 33 + 4 * lq;
        p[0] = v[it].x; p[1] = v[it].y; p[2] = v[it].z; p[3] = v[it].w;
      }
      __syncwarp();
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const float x = tile[lane * 33 + k];
        const float4 w = *reinterpret_cast<const float4*>(s_w + (half * 32 + k) * 4);
        o0 = fmaf(x, w.x, o0);
        o1 = fmaf(x, w.y, o1);
        o2 = fmaf(x, w.z, o2);
      }
    }
    const int64_t s = base + lane;
    if (s < m) {
      rgb[s * 3 + 0] = tanhf(o0);
      rgb[s * 3 + 1] = tanhf(o1);
      rgb[s * 3 + 2] = tanhf(o2);
    }
  }
}

// Backward of Dense_4: dp = d_rgb (1 - rgb^2); g3 = (dp @ W4^T) * [h3 > 0]; dW4 += h3^T dp; db4 += sum dp; max|g3|.
__global__ void __launch_bounds__(256)
ngp_head_bwd_kernel(const float* __restrict__ h3, const float* __restrict__ rgb, const float* __restrict__ d_rgb,
                    const float* __restrict__ w4, int64_t m, float* __restrict__ g3, float* __restrict__ dw4,
                    float* __restrict__ db4, float* __restrict__ g3_amax) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float w[2][3], gw[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, gb[3] = {0.f, 0.f, 0.f}, mx = 0.0f;
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) w[k][j] = __ldg(w4 + (lane * 2 + k) * 3 + j);
  for (int64_t base = warp * 32; base < m; base += nwarps * 32) {
    const int64_t s = base + lane;
    float dp[3] = {0.f, 0.f, 0.f};
    if (s < m) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float y = __ldg(rgb + s * 3 + j);
        dp[j] = __ldg(d_rgb + s * 3 + j) * (1.0f - y * y);
        gb[j] += dp[j];
      }
    }
    const int cnt = m - base < 32 ? int(m - base) : 32;
#pragma unroll 4
    for (int t = 0; t < cnt; ++t) {
      const float d0 = __shfl_sync(0xffffffffu, dp[0], t), d1 = __shfl_sync(0xffffffffu, dp[1], t),
                  d2 = __shfl_sync(0xffffffffu, dp[2], t);
      const float2 a = __ldg(reinterpret_cast<const float2*>(h3 + (base + t) * kNgpHidden) + lane);
      const float av[2] = {a.x, a.y};
      float g[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float tt = d0 * w[k][0] + d1 * w[k][1] + d2 * w[k][2];
        g[k] = av[k] > 0.0f ? tt : 0.0f;
        mx = fmaxf(mx, fabsf(g[k]));
        gw[k][0] = fmaf(av[k], d0, gw[k][0]);
        gw[k][1] = fmaf(av[k], d1, gw[k][1]);
        gw[k][2] = fmaf(av[k], d2, gw[k][2]);
      }
      reinterpret_cast<float2*>(g3 + (base + t) * kNgpHidden)[lane] = make_float2(g[0], g[1]);
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
#pragma unroll
  for (int j = 0; j < 3; ++j) gb[j] = warp_sum(gb[j]);
  const uint32_t mw = __reduce_max_sync(0xffffffffu, __float_as_uint(mx < 3.0e38f ? mx : 0.0f));
  if (lane == 0) {
    atomicAdd(db4 + 0, gb[0]);
    atomicAdd(db4 + 1, gb[1]);
    atomicAdd(db4 + 2, gb[2]);
    if (mw != 0u) atomicMax(reinterpret_cast<uint32_t*>(g3_amax), mw);
  }
}

// density = exp(out[0]) (:49): g_out1[:, 0] += d_dens * density
__global__ void __launch_bounds__(256)
ngp_dens_bwd_kernel(const float* __restrict__ dens, const float* __restrict__ d_dens, int64_t m, float* __restrict__ go1,
                    float* __restrict__ go1_amax) {
  float mx = 0.0f;
  for (int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < m; s += int64_t(gridDim.x) * blockDim.x) {
    const float v = go1[s * kNgpDensity] + __ldg(d_dens + s) * __ldg(dens + s);
    go1[s * kNgpDensity] = v;
    mx = fmaxf(mx, fabsf(v));
  }
  const uint32_t mw = __reduce_max_sync(0xffffffffu, __float_as_uint(mx < 3.0e38f ? mx : 0.0f));
  if ((threadIdx.x & 31) == 0 && mw != 0u) atomicMax(reinterpret_cast<uint32_t*>(go1_amax), mw);
}

static int ngp_fwd_engine(cudaStream_t st, const float* P, const NgpLayout& nl, int E, const float* enc, const float* d,
                          const float* rays, int T, int64_t m, const NgpWs& w, float* dens, float* rgb) {
  int rc;
  float* am = w.amax;
  LNRF_CUDA(cudaMemsetAsync(am + 4, 0, 8 * sizeof(float), st));
  if ((rc = tcg_amax(st, enc, m * E, am + 4))) return rc;  // fresh hash tables hold ~1e-4: the encoding needs its scale
  // Dense_0 + relu                                                          instant_ngp.py:46-47
  if ((rc = tcg_rows(st, TCG_BIAS_RELU, false, m, kNgpHidden, enc, E, E, nullptr, 0, 0, P + nl.w[0], kNgpHidden, w.h0,
                     kNgpHidden, P + nl.b[0], nullptr, 0, nullptr, nullptr, am + 4, nullptr, am + 5, nullptr, w.mask0)))
    return rc;
  // Dense_1 -> in2[:, 24:40]                                                :48
  if ((rc = tcg_rows(st, TCG_BIAS, false, m, kNgpDensity, w.h0, kNgpHidden, kNgpHidden, nullptr, 0, 0, P + nl.w[1],
                     kNgpDensity, w.in2 + kNgpDE, kNgpIn2, P + nl.b[1], nullptr, 0, nullptr, nullptr, am + 5, nullptr,
                     am + 8)))
    return rc;
  ngp_mid_kernel<<<ew_blocks(m, 256), 256, 0, st>>>(d, rays, T, m, w.in2, dens, am);
  LNRF_LAUNCH_CHECK("ngp_mid_kernel");
  // Dense_2 + relu, Dense_3 + relu                                          :51-52
  if ((rc = tcg_rows(st, TCG_BIAS_RELU, false, m, kNgpHidden, w.in2, kNgpIn2, kNgpIn2, nullptr, 0, 0, P + nl.w[2],
                     kNgpHidden, w.h2, kNgpHidden, P + nl.b[2], nullptr, 0, nullptr, nullptr, am + 6, nullptr, am + 7,
                     nullptr, w.mask2)))
    return rc;
  if ((rc = tcg_rows(st, TCG_BIAS_RELU, false, m, kNgpHidden, w.h2, kNgpHidden, kNgpHidden, nullptr, 0, 0, P + nl.w[3],
                     kNgpHidden, w.h3, kNgpHidden, P + nl.b[3], nullptr, 0, nullptr, nullptr, am + 7, nullptr, nullptr)))
    return rc;
  ngp_rgb_kernel<<<ew_blocks(m, 8 * 128), 256, 0, st>>>(w.h3, P + nl.w[4], P + nl.b[4], m, rgb);
  LNRF_LAUNCH_CHECK("ngp_rgb_kernel");
  return LNRF_OK;
}

static int ngp_bwd_engine(cudaStream_t st, const float* P, const NgpLayout& nl, int E, const float* enc, int64_t m,
                          const NgpWs& w, const float* dens, const float* rgb, const float* d_dens, const float* d_rgb,
                          float* G, float* d_enc) {
  int rc;
  float* am = w.amax;
  LNRF_CUDA(cudaMemsetAsync(am, 0, 4 * sizeof(float), st));
  ngp_head_bwd_kernel<<<ew_blocks(m, 8 * 128), 256, 0, st>>>(w.h3, rgb, d_rgb, P + nl.w[4], m, w.g3, G + nl.w[4],
                                                             G + nl.b[4], am + 3);
  LNRF_LAUNCH_CHECK("ngp_head_bwd_kernel");
  // g2 = (g3 @ W3^T) * [h2 > 0]
  if ((rc = tcg_rows(st, TCG_MASKBITS, true, m, kNgpHidden, w.g3, kNgpHidden, kNgpHidden, nullptr, 0, 0, P + nl.w[3],
                     kNgpHidden, w.g2, kNgpHidden, nullptr, nullptr, 0, nullptr, nullptr, am + 3, nullptr, am + 2,
                     w.mask2, nullptr)))
    return rc;
  // g_out1 = g2 @ W2[24:40]^T (d_emb carries no parameters), + d_dens * density on column 0
  if ((rc = tcg_rows(st, TCG_STORE, true, m, kNgpDensity, w.g2, kNgpHidden, kNgpHidden, nullptr, 0, 0,
                     P + nl.w[2] + kNgpDE * kNgpHidden, kNgpHidden, w.go1, kNgpDensity, nullptr, nullptr, 0, nullptr,
                     nullptr, am + 2, nullptr, am + 1)))
    return rc;
  ngp_dens_bwd_kernel<<<ew_blocks(m, 256), 256, 0, st>>>(dens, d_dens, m, w.go1, am + 1);
  LNRF_LAUNCH_CHECK("ngp_dens_bwd_kernel");
  // g0 = (g_out1 @ W1^T) * [h0 > 0];  d_enc = g0 @ W0^T
  if ((rc = tcg_rows(st, TCG_MASKBITS, true, m, kNgpHidden, w.go1, kNgpDensity, kNgpDensity, nullptr, 0, 0, P + nl.w[1],
                     kNgpDensity, w.g0, kNgpHidden, nullptr, nullptr, 0, nullptr, nullptr, am + 1, nullptr, am + 0,
                     w.mask0, nullptr)))
    return rc;
  if ((rc = tcg_rows(st, TCG_STORE, true, m, E, w.g0, kNgpHidden, kNgpHidden, nullptr, 0, 0, P + nl.w[0], kNgpHidden,
                     d_enc, E, nullptr, nullptr, 0, nullptr, nullptr, am + 0, nullptr, nullptr)))
    return rc;
  // dW_l = input_l^T g_l, db_l = column sums of g_l
  if ((rc = tcg_tn_acc(st, kNgpHidden, kNgpHidden, w.h2, kNgpHidden, w.g3, kNgpHidden, m, G + nl.w[3], kNgpHidden,
                       G + nl.b[3], am + 7, am + 3))) return rc;
  if ((rc = tcg_tn_acc(st, kNgpIn2, kNgpHidden, w.in2, kNgpIn2, w.g2, kNgpHidden, m, G + nl.w[2], kNgpHidden,
                       G + nl.b[2], am + 6, am + 2))) return rc;
  if ((rc = tcg_tn_acc(st, kNgpHidden, kNgpDensity, w.h0, kNgpHidden, w.go1, kNgpDensity, m, G + nl.w[1], kNgpDensity,
                       G + nl.b[1], am + 5, am + 1))) return rc;
  return tcg_tn_acc(st, E, kNgpHidden, enc, E, w.g0, kNgpHidden, m, G + nl.w[0], kNgpHidden, G + nl.b[0], am + 4, am + 0);
}

static int ngp_grid(int64_t m) {
  int64_t blocks = ceil_div(m, kTM);
  const int64_t cap = int64_t(sm_count()) * 2;  // two ~100 KB blocks per SM, persistent over tiles
  return int(blocks < cap ? blocks : cap);
}
static size_t ngp_fwd_smem(const NgpLayout& nl) { return size_t(align_up(nl.total, 4) + 2 * kNgpXFloats) * sizeof(float); }

}  // namespace lnrf

namespace lnrf {
// per-device setup, called from lnrf_init (no lazily-set process state)
int init_ngp_mlp() {
  LNRF_CUDA(cudaFuncSetAttribute(ngp_mlp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  LNRF_CUDA(cudaFuncSetAttribute(ngp_mlp_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  LNRF_CUDA(cudaFuncSetAttribute(ngp_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  return LNRF_OK;
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
  if (save && !fp32_ffma() && tcg_supported(kNgpHidden, 2 * L, 0))
    return ngp_fwd_engine(as_stream(stream), params, nl, 2 * L, enc, d, rays, T, m, w, dens, rgb);
  NgpFwdArgs a{params, nl, 2 * L, enc, d, rays, T, m, w, dens, rgb};
  const size_t smem = ngp_fwd_smem(nl);
  if (save) LNRF_CUDA(cudaMemsetAsync(w.amax + 4, 0, 4 * sizeof(float), as_stream(stream)));
  if (save) ngp_mlp_fwd_kernel<true><<<ngp_grid(m), kNgpFwdThreads, smem, as_stream(stream)>>>(a);
  else ngp_mlp_fwd_kernel<false><<<ngp_grid(m), kNgpFwdThreads, smem, as_stream(stream)>>>(a);
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
  if (!fp32_ffma() && tcg_supported(kNgpHidden, 2 * L, 0))
    return ngp_bwd_engine(st, params, nl, 2 * L, enc, m, w, dens, rgb, d_dens, d_rgb, G, d_enc);
  NgpBwdArgs a{params, nl, 2 * L, m, w, dens, rgb, d_dens, d_rgb, d_enc};
  LNRF_CUDA(cudaMemsetAsync(w.amax, 0, 4 * sizeof(float), st));
  ngp_mlp_bwd_kernel<<<ngp_grid(m), kNgpBwdThreads, size_t(ngp_bwd_smem(2 * L).total) * sizeof(float), st>>>(a);
  LNRF_LAUNCH_CHECK("ngp_mlp_bwd_kernel");
  // weight / bias gradients: dW_l = input_l^T g_l (split-K FFMA GEMM), db_l = column sums
  int rc;
  ngp_dw4_kernel<<<ew_blocks(m, 8 * 128), 256, 0, st>>>(w.h3, rgb, d_rgb, m, G + nl.w[4], G + nl.b[4]);
  LNRF_LAUNCH_CHECK("ngp_dw4_kernel");
  if (!fp32_ffma()) {
    // operand ranges: layer inputs from the forward (slots 4..7), gradients from the dX kernel above (0..3)
    if ((rc = tcg_tn_acc(st, kNgpHidden, kNgpHidden, w.h2, kNgpHidden, w.g3, kNgpHidden, m, G + nl.w[3], kNgpHidden,
                         G + nl.b[3], w.amax + 7, w.amax + 3))) return rc;
    if ((rc = tcg_tn_acc(st, kNgpIn2, kNgpHidden, w.in2, kNgpIn2, w.g2, kNgpHidden, m, G + nl.w[2], kNgpHidden,
                         G + nl.b[2], w.amax + 6, w.amax + 2))) return rc;
    if ((rc = tcg_tn_acc(st, kNgpHidden, kNgpDensity, w.h0, kNgpHidden, w.go1, kNgpDensity, m, G + nl.w[1],
                         kNgpDensity, G + nl.b[1], w.amax + 5, w.amax + 1))) return rc;
    return tcg_tn_acc(st, 2 * L, kNgpHidden, enc, 2 * L, w.g0, kNgpHidden, m, G + nl.w[0], kNgpHidden, G + nl.b[0],
                      w.amax + 4, w.amax + 0);
  }
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
