// InstantNGPModel heads (learn_nerf/instant_ngp.py:37,46-53) on the tensor cores (bf16 operands, fp32
// accumulation in TMEM; the 2e-2 precision contract of the bf16 NeRF path):
//   Dense_0: 2L -> 64 relu;  Dense_1: 64 -> 16 (col 0 -> exp -> density);
//   [d_emb(24) | out(16)] -> Dense_2: 40 -> 64 relu;  Dense_3: 64 -> 64 relu;  Dense_4: 64 -> 3 tanh
//
// The five layers are 9,920 MAC per sample: on the FP32 pipes the fused FFMA kernels (ngp_mlp.cu) are
// issue-bound and cost 2/3 of an Instant-NGP train step.  Here a CTA owns a tile of 128 samples; every
// layer input is a [128 x 64] bf16 K-major SW128 block in shared memory (K zero-padded to 64), all
// weights (28 KB forward / 40 KB transposed) stay resident, and a layer is 1..4 tcgen05.mma (M = 128,
// N = 64 or 16) followed by a TMEM read-back epilogue that writes the next block.
//   forward : 2 CTAs per SM; with save_for_backward the five blocks of a tile leave as ONE 80 KB bulk
//             store (the stash: 640 B/sample instead of 1.9 KB of fp32 activations).
//   backward: the dX chain (five small GEMMs on the transposed weights) AND every dW_l = in_l^T g_l: the
//             stash blocks and the g blocks are read as MN-major operands (K = the tile's 128 samples), the
//             five dW accumulators live in TMEM across all tiles of the CTA and are added to global memory
//             once.  A shared "ones" block appended as feature 64 of every A^T operand makes row 64 of each
//             accumulator the column sum of g, i.e. the bias gradient, for free.
#include "lnrf_math.cuh"
#include "tc_common.cuh"

namespace lnrf {

using namespace ptx;

constexpr int kNgpTcThreads = 256;
constexpr uint32_t kBlk = 16384;  // one [128 x 64] bf16 SW128 block
constexpr int kNgpStashBlocks = 5;  // ENC, H0, IN2 = [d_emb | out], H2, H3
constexpr int64_t kNgpStashTileBytes = kNgpStashBlocks * int64_t(kBlk);

// ---------------------------------------------------------------- packed weights
// forward images  B[n][k] = W_l[k][n]  ([N x 64 k] K-major SW128, zero padded):
//   W0 [64 x 64] @0, W1 [16 x 64] @8K, W2 [64 x 64] @10K, W3 [64 x 64] @18K, W4 [16 x 64] @26K    (28 KB)
// backward images B[n][k] = W_l[n][k]  (n = input feature, k = output feature), each [64 x 64]:
//   W4^T @28K, W3^T @36K, W2^T @44K, W1^T @52K, W0^T @60K                                          (40 KB)
// then the five biases (fp32): b0[64] b1[16] b2[64] b3[64] b4[4]                                      (1 KB)
__host__ __device__ constexpr uint32_t ngp_fwd_w(int l) {  // byte offset of layer l's forward image
  return l == 0 ? 0u : l == 1 ? 8192u : l == 2 ? 10240u : l == 3 ? 18432u : 26624u;
}
constexpr uint32_t kNgpFwdBytes = 28672;
__host__ __device__ constexpr uint32_t ngp_bwd_w(int l) {  // byte offset of layer l's transposed image
  return kNgpFwdBytes + uint32_t(4 - l) * 8192u;
}
constexpr uint32_t kNgpBwdBytes = 40960;
constexpr uint32_t kNgpBiasOff = kNgpFwdBytes + kNgpBwdBytes;  // 69,632
__host__ __device__ constexpr int ngp_bias(int l) {  // float offset of layer l's bias inside the bias image
  return l == 0 ? 0 : l == 1 ? 64 : l == 2 ? 80 : l == 3 ? 144 : 208;
}
constexpr int64_t kNgpPackedBytes = kNgpBiasOff + 1024;

struct NgpTcLayout {
  int in[5], out[5];
  int64_t w[5], b[5];
};
static NgpTcLayout ngp_tc_layout(int L) {
  NgpTcLayout n{};
  const int ins[5] = {2 * L, 64, 40, 64, 64};
  const int outs[5] = {64, 16, 64, 64, 3};
  int64_t off = 0;
  for (int i = 0; i < 5; ++i) {
    n.in[i] = ins[i];
    n.out[i] = outs[i];
    n.w[i] = off;
    off = align_up(off + int64_t(ins[i]) * outs[i], 4);
    n.b[i] = off;
    off = align_up(off + outs[i], 4);
  }
  return n;
}

__global__ void __launch_bounds__(256)
ngp_pack_kernel(const float* __restrict__ P, NgpTcLayout lay, uint8_t* __restrict__ packed) {
  const int img = blockIdx.y;  // 0..4 forward, 5..9 backward (layer img - 5), 10 = biases
  if (img == 10) {
    float* out = reinterpret_cast<float*>(packed + kNgpBiasOff);
    for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < 256; i += gridDim.x * blockDim.x) {
      float v = 0.0f;
#pragma unroll
      for (int l = 0; l < 5; ++l)
        if (i >= ngp_bias(l) && i < ngp_bias(l) + lay.out[l]) v = __ldg(P + lay.b[l] + (i - ngp_bias(l)));
      out[i] = v;
    }
    return;
  }
  const bool bwd = img >= 5;
  const int l = bwd ? img - 5 : img;
  const int rows = bwd ? 64 : ((l == 1 || l == 4) ? 16 : 64);
  const uint32_t base = bwd ? ngp_bwd_w(l) : ngp_fwd_w(l);
  const float* W = P + lay.w[l];
  const int in = lay.in[l], out = lay.out[l];
  for (int it = threadIdx.x + blockIdx.x * blockDim.x; it < rows * 8; it += gridDim.x * blockDim.x) {
    const int n = it >> 3, kg = it & 7;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kg * 8 + j;
      float w = 0.0f;
      if (!bwd) {
        if (k < in && n < out) w = __ldg(W + int64_t(k) * out + n);  // B[n][k] = W[k][n]
      } else {
        if (n < in && k < out) w = __ldg(W + int64_t(n) * out + k);  // B[n][k] = W[n][k]
      }
      v[j] = w;
    }
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]);
    q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]);
    q.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(packed + base + sw128_offset(n, kg * 8)) = q;
  }
}

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float4 ngp_lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 ngp_lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ngp_sincos(float a, float* s, float* c) {
  const float k = rintf(a * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, a);
  r = fmaf(k, 1.7484555e-7f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}
// K-major A x K-major B, `ksteps` K=16 steps
__device__ __forceinline__ void ngp_mma_kk(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, int N, int ksteps) {
  const uint32_t idesc = umma_idesc_bf16(128, N);
  for (int k = 0; k < ksteps; ++k)
    umma_bf16(tmem_d, umma_desc_sw128_kmajor(a_addr + k * 32), umma_desc_sw128_kmajor(b_addr + k * 32), idesc,
              k ? 1u : 0u);
}
// D[128 x N] (+)= A^T B over the tile's 128 samples: A^T = features 0..63 of block `a_addr` + the block
// `ones_addr` as features 64..127, B = columns 0..N-1 of block `b_addr`; both operands MN-major.
__device__ __forceinline__ void ngp_mma_tn(uint32_t tmem_d, uint32_t a_addr, uint32_t ones_addr, uint32_t b_addr, int N,
                                           bool accumulate) {
  const uint32_t idesc = umma_idesc_bf16_mn(128, N);
  const uint32_t lbo = ones_addr - a_addr;
  for (int k = 0; k < 8; ++k)  // 8 x 16 samples
    umma_bf16(tmem_d, umma_desc_sw128_mnmajor(a_addr + k * 2048, lbo), umma_desc_sw128_mnmajor(b_addr + k * 2048, kBlk),
              idesc, (accumulate || k) ? 1u : 0u);
}

// bias (+ReLU) + bf16 pack of this warp's 32 accumulator columns -> four 16-byte chunks of a block row
template <bool RELU>
__device__ __forceinline__ void ngp_epi32(uint32_t taddr, uint32_t sbias, uint32_t blk, int r, int ch) {
  uint32_t v[32];
  tmem_ld32(taddr, v);
  tmem_wait_ld();
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b = ngp_lds_f4(sbias + (ch * 32 + j) * 4);
    const float f0 = __uint_as_float(v[j]) + b.x, f1 = __uint_as_float(v[j + 1]) + b.y;
    const float f2 = __uint_as_float(v[j + 2]) + b.z, f3 = __uint_as_float(v[j + 3]) + b.w;
    pk[j / 2] = RELU ? pack_bf16x2_relu(f0, f1) : pack_bf16x2(f0, f1);
    pk[j / 2 + 1] = RELU ? pack_bf16x2_relu(f2, f3) : pack_bf16x2(f2, f3);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) store_row_chunk(blk, r, ch * 4 + q, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
}

// ================================================================ forward
struct NgpTcFwdArgs {
  const uint8_t* packed;
  const float* enc;   // [m, E]
  const float* d;     // [m,3] or null
  const float* rays;  // [n,2,3] (ray mode)
  int T, E;
  int64_t m;
  float* dens;
  float* rgb;
  uint8_t* stash;     // null: nothing saved
};
struct NgpFwdSmem {
  static constexpr uint32_t act = 0;                       // ENC, H0, IN2, H2, H3
  static constexpr uint32_t w = 5 * kBlk;                  // 81,920
  static constexpr uint32_t bias = w + kNgpFwdBytes;       // 110,592
  static constexpr uint32_t bar = bias + 1024;             // 111,616: [0] weights landed, [8] MMA done, [16] tmem slot
  static constexpr uint32_t total = bar + 64;
};

template <bool SAVE>
__global__ void __launch_bounds__(kNgpTcThreads, 2)
ngp_fwd_tc_kernel(const __grid_constant__ NgpTcFwdArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sb = smem_u32(smem_raw);
  if (sb & 1023u) __trap();
  const uint32_t ENC = sb, H0 = sb + kBlk, IN2 = sb + 2 * kBlk, H2 = sb + 3 * kBlk, H3 = sb + 4 * kBlk;
  const uint32_t sW = sb + NgpFwdSmem::w, sBias = sb + NgpFwdSmem::bias;
  const uint32_t bar_w = sb + NgpFwdSmem::bar, bar_mma = bar_w + 8, tmem_slot = bar_w + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & 127, ch = warp >> 2;  // tile row (= TMEM lane) and 32-column half of this warp
  const int E = args.E;
  const int64_t tiles = (args.m + 127) / 128;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_w, kNgpFwdBytes);
    bulk_g2s(sW, args.packed, kNgpFwdBytes, bar_w);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  // zero the activation blocks once (K padding: ENC cols >= E, IN2 cols >= 40 stay zero), stage the biases
  for (uint32_t i = tid * 16; i < 5 * kBlk; i += kNgpTcThreads * 16) st_shared_v4(sb + i, 0u, 0u, 0u, 0u);
  if (tid < 256) {
    const float bv = __ldg(reinterpret_cast<const float*>(args.packed + kNgpBiasOff) + tid);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sBias + tid * 4), "f"(bv) : "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + NgpFwdSmem::bar + 16);
  const uint32_t tlane = tmem + (uint32_t((warp & 3) * 32) << 16);
  mbar_wait(bar_w, 0);
  uint32_t ph = 0;
  auto mma_done = [&]() {  // all threads: wait for the committed MMAs, then the accumulator may be read
    mbar_wait(bar_mma, ph);
    ph ^= 1;
    tc_fence_after();
  };
  auto publish = [&]() {  // smem written by this thread -> visible to the tensor core; TMEM reads done
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
  };
  const int e4 = E >> 2;  // float4 per encoding row

  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t s = tile * 128 + r;
    const bool valid = s < args.m;
    if (SAVE) {  // the previous tile's stash store must have read the blocks
      if (tid == 0) bulk_wait_read0();
      __syncthreads();
    }
    // ---- P0: encoding tile (fp32 -> bf16) and d_emb
    for (int i = tid; i < 128 * e4; i += kNgpTcThreads) {
      const int row = i / e4, c = (i - row * e4) * 4;
      const int64_t sr = tile * 128 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sr < args.m) v = __ldg(reinterpret_cast<const float4*>(args.enc + sr * E + c));
      const uint32_t a = ENC + sw128_offset(row, c);
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(pack_bf16x2(v.x, v.y)), "r"(pack_bf16x2(v.z, v.w))
                   : "memory");
    }
    if (ch == 0) {
      float dv[3] = {0.f, 0.f, 0.f};
      if (valid) {
        const float* dp = args.d ? args.d + s * 3 : args.rays + (s / args.T) * 6 + 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) dv[k] = __ldg(dp + k);
      }
      uint32_t de[12];
#pragma unroll
      for (int dim = 0; dim < 3; ++dim) {  // sinusoidal_emb(d, 4): per coordinate [sin 2^0..2^3 | cos 2^0..2^3]
        float sn[4], cs[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) ngp_sincos(dv[dim] * float(1 << f), &sn[f], &cs[f]);
        de[dim * 4 + 0] = pack_bf16x2(sn[0], sn[1]);
        de[dim * 4 + 1] = pack_bf16x2(sn[2], sn[3]);
        de[dim * 4 + 2] = pack_bf16x2(cs[0], cs[1]);
        de[dim * 4 + 3] = pack_bf16x2(cs[2], cs[3]);
      }
      store_row_chunk(IN2, r, 0, de[0], de[1], de[2], de[3]);
      store_row_chunk(IN2, r, 1, de[4], de[5], de[6], de[7]);
      store_row_chunk(IN2, r, 2, de[8], de[9], de[10], de[11]);
    }
    publish();
    // ---- Dense_0 + ReLU
    if (tid == 0) {
      tc_fence_after();
      ngp_mma_kk(tmem, ENC, sW + ngp_fwd_w(0), 64, 4);
      umma_commit(bar_mma);
    }
    mma_done();
    ngp_epi32<true>(tlane + ch * 32, sBias + ngp_bias(0) * 4, H0, r, ch);
    publish();
    // ---- Dense_1: 16 outputs, column 0 -> exp -> density; the 16 values feed Dense_2 next to d_emb
    if (tid == 0) {
      tc_fence_after();
      ngp_mma_kk(tmem, H0, sW + ngp_fwd_w(1), 16, 4);
      umma_commit(bar_mma);
    }
    mma_done();
    if (ch == 0) {
      uint32_t v[16];
      tmem_ld16(tlane, v);
      tmem_wait_ld();
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b = ngp_lds_f4(sBias + (ngp_bias(1) + j) * 4);
        o[j] = __uint_as_float(v[j]) + b.x; o[j + 1] = __uint_as_float(v[j + 1]) + b.y;
        o[j + 2] = __uint_as_float(v[j + 2]) + b.z; o[j + 3] = __uint_as_float(v[j + 3]) + b.w;
      }
      if (valid) args.dens[s] = expf(o[0]);  // instant_ngp.py:49
      store_row_chunk(IN2, r, 3, pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                      pack_bf16x2(o[6], o[7]));
      store_row_chunk(IN2, r, 4, pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]), pack_bf16x2(o[12], o[13]),
                      pack_bf16x2(o[14], o[15]));
    }
    publish();
    // ---- Dense_2 + ReLU, Dense_3 + ReLU
    if (tid == 0) {
      tc_fence_after();
      ngp_mma_kk(tmem, IN2, sW + ngp_fwd_w(2), 64, 4);
      umma_commit(bar_mma);
    }
    mma_done();
    ngp_epi32<true>(tlane + ch * 32, sBias + ngp_bias(2) * 4, H2, r, ch);
    publish();
    if (tid == 0) {
      tc_fence_after();
      ngp_mma_kk(tmem, H2, sW + ngp_fwd_w(3), 64, 4);
      umma_commit(bar_mma);
    }
    mma_done();
    ngp_epi32<true>(tlane + ch * 32, sBias + ngp_bias(3) * 4, H3, r, ch);
    publish();
    // ---- Dense_4 + tanh; the five blocks of the tile are final: stash them while the head runs
    if (tid == 0) {
      tc_fence_after();
      ngp_mma_kk(tmem, H3, sW + ngp_fwd_w(4), 16, 4);
      umma_commit(bar_mma);
      if (SAVE) {
        bulk_s2g(args.stash + tile * kNgpStashTileBytes, sb, uint32_t(kNgpStashTileBytes));
        bulk_commit();
      }
    }
    mma_done();
    if (ch == 0) {
      uint32_t v[16];
      tmem_ld16(tlane, v);
      tmem_wait_ld();
      if (valid) {
        const float4 b = ngp_lds_f4(sBias + ngp_bias(4) * 4);
        args.rgb[s * 3 + 0] = tanhf(__uint_as_float(v[0]) + b.x);  // instant_ngp.py:53
        args.rgb[s * 3 + 1] = tanhf(__uint_as_float(v[1]) + b.y);
        args.rgb[s * 3 + 2] = tanhf(__uint_as_float(v[2]) + b.z);
      }
    }
    tc_fence_before();  // these TMEM reads precede the next tile's first MMA (ordered by its publish())
  }
  if (SAVE && tid == 0) bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 64);
}

// ================================================================ backward
struct NgpTcBwdArgs {
  const uint8_t* packed;
  const uint8_t* stash;
  const float* dens;
  const float* rgb;
  const float* d_dens;
  const float* d_rgb;
  int E;
  int64_t m;
  float* d_params;  // flat head gradients (ACCUMULATED): layout = NgpTcLayout
  float* d_enc;     // [m, E] (overwritten)
  NgpTcLayout lay;
};
struct NgpBwdSmem {
  static constexpr uint32_t act = 0;                   // stash blocks ENC, H0, IN2, H2, H3
  static constexpr uint32_t ones = 5 * kBlk;           // feature 64 of every A^T operand: column 0 = 1
  static constexpr uint32_t g = 6 * kBlk;              // D4, G3, G2, DZ1, G0
  static constexpr uint32_t w = 11 * kBlk;             // 180,224: five transposed weight images
  static constexpr uint32_t bar = w + kNgpBwdBytes;    // 221,184: [0] weights, [8] stash tile, [16] chain MMA, [24] dW MMAs, [32] tmem
  static constexpr uint32_t total = bar + 128;         // [64..103] one barrier per stash block (progressive refill)
};
// TMEM columns: chain accumulator, then the five dW accumulators
constexpr uint32_t kTmChain = 0;
__host__ __device__ constexpr uint32_t tm_dw(int l) { return l == 0 ? 64u : l == 1 ? 128u : l == 2 ? 144u : l == 3 ? 208u : 272u; }

__global__ void __launch_bounds__(kNgpTcThreads, 1)
ngp_bwd_tc_kernel(const __grid_constant__ NgpTcBwdArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sb = smem_u32(smem_raw);
  if (sb & 1023u) __trap();
  const uint32_t ENC = sb, H0 = sb + kBlk, IN2 = sb + 2 * kBlk, H2 = sb + 3 * kBlk, H3 = sb + 4 * kBlk;
  const uint32_t ONES = sb + NgpBwdSmem::ones;
  const uint32_t D4 = sb + NgpBwdSmem::g, G3 = D4 + kBlk, G2 = D4 + 2 * kBlk, DZ1 = D4 + 3 * kBlk, G0 = D4 + 4 * kBlk;
  const uint32_t sW = sb + NgpBwdSmem::w;
  const uint32_t bar_w = sb + NgpBwdSmem::bar, bar_ld = bar_w + 8, bar_mma = bar_w + 16, bar_dw = bar_w + 24,
                 tmem_slot = bar_w + 32, bar_blk = bar_w + 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & 127, ch = warp >> 2;
  const int E = args.E;
  const int64_t tiles = (args.m + 127) / 128;
  const int64_t my_tiles = tiles > blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_dw, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_w, kNgpBwdBytes);
    bulk_g2s(sW, args.packed + kNgpFwdBytes, kNgpBwdBytes, bar_w);
    for (int b = 0; b < 5; ++b) mbar_init(bar_blk + 8 * b, 1);
    fence_barrier_init();
    if (my_tiles > 0)
      for (int b = 0; b < 5; ++b) {
        mbar_arrive_expect_tx(bar_blk + 8 * b, kBlk);
        bulk_g2s(sb + b * kBlk, args.stash + int64_t(blockIdx.x) * kNgpStashTileBytes + b * kBlk, kBlk, bar_blk + 8 * b);
      }
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // ones block (column 0 = 1.0, the rest 0) and zeroed g blocks (K padding of D4 / DZ1 stays zero)
  for (uint32_t i = tid * 16; i < 6 * kBlk; i += kNgpTcThreads * 16) st_shared_v4(ONES + i, 0u, 0u, 0u, 0u);
  __syncthreads();
  if (tid < 128) {
    const uint32_t a = ONES + sw128_offset(tid, 0);
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(uint16_t(0x3F80)) : "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + NgpBwdSmem::bar + 32);
  const uint32_t tlane = tmem + (uint32_t((warp & 3) * 32) << 16);
  mbar_wait(bar_w, 0);
  uint32_t ph_mma = 0, ph_ld = 0, ph_dw = 0;
  // The five stash blocks of the NEXT tile are fetched one by one as soon as the last reader of the block
  // (the dW MMA of its step, retired once the following step's chain MMA has committed) is done, so the loads
  // run under the remaining steps of the current tile instead of in front of the next one.
  auto block_landed = [&](int b) { mbar_wait(bar_blk + 8 * b, ph_ld); };
  auto refill = [&](int b, int64_t next_tile) {  // thread 0 only
    mbar_arrive_expect_tx(bar_blk + 8 * b, kBlk);
    bulk_g2s(sb + b * kBlk, args.stash + next_tile * kNgpStashTileBytes + b * kBlk, kBlk, bar_blk + 8 * b);
  };
  auto chain_done = [&]() {
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
  };
  auto publish = [&]() {
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
  };
  // masked copy of this warp's 32 accumulator columns into block `dst`: g = acc where act[row, col] > 0
  auto epi_mask32 = [&](uint32_t act_blk, uint32_t dst) {
    uint32_t v[32];
    tmem_ld32(tlane + kTmChain + ch * 32, v);
    tmem_wait_ld();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int chunk = ch * 4 + q;
      const uint4 a = ngp_lds_u4(act_blk + r * 128 + (((chunk ^ (r & 7)) & 7) << 4));  // 8 bf16 activations
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float f0 = (aw[j] & 0x0000FFFFu) ? __uint_as_float(v[q * 8 + 2 * j]) : 0.0f;
        const float f1 = (aw[j] & 0xFFFF0000u) ? __uint_as_float(v[q * 8 + 2 * j + 1]) : 0.0f;
        pk[j] = pack_bf16x2(f0, f1);
      }
      store_row_chunk(dst, r, chunk, pk[0], pk[1], pk[2], pk[3]);
    }
  };

  for (int64_t t = 0; t < my_tiles; ++t) {
    const int64_t tile = blockIdx.x + t * gridDim.x;
    const int64_t s = tile * 128 + r;
    const bool valid = s < args.m;
    // ---- head gradient d4 = d_rgb * (1 - rgb^2) -> D4 cols 0..2 (instant_ngp.py:53)
    if (ch == 0) {
      float d4[3] = {0.f, 0.f, 0.f};
      if (valid) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float y = __ldg(args.rgb + s * 3 + j);
          d4[j] = __ldg(args.d_rgb + s * 3 + j) * (1.0f - y * y);
        }
      }
      store_row_chunk(D4, r, 0, pack_bf16x2(d4[0], d4[1]), pack_bf16x2(d4[2], 0.0f), 0u, 0u);
    }
    const bool more = t + 1 < my_tiles;
    const int64_t next_tile = tile + gridDim.x;
    block_landed(4);  // H3
    publish();
    // ---- g3 = (d4 @ W4^T) * (h3 > 0);  dW4 += [H3 | 1]^T d4
    if (tid == 0) {
      tc_fence_after();
      ngp_mma_kk(tmem + kTmChain, D4, sW + (ngp_bwd_w(4) - kNgpFwdBytes), 64, 1);
      umma_commit(bar_mma);
      ngp_mma_tn(tmem + tm_dw(4), H3, ONES, D4, 16, t > 0);
    }
    chain_done();
    epi_mask32(H3, G3);
    block_landed(3);  // H2
    publish();
    // ---- g2 = (g3 @ W3^T) * (h2 > 0);  dW3 += [H2 | 1]^T g3
    if (tid == 0) {
      tc_fence_after();
      ngp_mma_kk(tmem + kTmChain, G3, sW + (ngp_bwd_w(3) - kNgpFwdBytes), 64, 4);
      umma_commit(bar_mma);
      ngp_mma_tn(tmem + tm_dw(3), H2, ONES, G3, 64, t > 0);
    }
    chain_done();  // dW4's MMAs retired with it: H3 is free
    if (tid == 0 && more) refill(4, next_tile);
    epi_mask32(H2, G2);
    block_landed(2);  // IN2
    publish();
    // ---- d_in2 = g2 @ W2^T: columns 24..39 are d_out1 (+ the density term on column 24);  dW2 += [IN2 | 1]^T g2
    if (tid == 0) {
      tc_fence_after();
      ngp_mma_kk(tmem + kTmChain, G2, sW + (ngp_bwd_w(2) - kNgpFwdBytes), 64, 4);
      umma_commit(bar_mma);
      ngp_mma_tn(tmem + tm_dw(2), IN2, ONES, G2, 64, t > 0);
    }
    chain_done();
    if (tid == 0 && more) refill(3, next_tile);  // H2
    {
      uint32_t v[16];
      tmem_ld16(tlane + kTmChain + (ch == 0 ? 16 : 32), v);  // ch 0: columns 16..31, ch 1: columns 32..47
      tmem_wait_ld();
      float z[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = __uint_as_float(v[(ch == 0 ? 8 : 0) + j]);  // columns 24..31 / 32..39
      if (ch == 0 && valid) z[0] += __ldg(args.d_dens + s) * __ldg(args.dens + s);  // d exp(out0) = density (:49)
      store_row_chunk(DZ1, r, ch, pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]), pack_bf16x2(z[4], z[5]),
                      pack_bf16x2(z[6], z[7]));
    }
    block_landed(1);  // H0
    publish();
    // ---- g0 = (dz1 @ W1^T) * (h0 > 0);  dW1 += [H0 | 1]^T dz1
    if (tid == 0) {
      tc_fence_after();
      ngp_mma_kk(tmem + kTmChain, DZ1, sW + (ngp_bwd_w(1) - kNgpFwdBytes), 64, 1);
      umma_commit(bar_mma);
      ngp_mma_tn(tmem + tm_dw(1), H0, ONES, DZ1, 16, t > 0);
    }
    chain_done();
    if (tid == 0 && more) refill(2, next_tile);  // IN2
    epi_mask32(H0, G0);
    block_landed(0);  // ENC
    publish();
    // ---- d_enc = g0 @ W0^T;  dW0 += [ENC | 1]^T g0
    if (tid == 0) {
      tc_fence_after();
      ngp_mma_kk(tmem + kTmChain, G0, sW + (ngp_bwd_w(0) - kNgpFwdBytes), 64, 4);
      umma_commit(bar_mma);
      ngp_mma_tn(tmem + tm_dw(0), ENC, ONES, G0, 64, t > 0);
      umma_commit(bar_dw);  // every dW MMA of this tile has retired: its operand blocks may be overwritten
    }
    chain_done();
    if (tid == 0 && more) refill(1, next_tile);  // H0
    if (ch * 32 < E) {
      uint32_t v[32];
      tmem_ld32(tlane + kTmChain + ch * 32, v);
      tmem_wait_ld();
      if (valid) {
        float* dst = args.d_enc + s * E + ch * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (ch * 32 + j < E)
            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                             __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
    tc_fence_before();
    mbar_wait(bar_dw, ph_dw);
    ph_dw ^= 1;
    __syncthreads();
    if (tid == 0 && more) refill(0, next_tile);  // ENC
    ph_ld ^= 1;
  }
  // ---- drain: dW_l[k][n] += D_l[row k][col n] for k < in_l; db_l[n] += D_l[row 64][col n]
  tc_fence_after();
  if (my_tiles > 0) {
#pragma unroll
    for (int l = 0; l < 5; ++l) {
      const int in = args.lay.in[l], out = args.lay.out[l];
      const int ncols = (l == 1 || l == 4) ? 16 : 64;
      const int c0 = ch * 32;  // this warp's columns [c0, c0 + 32) (only ch 0 for the 16-column layers)
      if (c0 >= ncols) continue;
      const bool wrow = r < in, brow = r == 64;
      if (ncols == 16) {
        uint32_t v[16];
        tmem_ld16(tlane + tm_dw(l), v);
        tmem_wait_ld();
        if (wrow || brow) {
          float* dst = wrow ? args.d_params + args.lay.w[l] + int64_t(r) * out : args.d_params + args.lay.b[l];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < out) atomicAdd(dst + j, __uint_as_float(v[j]));
        }
      } else {
        uint32_t v[32];
        tmem_ld32(tlane + tm_dw(l) + c0, v);
        tmem_wait_ld();
        if (wrow || brow) {
          float* dst = (wrow ? args.d_params + args.lay.w[l] + int64_t(r) * out : args.d_params + args.lay.b[l]) + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ================================================================ host side
int64_t ngp_tc_packed_bytes() { return kNgpPackedBytes; }
int64_t ngp_tc_workspace_bytes(int64_t m) { return ceil_div(m, 128) * kNgpStashTileBytes; }

int init_ngp_tc() {
  LNRF_CUDA(cudaFuncSetAttribute(ngp_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NgpFwdSmem::total));
  LNRF_CUDA(cudaFuncSetAttribute(ngp_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NgpFwdSmem::total));
  LNRF_CUDA(cudaFuncSetAttribute(ngp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NgpBwdSmem::total));
  return LNRF_OK;
}

}  // namespace lnrf

extern "C" {

int64_t lnrf_ngp_packed_bytes(void) { return lnrf::ngp_tc_packed_bytes(); }

int lnrf_ngp_pack_weights(const float* params, int32_t L, void* packed, lnrf_stream_t stream) {
  using namespace lnrf;
  LNRF_REQUIRE(params && packed && L >= 1 && L <= 16 && (2 * L) % 4 == 0, LNRF_E_INVALID, "lnrf_ngp_pack_weights: L=%d", L);
  LNRF_REQUIRE((uintptr_t)packed % 1024 == 0, LNRF_E_INVALID, "lnrf_ngp_pack_weights: packed must be 1024-byte aligned");
  ngp_pack_kernel<<<dim3(2, 11), 256, 0, as_stream(stream)>>>(params, ngp_tc_layout(L), reinterpret_cast<uint8_t*>(packed));
  LNRF_LAUNCH_CHECK("ngp_pack_kernel");
  return LNRF_OK;
}

int lnrf_ngp_mlp_tc_workspace_bytes(int64_t m, int64_t* bytes_out_host) {
  LNRF_REQUIRE(m >= 0 && bytes_out_host, LNRF_E_INVALID, "lnrf_ngp_mlp_tc_workspace_bytes: bad args");
  *bytes_out_host = lnrf::ngp_tc_workspace_bytes(m);
  return LNRF_OK;
}

int lnrf_ngp_mlp_fwd_tc(const void* packed, int32_t L, const float* enc, const float* d, const float* rays, int64_t n,
                        int32_t T, int32_t save_for_backward, void* workspace, int64_t workspace_bytes, float* dens,
                        float* rgb, lnrf_stream_t stream) {
  using namespace lnrf;
  LNRF_REQUIRE(n >= 0 && T >= 1 && L >= 1 && L <= 16 && (2 * L) % 4 == 0, LNRF_E_INVALID, "lnrf_ngp_mlp_fwd_tc: n=%lld T=%d L=%d",
               (long long)n, T, L);
  const int64_t m = n * T;
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(packed && enc && dens && rgb && ((d != nullptr) != (rays != nullptr)), LNRF_E_INVALID,
               "lnrf_ngp_mlp_fwd_tc: null pointer / pass either d or rays");
  LNRF_REQUIRE((uintptr_t)packed % 1024 == 0, LNRF_E_INVALID, "lnrf_ngp_mlp_fwd_tc: packed not 1024-byte aligned");
  const bool save = save_for_backward != 0;
  if (save)
    LNRF_REQUIRE(workspace && (uintptr_t)workspace % 1024 == 0 && workspace_bytes >= ngp_tc_workspace_bytes(m),
                 LNRF_E_WORKSPACE, "lnrf_ngp_mlp_fwd_tc: workspace %lld < %lld bytes or not 1024-byte aligned",
                 (long long)workspace_bytes, (long long)ngp_tc_workspace_bytes(m));
  NgpTcFwdArgs a{reinterpret_cast<const uint8_t*>(packed), enc, d, rays, T, 2 * L, m, dens, rgb,
                 save ? reinterpret_cast<uint8_t*>(workspace) : nullptr};
  int64_t grid = int64_t(sm_count()) * 2;
  const int64_t tiles = ceil_div(m, 128);
  if (grid > tiles) grid = tiles;
  if (save) ngp_fwd_tc_kernel<true><<<(unsigned)grid, kNgpTcThreads, NgpFwdSmem::total, as_stream(stream)>>>(a);
  else ngp_fwd_tc_kernel<false><<<(unsigned)grid, kNgpTcThreads, NgpFwdSmem::total, as_stream(stream)>>>(a);
  LNRF_LAUNCH_CHECK("ngp_fwd_tc_kernel");
  return LNRF_OK;
}

int lnrf_ngp_mlp_bwd_tc(const void* packed, int32_t L, int64_t m, const void* workspace, int64_t workspace_bytes,
                        const float* dens, const float* rgb, const float* d_dens, const float* d_rgb, float* d_params,
                        float* d_enc, lnrf_stream_t stream) {
  using namespace lnrf;
  LNRF_REQUIRE(m >= 0 && L >= 1 && L <= 16 && (2 * L) % 4 == 0, LNRF_E_INVALID, "lnrf_ngp_mlp_bwd_tc: m=%lld L=%d", (long long)m, L);
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(packed && workspace && dens && rgb && d_dens && d_rgb && d_params && d_enc, LNRF_E_INVALID,
               "lnrf_ngp_mlp_bwd_tc: null pointer");
  LNRF_REQUIRE((uintptr_t)workspace % 1024 == 0 && workspace_bytes >= ngp_tc_workspace_bytes(m), LNRF_E_WORKSPACE,
               "lnrf_ngp_mlp_bwd_tc: workspace %lld < %lld bytes or misaligned", (long long)workspace_bytes,
               (long long)ngp_tc_workspace_bytes(m));
  NgpTcBwdArgs a{reinterpret_cast<const uint8_t*>(packed), reinterpret_cast<const uint8_t*>(workspace), dens, rgb, d_dens,
                 d_rgb, 2 * L, m, d_params, d_enc, ngp_tc_layout(L)};
  int64_t grid = sm_count();
  const int64_t tiles = ceil_div(m, 128);
  if (grid > tiles) grid = tiles;
  ngp_bwd_tc_kernel<<<(unsigned)grid, kNgpTcThreads, NgpBwdSmem::total, as_stream(stream)>>>(a);
  LNRF_LAUNCH_CHECK("ngp_bwd_tc_kernel");
  return LNRF_OK;
}

}  // extern "C"
