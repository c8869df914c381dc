// K2 on the tensor cores: the whole NeRFModel forward (model.py:42-62) as ONE fused
// tcgen05/TMEM kernel.  Activations never leave the SM: each CTA owns a tile of 128
// samples, keeps the 128x256 bf16 activation tile in shared memory as the UMMA
// A operand (K-major, 128B swizzle), streams the pre-packed bf16 weight chunks
// (B operand, 64(K) x N) through a ring filled by the bulk-copy (TMA) engine, and
// accumulates in TMEM (128 lanes x 256 fp32 columns).  The epilogue warps read the
// accumulator back with tcgen05.ld, add the bias, apply ReLU, convert to bf16 and
// write the next layer's A operand.  Positional encoding is computed in-kernel
// straight into the A operand; the density head rides as column 128 of the colour
// layer GEMM, the 128->3 rgb head is done in fp32 FMAs in the last epilogue.
// With SAVE the per-layer activation tiles (the shared-memory images themselves) and
// the ReLU bit masks are streamed to the stash with bulk stores for the backward.
//
// Warp roles (192 threads): warps 0-3 = epilogue (thread r <-> tile row r <-> TMEM
// lane r), warp 4 = weight producer (bulk copies), warp 5 = MMA issuer (one thread).
#include "tc_common.cuh"

namespace lnrf {

using namespace ptx;

// One thread per (chunk, n, 16-byte group of 8 k): writes the bf16 B-operand image in the
// SW128 K-major layout the MMA expects (see tc_common.cuh for the two chunk forms).
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ P, uint8_t* __restrict__ packed) {
  const int ci = blockIdx.y;
  if (ci == kAllChunks) {  // gather the small fp32 parameters (SmallParams image)
    float* out = reinterpret_cast<float*>(packed + kSmallOffset);
    const int total = int(sizeof(SmallParams) / 4);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      float v;
      if (i < 9 * 256) v = __ldg(P + c_nerf.b[i >> 8] + (i & 255));
      else if (i < 9 * 256 + 128) v = __ldg(P + c_nerf.b[10] + (i - 9 * 256));
      else if (i < 9 * 256 + 128 + 384) v = __ldg(P + c_nerf.w[11] + (i - 9 * 256 - 128));
      else if (i == 9 * 256 + 128 + 384) v = __ldg(P + c_nerf.b[9]);
      else v = __ldg(P + c_nerf.b[11] + (i - 9 * 256 - 128 - 384 - 1));
      out[i] = v;
    }
    return;
  }
  const ChunkInfo c = ci < kTcChunks ? c_chunks.f[ci]
                      : ci < kTcChunks + kBwChunks ? c_chunks.b[ci - kTcChunks]
                      : ci < kTcChunks + kBwChunks + kF2Chunks ? c_chunks.f2[ci - kTcChunks - kBwChunks]
                                                                : c_chunks.b2[ci - kTcChunks - kBwChunks - kF2Chunks];
  const int items = c.n * 8;
  const int out_dim = c_nerf.out[c.layer];
  const float* W = P + c_nerf.w[c.layer];
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
    const int n = it >> 3, kg = it & 7;
    const int ng = c.n0 + n;  // global row index of this B-operand row
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kg * 8 + j;
      float w = 0.0f;
      if (k < c.kvalid) {
        if (c.transposed) {
          w = __ldg(W + int64_t(ng) * out_dim + (c.k0 + k));
        } else if (ng < out_dim) {
          w = __ldg(W + int64_t(c.k0 + k) * out_dim + ng);
        } else if (c.layer == 10 && ng == kHC && c.ablock < 4) {  // density column (Dense_9)
          w = __ldg(P + c_nerf.w[9] + (c.k0 + k));
        }
      }
      v[j] = w;
    }
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]);
    q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]);
    q.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(packed + c.offset + sw128_offset(n, kg * 8)) = q;
  }
}

// ---------------------------------------------------------------- debug GEMMs
// D[128,N] = A[128,K] * B[N,K]^T, one CTA, same operand layouts / descriptors /
// TMEM read-back as the fused kernel.  K multiple of 64 (<= 256), N multiple of 16.
__global__ void __launch_bounds__(128)
debug_umma_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, int N, int K,
                       float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kb = K / 64;
  uint8_t* sA = smem;               // kb blocks of 128 x 128 B
  uint8_t* sB = smem + kb * 16384;  // kb blocks of N x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K; i += 128) {
    int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + (k / 64) * 16384 + sw128_offset(r, k % 64)) = __float2bfloat16(A[i]);
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + (k / 64) * (N * 128) + sw128_offset(r, k % 64)) = __float2bfloat16(B[i]);
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 256);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    for (int b = 0; b < kb; ++b) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t ad = umma_desc_sw128_kmajor(smem_u32(sA + b * 16384) + k * 32);
        uint64_t bd = umma_desc_sw128_kmajor(smem_u32(sB + b * (N * 128)) + k * 32);
        umma_bf16(tmem, ad, bd, idesc, (b | k) ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 32 && c0 + j < N; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// D[M,N] = At[K,M]^T * Bt[K,N] with BOTH operands MN-major (the dW = act^T @ grad shape:
// K = 128 samples are the rows of the same [128 x 64] SW128 block images the forward
// writes).  M in {128, 256} (two M halves -> two TMEM column ranges), N multiple of 64.
__global__ void __launch_bounds__(128)
debug_umma_gemm_tn_kernel(const float* __restrict__ At, const float* __restrict__ Bt, int M, int N,
                          float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                     // M/64 blocks of [128 samples x 64 features]
  uint8_t* sB = smem + (M / 64) * 16384;  // N/64 blocks
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * M; i += 128) {
    int r = i / M, c = i % M;
    *reinterpret_cast<__nv_bfloat16*>(sA + (c / 64) * 16384 + sw128_offset(r, c % 64)) = __float2bfloat16(At[i]);
  }
  for (int i = tid; i < 128 * N; i += 128) {
    int r = i / N, c = i % N;
    *reinterpret_cast<__nv_bfloat16*>(sB + (c / 64) * 16384 + sw128_offset(r, c % 64)) = __float2bfloat16(Bt[i]);
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16_mn(128, N);
    for (int h = 0; h < M / 128; ++h) {
      for (int k = 0; k < 8; ++k) {  // 8 x 16 samples
        uint64_t ad = umma_desc_sw128_mnmajor(smem_u32(sA + h * 2 * 16384) + k * 2048, 16384);
        uint64_t bd = umma_desc_sw128_mnmajor(smem_u32(sB) + k * 2048, 16384);
        umma_bf16(tmem + h * N, ad, bd, idesc, k ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int h = 0; h < M / 128; ++h) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + h * N + c0, v);
      tmem_wait_ld();
      for (int j = 0; j < 32; ++j) D[(h * 128 + tid) * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------- fused forward
struct TcFwdArgs {
  const uint8_t* packed;  // packed bf16 weight image (kPackedBytes)
  const float* P;         // fp32 params (biases, Dense_11)
  const float* x;         // [m,3] or null
  const float* d;         // [m,3] or null
  const float* rays;      // [n,2,3] (ray mode)
  const float* ts;        // [m]
  int T;
  int64_t m;
  float* dens;
  float* rgb;
  TcStash stash;          // used when SAVE
};

constexpr uint32_t kABytes = 5 * kABlockBytes;

template <int STAGES>
struct TcSmem {
  static constexpr uint32_t a_off = 0;
  static constexpr uint32_t w_off = kABytes;
  static constexpr uint32_t bar_off = w_off + STAGES * kChunkBytes256;
  static constexpr uint32_t total = bar_off + 128;  // barriers + tmem ptr
  static constexpr uint32_t alloc = total;          // dynamic smem base is declared 1024-aligned
};

template <int STAGES, bool SAVE>
__global__ void __launch_bounds__(kTcThreads, STAGES <= 1 ? 2 : 1)
nerf_fwd_tc_kernel(TcFwdArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using S = TcSmem<STAGES>;
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) {  // SW128 operands need 1024-byte aligned blocks
    if (threadIdx.x == 0) printf("lnrf: dynamic smem base 0x%x not 1024-aligned\n", smem_base);
    __trap();
  }
  const uint32_t sA = smem_base + S::a_off;
  const uint32_t sW = smem_base + S::w_off;
  const uint32_t bars = smem_base + S::bar_off;
  // barrier map (8 B each): full[s] = bars + 8 s; empty[s] = bars + 8 (STAGES + s);
  const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES;
  const uint32_t bar_a_ready = bars + 16 * STAGES, bar_acc_ready = bar_a_ready + 8;
  const uint32_t tmem_slot = bar_acc_ready + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (args.m + 127) / 128;
  const int64_t my_tiles = (tiles > blockIdx.x) ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_a_ready, 128);
    mbar_init(bar_acc_ready, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 4) {
    // ===== weight producer: bulk-copy chunk after chunk into the ring =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t t = 0; t < my_tiles; ++t) {
        for (int ci = 0; ci < kTcChunks; ++ci) {
          const uint32_t bytes = uint32_t(c_chunks.f[ci].n) * 128u;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          mbar_arrive_expect_tx(bar_full + 8 * stage, bytes);
          bulk_g2s(sW + stage * kChunkBytes256, args.packed + c_chunks.f[ci].offset, bytes,
                   bar_full + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, ev = 0;
      for (int64_t t = 0; t < my_tiles; ++t) {
        int ci = 0;
        for (int tl = 0; tl < kTcLayers; ++tl) {
          mbar_wait(bar_a_ready, ev & 1);  // A operand of this layer is in smem, accumulator is free
          tc_fence_after();
          bool first = true;
          while (ci < kTcChunks && c_chunks.f[ci].tlayer == tl) {
            const ChunkInfo c = c_chunks.f[ci];
            const uint32_t idesc = umma_idesc_bf16(128, c.n);
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t a_base = sA + c.ablock * kABlockBytes;
            const uint32_t b_base = sW + stage * kChunkBytes256;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(tmem, umma_desc_sw128_kmajor(a_base + k * 32),
                        umma_desc_sw128_kmajor(b_base + k * 32), idesc, (first && k == 0) ? 0u : 1u);
            }
            first = false;
            umma_commit(bar_empty + 8 * stage);  // ring slot reusable once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            ++ci;
          }
          umma_commit(bar_acc_ready);  // accumulator of layer tl complete
          ++ev;
        }
      }
    }
  } else {
    // ===== epilogue warps: thread r owns tile row r =====
    const int r = tid;
    const uint32_t tm_lane = tmem + (uint32_t(warp * 32) << 16);
    const float* P = args.P;
    uint32_t ev = 0;
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t tile = blockIdx.x + t * gridDim.x;
      const int64_t s = tile * 128 + r;
      const bool valid = s < args.m;
      uint32_t* mask_tile = SAVE ? args.stash.MASK + (tile * 9) * 1024 + warp * 256 : nullptr;
      // ---- inputs: point and direction of this sample
      float px[3] = {0.f, 0.f, 0.f}, dv[3] = {0.f, 0.f, 0.f};
      if (valid) {
        if (args.x) {
#pragma unroll
          for (int k = 0; k < 3; ++k) { px[k] = __ldg(args.x + s * 3 + k); dv[k] = __ldg(args.d + s * 3 + k); }
        } else {
          const int64_t ray = s / args.T;
          const float tt = __ldg(args.ts + s);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            dv[k] = __ldg(args.rays + ray * 6 + 3 + k);
            px[k] = __fadd_rn(__ldg(args.rays + ray * 6 + k), __fmul_rn(dv[k], tt));  // render.py:153
          }
        }
      }
      if (SAVE) {  // block 4 may still be read by the previous tile's d_emb bulk store
        if (tid == 0) bulk_wait_read0();
        epi_bar();
      }
      // ---- sinusoidal_emb(x, 10) -> A block 4 (cols dim*20 + [sin f | cos f]), cols 60..63 = 0
      {
        uint32_t pk[32];
#pragma unroll
        for (int dim = 0; dim < 3; ++dim) {
          float sn[kXFreqs], cs[kXFreqs];
#pragma unroll
          for (int f = 0; f < kXFreqs; ++f) sincosf(px[dim] * float(1 << f), &sn[f], &cs[f]);
#pragma unroll
          for (int f = 0; f < kXFreqs; f += 2) {
            pk[dim * 10 + f / 2] = pack_bf16x2(sn[f], sn[f + 1]);
            pk[dim * 10 + 5 + f / 2] = pack_bf16x2(cs[f], cs[f + 1]);
          }
        }
        pk[30] = 0u; pk[31] = 0u;
        const uint32_t blk = sA + 4 * kABlockBytes;
#pragma unroll
        for (int c = 0; c < 8; ++c) store_row_chunk(blk, r, c, pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      }
      // ---- sinusoidal_emb(d, 4), kept in registers until the colour layer
      uint32_t de[12];
#pragma unroll
      for (int dim = 0; dim < 3; ++dim) {
        float sn[kDFreqs], cs[kDFreqs];
#pragma unroll
        for (int f = 0; f < kDFreqs; ++f) sincosf(dv[dim] * float(1 << f), &sn[f], &cs[f]);
        de[dim * 4 + 0] = pack_bf16x2(sn[0], sn[1]);
        de[dim * 4 + 1] = pack_bf16x2(sn[2], sn[3]);
        de[dim * 4 + 2] = pack_bf16x2(cs[0], cs[1]);
        de[dim * 4 + 3] = pack_bf16x2(cs[2], cs[3]);
      }
      fence_proxy_async_smem();
      if (SAVE) {
        epi_bar();
        if (tid == 0) {
          bulk_s2g(args.stash.XE + tile * kABlockBytes, sA + 4 * kABlockBytes, kABlockBytes);
          bulk_commit();
        }
      }
      tc_fence_before();
      mbar_arrive(bar_a_ready);  // event: layer T0 may start
      // ---- hidden layers T0..T8
      for (int tl = 0; tl < 9; ++tl) {
        mbar_wait(bar_acc_ready, ev & 1);
        ++ev;
        tc_fence_after();
        if (SAVE) {  // the previous layer's tile image must have left smem before we overwrite it
          if (tid == 0) bulk_wait_read0();
          epi_bar();
        }
        const float* bias = P + c_nerf.b[tl];
        const bool relu = tl < 8;  // Dense_8's output feeds the heads raw (model.py:53-58)
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tm_lane + c0, v);
          tmem_wait_ld();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
            const float f0 = __uint_as_float(v[j]) + b.x, f1 = __uint_as_float(v[j + 1]) + b.y;
            const float f2 = __uint_as_float(v[j + 2]) + b.z, f3 = __uint_as_float(v[j + 3]) + b.w;
            pk[j / 2] = relu ? pack_bf16x2_relu(f0, f1) : pack_bf16x2(f0, f1);
            pk[j / 2 + 1] = relu ? pack_bf16x2_relu(f2, f3) : pack_bf16x2(f2, f3);
            if (SAVE) {  // reuse v[] for the mask words: bit `lane` of word j = (h[row, c0+j] > 0)
              v[j] = __ballot_sync(0xffffffffu, f0 > 0.0f);
              v[j + 1] = __ballot_sync(0xffffffffu, f1 > 0.0f);
              v[j + 2] = __ballot_sync(0xffffffffu, f2 > 0.0f);
              v[j + 3] = __ballot_sync(0xffffffffu, f3 > 0.0f);
            }
          }
          const uint32_t blk = sA + (c0 >> 6) * kABlockBytes;
          const int cbase = (c0 & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            store_row_chunk(blk, r, cbase + q, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          if (SAVE && relu && lane == 0) {
            uint4* dst = reinterpret_cast<uint4*>(mask_tile + tl * 1024 + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          }
        }
        if (tl == 8) {  // x_emb is dead after T5: block 4 now carries d_emb (24 cols) + zeros
          const uint32_t blk = sA + 4 * kABlockBytes;
          store_row_chunk(blk, r, 0, de[0], de[1], de[2], de[3]);
          store_row_chunk(blk, r, 1, de[4], de[5], de[6], de[7]);
          store_row_chunk(blk, r, 2, de[8], de[9], de[10], de[11]);
#pragma unroll
          for (int c = 3; c < 8; ++c) store_row_chunk(blk, r, c, 0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
        if (SAVE) {
          epi_bar();
          if (tid == 0) {
            bulk_s2g(args.stash.H[tl] + tile * kTileBytes, sA, kTileBytes);
            if (tl == 8) bulk_s2g(args.stash.DE + tile * kABlockBytes, sA + 4 * kABlockBytes, kABlockBytes);
            bulk_commit();
          }
        }
        tc_fence_before();
        mbar_arrive(bar_a_ready);
      }
      // ---- T9: colour layer (+ density column) and the fp32 rgb head
      mbar_wait(bar_acc_ready, ev & 1);
      ++ev;
      tc_fence_after();
      if (SAVE) {  // blocks 0,1 get the colour-hidden image once the z8 image has left smem
        if (tid == 0) bulk_wait_read0();
        epi_bar();
      }
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      const float* b10 = P + c_nerf.b[10];
      const float* w11 = P + c_nerf.w[11];
#pragma unroll 1
      for (int c0 = 0; c0 < kHC; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tm_lane + c0, v);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float h0 = fmaxf(__uint_as_float(v[j]) + __ldg(b10 + c0 + j), 0.0f);  // model.py:59
          const float h1 = fmaxf(__uint_as_float(v[j + 1]) + __ldg(b10 + c0 + j + 1), 0.0f);
          o0 = fmaf(h0, __ldg(w11 + (c0 + j) * 3 + 0), o0);
          o1 = fmaf(h0, __ldg(w11 + (c0 + j) * 3 + 1), o1);
          o2 = fmaf(h0, __ldg(w11 + (c0 + j) * 3 + 2), o2);
          o0 = fmaf(h1, __ldg(w11 + (c0 + j) * 3 + 3), o0);
          o1 = fmaf(h1, __ldg(w11 + (c0 + j) * 3 + 4), o1);
          o2 = fmaf(h1, __ldg(w11 + (c0 + j) * 3 + 5), o2);
          if (SAVE) {
            pk[j / 2] = pack_bf16x2(h0, h1);
            v[j] = __ballot_sync(0xffffffffu, h0 > 0.0f);
            v[j + 1] = __ballot_sync(0xffffffffu, h1 > 0.0f);
          }
        }
        if (SAVE) {
          const uint32_t blk = sA + (c0 >> 6) * kABlockBytes;
          const int cbase = (c0 & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            store_row_chunk(blk, r, cbase + q, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          if (lane == 0) {
            uint4* dst = reinterpret_cast<uint4*>(mask_tile + 8 * 1024 + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          }
        }
      }
      {
        uint32_t v[32];
        tmem_ld32(tm_lane + kHC, v);  // column 128 = Dense_9 pre-activation
        tmem_wait_ld();
        if (valid) {
          args.dens[s] = softplus_f(__uint_as_float(v[0]) + __ldg(P + c_nerf.b[9]));  // model.py:57
          const float* b11 = P + c_nerf.b[11];
          args.rgb[s * 3 + 0] = tanhf(o0 + __ldg(b11 + 0));  // model.py:60
          args.rgb[s * 3 + 1] = tanhf(o1 + __ldg(b11 + 1));
          args.rgb[s * 3 + 2] = tanhf(o2 + __ldg(b11 + 2));
        }
      }
      if (SAVE) {
        fence_proxy_async_smem();
        epi_bar();
        if (tid == 0) {
          bulk_s2g(args.stash.C + tile * 2 * kABlockBytes, sA, 2 * kABlockBytes);
          bulk_commit();
        }
      }
      tc_fence_before();  // orders these TMEM reads before the next a_ready arrive
    }
    if (SAVE && tid == 0) bulk_wait0();  // all stash stores complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

static bool g_tc_ready = false;
static int g_tc_stages = 0;  // 0 = pair kernel (default); 1 / 4 = single-tile kernels (render only)

int init_mlp_tc_bwd();  // mlp_tc_bwd.cu
int init_mlp_tc_fwd2();  // mlp_tc_fwd2.cu
int init_mlp_tc_bwd2();  // mlp_tc_bwd2.cu
int nerf_fwd_pair(const void* packed, const float* x, const float* d, const float* rays, const float* ts,
                  int64_t m, int T, bool save, const TcStash& stash, float* dens, float* rgb, cudaStream_t st);
void set_dw_debug(int flags);
void set_fwd_debug(int flags);

template <typename K>
static int set_smem(K kernel, int bytes) {
  LNRF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  LNRF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
  return LNRF_OK;
}

int init_mlp_tc() {
  int rc = upload_tc_tables();
  if (rc) return rc;
  if ((rc = set_smem(nerf_fwd_tc_kernel<1, false>, (int)TcSmem<1>::alloc))) return rc;
  if ((rc = set_smem(nerf_fwd_tc_kernel<1, true>, (int)TcSmem<1>::alloc))) return rc;
  if ((rc = set_smem(nerf_fwd_tc_kernel<4, false>, (int)TcSmem<4>::alloc))) return rc;
  if ((rc = set_smem(nerf_fwd_tc_kernel<4, true>, (int)TcSmem<4>::alloc))) return rc;
  if ((rc = set_smem(debug_umma_gemm_kernel, 200 * 1024))) return rc;
  if ((rc = set_smem(debug_umma_gemm_tn_kernel, 200 * 1024))) return rc;
  if ((rc = init_mlp_tc_bwd())) return rc;
  if ((rc = init_mlp_tc_fwd2())) return rc;
  if ((rc = init_mlp_tc_bwd2())) return rc;
  g_tc_ready = true;
  return LNRF_OK;
}

bool tc_ready() { return g_tc_ready; }

int64_t tc_workspace_bytes(int64_t m, bool save) {
  return save ? carve_stash(nullptr, m).bytes : 0;  // render keeps every activation on chip
}

int nerf_fwd_tc(const float* P, const void* packed, const float* x, const float* d, const float* rays,
                const float* ts, int64_t m, int T, bool save, void* ws, int64_t ws_bytes, float* dens,
                float* rgb, cudaStream_t st) {
  LNRF_REQUIRE(g_tc_ready, LNRF_E_INVALID, "lnrf_nerf_mlp_fwd(bf16): call lnrf_init first");
  TcFwdArgs a{reinterpret_cast<const uint8_t*>(packed), P, x, d, rays, ts, T, m, dens, rgb, TcStash{}};
  if (save) {
    LNRF_REQUIRE(ws && (uintptr_t)ws % 1024 == 0 && ws_bytes >= tc_workspace_bytes(m, true),
                 LNRF_E_WORKSPACE, "lnrf_nerf_mlp_fwd(bf16): workspace %lld < %lld bytes or not "
                 "1024-byte aligned", (long long)ws_bytes, (long long)tc_workspace_bytes(m, true));
    a.stash = carve_stash(ws, m);
  }
  if (save || g_tc_stages == 0)  // the stash (row-major masks) is only written by the pair kernel
    return nerf_fwd_pair(packed, x, d, rays, ts, m, T, save, a.stash, dens, rgb, st);
  const int64_t tiles = ceil_div(m, 128);
  int64_t grid = int64_t(sm_count()) * (g_tc_stages == 1 ? 2 : 1);
  if (grid > tiles) grid = tiles;
  if (g_tc_stages == 1) {
    if (save) nerf_fwd_tc_kernel<1, true><<<(unsigned)grid, kTcThreads, TcSmem<1>::alloc, st>>>(a);
    else nerf_fwd_tc_kernel<1, false><<<(unsigned)grid, kTcThreads, TcSmem<1>::alloc, st>>>(a);
  } else {
    if (save) nerf_fwd_tc_kernel<4, true><<<(unsigned)grid, kTcThreads, TcSmem<4>::alloc, st>>>(a);
    else nerf_fwd_tc_kernel<4, false><<<(unsigned)grid, kTcThreads, TcSmem<4>::alloc, st>>>(a);
  }
  LNRF_LAUNCH_CHECK("nerf_fwd_tc_kernel");
  return LNRF_OK;
}

int nerf_pack_weights(const float* P, void* packed, cudaStream_t st) {
  LNRF_REQUIRE(g_tc_ready, LNRF_E_INVALID, "lnrf_nerf_pack_weights: call lnrf_init first");
  dim3 grid(8, kAllChunks + 1);
  pack_weights_kernel<<<grid, 256, 0, st>>>(P, reinterpret_cast<uint8_t*>(packed));
  LNRF_LAUNCH_CHECK("pack_weights_kernel");
  return LNRF_OK;
}

int64_t nerf_packed_bytes() { return kPackedBytes; }

void set_tc_stages(int stages) { g_tc_stages = stages <= 0 ? 0 : (stages >= 2 ? 4 : 1); }

}  // namespace lnrf

extern "C" {

int lnrf_debug_umma_gemm(const float* a, const float* b, int32_t N, int32_t K, float* d_out,
                         lnrf_stream_t stream) {
  LNRF_REQUIRE(lnrf::g_tc_ready, LNRF_E_INVALID, "lnrf_debug_umma_gemm: call lnrf_init first");
  LNRF_REQUIRE(a && b && d_out, LNRF_E_INVALID, "lnrf_debug_umma_gemm: null pointer");
  LNRF_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 64 && K <= 256 && K % 64 == 0,
               LNRF_E_UNSUPPORTED, "lnrf_debug_umma_gemm: N=%d K=%d", N, K);
  size_t smem = size_t(K / 64) * (16384 + N * 128) + 1024;
  lnrf::debug_umma_gemm_kernel<<<1, 128, smem, lnrf::as_stream(stream)>>>(a, b, N, K, d_out);
  LNRF_LAUNCH_CHECK("debug_umma_gemm_kernel");
  return LNRF_OK;
}

int lnrf_debug_umma_gemm_tn(const float* at, const float* bt, int32_t M, int32_t N, float* d_out,
                            lnrf_stream_t stream) {
  LNRF_REQUIRE(lnrf::g_tc_ready, LNRF_E_INVALID, "lnrf_debug_umma_gemm_tn: call lnrf_init first");
  LNRF_REQUIRE(at && bt && d_out, LNRF_E_INVALID, "lnrf_debug_umma_gemm_tn: null pointer");
  LNRF_REQUIRE((M == 128 || M == 256) && N >= 64 && N <= 256 && N % 64 == 0 && (M / 128) * N <= 512,
               LNRF_E_UNSUPPORTED, "lnrf_debug_umma_gemm_tn: M=%d N=%d", M, N);
  size_t smem = size_t(M / 64 + N / 64) * 16384 + 1024;
  lnrf::debug_umma_gemm_tn_kernel<<<1, 128, smem, lnrf::as_stream(stream)>>>(at, bt, M, N, d_out);
  LNRF_LAUNCH_CHECK("debug_umma_gemm_tn_kernel");
  return LNRF_OK;
}

// tuning knob used by bench/tests: 1 = one ring stage, 2 CTAs/SM; >=2 = 4 stages, 1 CTA/SM
int lnrf_set_tc_stages(int32_t stages) {
  lnrf::set_tc_stages(stages);
  return LNRF_OK;
}

// ablation switches for profiling the dW kernel (results are wrong when non-zero)
int lnrf_set_debug_flags(int32_t flags) {
  lnrf::set_dw_debug(flags);
  lnrf::set_fwd_debug(flags >= 1000 ? 0 : flags);
  return LNRF_OK;
}

}  // extern "C"
