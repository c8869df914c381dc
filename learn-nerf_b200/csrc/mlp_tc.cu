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
//
// Warp roles (192 threads): warps 0-3 = epilogue (thread r <-> tile row r <-> TMEM
// lane r), warp 4 = weight producer (bulk copies), warp 5 = MMA issuer (one thread).
#include <cuda_bf16.h>

#include "lnrf_common.cuh"
#include "lnrf_math.cuh"
#include "nerf_layout.cuh"
#include "sm100_ptx.cuh"

namespace lnrf {

using namespace ptx;

// ---------------------------------------------------------------- packed weights
// Tensor layers of the fused kernel and their K chunks (64 rows of K each):
//   T0: Dense_0            K = x_emb block            N = 256
//   T1..T4: Dense_1..4     K = 4 act blocks           N = 256
//   T5: Dense_5            K = 4 act blocks + x_emb   N = 256
//   T6..T8: Dense_6..8     K = 4 act blocks           N = 256
//   T9: Dense_10 (+Dense_9 as column 128)  K = 4 act blocks + d_emb   N = 144
constexpr int kTcLayers = 10;
constexpr int kTcChunks = 39;
constexpr int kNColor = 144;                  // 128 colour units + density column + pad to 16
constexpr uint32_t kChunkBytes256 = 256 * 128;      // 32768
constexpr uint32_t kChunkBytes144 = kNColor * 128;  // 18432
constexpr int64_t kPackedBytes = 34 * int64_t(kChunkBytes256) + 5 * int64_t(kChunkBytes144);

struct ChunkInfo {
  int layer;    // Dense index providing the rows
  int k0;       // first kernel row of this chunk
  int kvalid;   // rows that exist (rest are zero padding)
  int n;        // B-operand rows (= output columns of the tensor layer)
  int ablock;   // which A block the MMA reads: 0..3 activations, 4 = embedding block
  int tlayer;   // tensor layer index 0..9
  uint32_t offset;  // byte offset inside the packed image
};

struct ChunkTable { ChunkInfo c[kTcChunks]; };

static ChunkTable build_chunk_table() {
  ChunkTable t{};
  int n = 0;
  uint32_t off = 0;
  auto add = [&](int layer, int k0, int kvalid, int ncols, int ablock, int tlayer) {
    t.c[n] = ChunkInfo{layer, k0, kvalid, ncols, ablock, tlayer, off};
    off += uint32_t(ncols) * 128u;
    ++n;
  };
  add(0, 0, kXE, 256, 4, 0);
  for (int l = 1; l <= 4; ++l)
    for (int b = 0; b < 4; ++b) add(l, b * 64, 64, 256, b, l);
  for (int b = 0; b < 4; ++b) add(5, b * 64, 64, 256, b, 5);
  add(5, 256, kXE, 256, 4, 5);
  for (int l = 6; l <= 8; ++l)
    for (int b = 0; b < 4; ++b) add(l, b * 64, 64, 256, b, l);
  for (int b = 0; b < 4; ++b) add(10, b * 64, 64, kNColor, b, 9);
  add(10, 256, kDE, kNColor, 4, 9);
  return t;
}

__constant__ ChunkTable c_chunks;
__constant__ NerfLayout c_nerf;  // device copy of kNerf (runtime-indexed in kernels)

// One thread per (chunk, n, 16-byte group of 8 k): writes the bf16 B-operand image
// B[n][k] = W[k0 + k][n] in the SW128 K-major layout the MMA expects.
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ P, uint8_t* __restrict__ packed) {
  const int ci = blockIdx.y;
  const ChunkInfo c = c_chunks.c[ci];
  const int items = c.n * 8;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
    const int n = it >> 3, kg = it & 7;
    const int out_dim = c_nerf.out[c.layer];
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kg * 8 + j;
      float w = 0.0f;
      if (k < c.kvalid) {
        if (n < out_dim) w = __ldg(P + c_nerf.w[c.layer] + int64_t(c.k0 + k) * out_dim + n);
        else if (c.layer == 10 && n == kHC && c.ablock < 4)  // density head column (Dense_9)
          w = __ldg(P + c_nerf.w[9] + (c.k0 + k));
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

// ---------------------------------------------------------------- debug GEMM
// D[128,N] = A[128,K] * B[N,K]^T, one CTA, same operand layouts / descriptors /
// TMEM read-back as the fused kernel.  K multiple of 64 (<= 256), N multiple of 16.
__global__ void __launch_bounds__(128)
debug_umma_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, int N, int K,
                       float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kb = K / 64;
  uint8_t* sA = smem;                       // kb blocks of 128 x 128 B
  uint8_t* sB = smem + kb * 16384;          // kb blocks of N x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K; i += 128) {
    int r = i / K, k = i % K;
    __nv_bfloat16 h = __float2bfloat16(A[i]);
    *reinterpret_cast<__nv_bfloat16*>(sA + (k / 64) * 16384 + sw128_offset(r, k % 64)) = h;
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    __nv_bfloat16 h = __float2bfloat16(B[i]);
    *reinterpret_cast<__nv_bfloat16*>(sB + (k / 64) * (N * 128) + sw128_offset(r, k % 64)) = h;
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
};

constexpr int kTcThreads = 192;
constexpr uint32_t kABlockBytes = 128 * 128;  // one [128 x 64] bf16 block
constexpr uint32_t kABytes = 5 * kABlockBytes;

template <int STAGES>
struct TcSmem {
  static constexpr uint32_t a_off = 0;
  static constexpr uint32_t w_off = kABytes;
  static constexpr uint32_t bar_off = w_off + STAGES * kChunkBytes256;
  static constexpr uint32_t total = bar_off + 128;  // barriers + tmem ptr
  static constexpr uint32_t alloc = total;          // dynamic smem base is declared 1024-aligned
};

// writes 8 packed bf16 pairs groups (64 values, one full row of a block) is done by callers
__device__ __forceinline__ void store_row_chunk(uint32_t block_base, int row, int chunk, uint32_t a,
                                                uint32_t b, uint32_t c, uint32_t d) {
  st_shared_v4(block_base + row * 128 + (((chunk ^ (row & 7)) & 7) << 4), a, b, c, d);
}

template <int STAGES>
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
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

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
          const uint32_t bytes = uint32_t(c_chunks.c[ci].n) * 128u;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          mbar_arrive_expect_tx(bar_full + 8 * stage, bytes);
          bulk_g2s(sW + stage * kChunkBytes256, args.packed + c_chunks.c[ci].offset, bytes,
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
          while (ci < kTcChunks && c_chunks.c[ci].tlayer == tl) {
            const ChunkInfo c = c_chunks.c[ci];
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
      tc_fence_before();
      mbar_arrive(bar_a_ready);  // event: layer T0 may start
      // ---- hidden layers T0..T8
      for (int tl = 0; tl < 9; ++tl) {
        mbar_wait(bar_acc_ready, ev & 1);
        ++ev;
        tc_fence_after();
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
          }
          const uint32_t blk = sA + (c0 >> 6) * kABlockBytes;
          const int cbase = (c0 & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            store_row_chunk(blk, r, cbase + q, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
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
        tc_fence_before();
        mbar_arrive(bar_a_ready);
      }
      // ---- T9: colour layer (+ density column) and the fp32 rgb head
      mbar_wait(bar_acc_ready, ev & 1);
      ++ev;
      tc_fence_after();
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      const float* b10 = P + c_nerf.b[10];
      const float* w11 = P + c_nerf.w[11];
#pragma unroll 1
      for (int c0 = 0; c0 < kHC; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tm_lane + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float h = fmaxf(__uint_as_float(v[j]) + __ldg(b10 + c0 + j), 0.0f);  // model.py:59
          o0 = fmaf(h, __ldg(w11 + (c0 + j) * 3 + 0), o0);
          o1 = fmaf(h, __ldg(w11 + (c0 + j) * 3 + 1), o1);
          o2 = fmaf(h, __ldg(w11 + (c0 + j) * 3 + 2), o2);
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
      tc_fence_before();  // orders these TMEM reads before the next a_ready arrive
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

static bool g_tc_ready = false;
static int g_tc_stages = 1;

int init_mlp_tc() {
  ChunkTable t = build_chunk_table();
  LNRF_CUDA(cudaMemcpyToSymbol(c_chunks, &t, sizeof(t)));
  NerfLayout lay = kNerf;
  LNRF_CUDA(cudaMemcpyToSymbol(c_nerf, &lay, sizeof(lay)));
  LNRF_CUDA(cudaFuncSetAttribute(nerf_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)TcSmem<1>::alloc));
  LNRF_CUDA(cudaFuncSetAttribute(nerf_fwd_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)TcSmem<4>::alloc));
  LNRF_CUDA(cudaFuncSetAttribute(nerf_fwd_tc_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
  LNRF_CUDA(cudaFuncSetAttribute(nerf_fwd_tc_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
  LNRF_CUDA(cudaFuncSetAttribute(debug_umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 200 * 1024));
  g_tc_ready = true;
  return LNRF_OK;
}

int64_t tc_workspace_bytes(int64_t m, bool save) {
  (void)m;
  (void)save;
  return 0;  // the fused forward keeps every activation on chip
}

int nerf_fwd_tc(const float* P, const void* packed, const float* x, const float* d, const float* rays,
                const float* ts, int64_t m, int T, bool save, void* ws, int64_t ws_bytes, float* dens,
                float* rgb, cudaStream_t st) {
  (void)ws;
  (void)ws_bytes;
  LNRF_REQUIRE(g_tc_ready, LNRF_E_INVALID, "lnrf_nerf_mlp_fwd(bf16): call lnrf_init first");
  LNRF_REQUIRE(!save, LNRF_E_UNSUPPORTED,
               "lnrf_nerf_mlp_fwd(bf16): save_for_backward is not implemented on the bf16 path yet");
  TcFwdArgs a{reinterpret_cast<const uint8_t*>(packed), P, x, d, rays, ts, T, m, dens, rgb};
  const int64_t tiles = ceil_div(m, 128);
  if (g_tc_stages == 1) {
    int64_t grid = int64_t(sm_count()) * 2;
    if (grid > tiles) grid = tiles;
    nerf_fwd_tc_kernel<1><<<(unsigned)grid, kTcThreads, TcSmem<1>::alloc, st>>>(a);
  } else {
    int64_t grid = sm_count();
    if (grid > tiles) grid = tiles;
    nerf_fwd_tc_kernel<4><<<(unsigned)grid, kTcThreads, TcSmem<4>::alloc, st>>>(a);
  }
  LNRF_LAUNCH_CHECK("nerf_fwd_tc_kernel");
  return LNRF_OK;
}

int nerf_bwd_tc(const float* P, const void* packed, int64_t m, void* ws, int64_t ws_bytes,
                const float* dens, const float* rgb, const float* d_dens, const float* d_rgb, float* G,
                cudaStream_t st) {
  (void)P; (void)packed; (void)m; (void)ws; (void)ws_bytes; (void)dens; (void)rgb; (void)d_dens;
  (void)d_rgb; (void)G; (void)st;
  LNRF_REQUIRE(false, LNRF_E_UNSUPPORTED, "lnrf_nerf_mlp_bwd(bf16): not implemented yet");
  return LNRF_OK;
}

int nerf_pack_weights(const float* P, void* packed, cudaStream_t st) {
  LNRF_REQUIRE(g_tc_ready, LNRF_E_INVALID, "lnrf_nerf_pack_weights: call lnrf_init first");
  dim3 grid(8, kTcChunks);
  pack_weights_kernel<<<grid, 256, 0, st>>>(P, reinterpret_cast<uint8_t*>(packed));
  LNRF_LAUNCH_CHECK("pack_weights_kernel");
  return LNRF_OK;
}

int64_t nerf_packed_bytes() { return kPackedBytes; }

void set_tc_stages(int stages) { g_tc_stages = (stages >= 2) ? 4 : 1; }

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

// tuning knob used by bench/tests: 1 = one ring stage, 2 CTAs/SM; >=2 = 4 stages, 1 CTA/SM
int lnrf_set_tc_stages(int32_t stages) {
  lnrf::set_tc_stages(stages);
  return LNRF_OK;
}

}  // extern "C"
