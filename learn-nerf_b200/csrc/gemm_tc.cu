// fp32-accurate GEMMs on tcgen05 (see gemm_tc.cuh).
//
// Split-fp16 arithmetic: x * s = hi + lo with hi = fp16(x s), lo = fp16(x s - hi): 22 significant bits as
// long as x s sits well inside the fp16 range, which the power-of-two scale s guarantees (max |x s| in
// [64, 128): lo keeps full precision down to 2^-14, i.e. 2^-21 of the operand's maximum, and degrades
// gracefully through the fp16 subnormals below that).  A product needs three MMAs:
//     main += A_hi B_hi          corr += A_hi B_lo + A_lo B_hi          (A_lo B_lo ~ 2^-22: dropped)
// `main` and `corr` are separate fp32 accumulators in TMEM, added once in the epilogue, so the small
// correction terms are not absorbed one by one into the large sum.
//
//  tcg_rows_kernel  C[m, n] = epi(A[m, :] W): one CTA owns a 128-column half of W, converted ONCE into
//      shared memory as resident UMMA B operands (hi + lo, K-major SW128, <= 160 KB), and walks over
//      128-row tiles of A: eight producer warps load fp32 rows (coalesced), split them and write the
//      swizzled A operand chunks into a small ring, one thread issues the MMAs, four epilogue warps read
//      the double-buffered accumulators (tcgen05.ld) and store fp32 rows.  HBM-bound by design: every
//      element of A and C crosses HBM once (the sibling CTA of the other column half re-reads A from L2).
//  tcg_tn_kernel    C += A^T B over K = samples: both operands are converted into MN-major SW128 blocks
//      (the byte image of a K-major block read transposed, as in mlp_tc_bwd.cu), the 128 x N result stays
//      in TMEM over the CTA's whole sample range and is added to global memory once; the producers also
//      form the column sums of B (bias gradients) on the way.
#include <cuda.h>
#include <cuda_fp16.h>

#include "gemm_tc.cuh"
#include "sm100_ptx.cuh"

namespace lnrf {

using namespace ptx;

namespace {

// instruction descriptor, kind::f16 with fp16 A / B (format 0), fp32 D; K-major operands
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_f16_mn(int M, int N) { return idesc_f16(M, N) | (1u << 15) | (1u << 16); }

// power of two s with amax * s in [64, 128); 1 for amax = 0 / inf / nan
__device__ __forceinline__ float pow2_scale(float amax) {
  if (!(amax > 0.0f) || !(amax < 3.0e38f)) return 1.0f;
  int e = int((__float_as_uint(amax) >> 23) & 0xffu) - 127;
  int se = 6 - e;
  se = se < -60 ? -60 : (se > 60 ? 60 : se);
  return __uint_as_float(uint32_t(se + 127) << 23);
}

__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// eight scaled floats -> 16 bytes of hi and 16 bytes of lo
__device__ __forceinline__ void split8_store(const float4& a, const float4& b, float s, uint32_t hi_addr, uint32_t lo_addr) {
  uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
  split2(a.x * s, a.y * s, h0, l0);
  split2(a.z * s, a.w * s, h1, l1);
  split2(b.x * s, b.y * s, h2, l2);
  split2(b.z * s, b.w * s, h3, l3);
  st_shared_v4(hi_addr, h0, h1, h2, h3);
  st_shared_v4(lo_addr, l0, l1, l2, l3);
}

// 2-D tiled TMA load (box = [128 rows x 64 fp32]) into shared memory, completion counted in bytes on `bar`;
// rows / columns outside the tensor arrive as zeros (UTMALDG in SASS)
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tm, int32_t col, int32_t row, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst_smem), "l"(tm), "r"(col), "r"(row), "r"(bar)
      : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A B with the two shared-memory descriptors given as 32-bit halves: the high words are
// constant per layout, the low words are start addresses that the issue loop only increments (building
// 64-bit descriptors per MMA in the single issuing thread costs more than the MMAs themselves).
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ================================================================ row GEMM
constexpr int kRgThreads = 576;            // warps 0-7 epilogue, 8-15 converters, 16 MMA issuer, 17 TMA issuer
constexpr int kRgProducers = 256;
constexpr int kRgMaxChunks = 5;            // K <= 320
constexpr uint32_t kRgBlock = 128 * 128;   // one [128 x 64] fp16 K-major SW128 block
constexpr uint32_t kRgChunk = 2 * kRgBlock;  // hi + lo
constexpr uint32_t kRgData = 7 * kRgChunk;   // resident B chunks first, the A ring behind them
constexpr uint32_t kRgSmem = kRgData + 256;
// barrier block (byte offsets inside it)
constexpr uint32_t kRgAFull = 0, kRgAEmpty = 32, kRgAccFull = 64, kRgAccEmpty = 80, kRgTmemSlot = 96, kRgBMax = 100,
                   kRgTmaFull = 104;

struct RgChunkDesc {
  int seg;         // which K segment (tensor map) of A
  int k0;          // first column inside the segment
  int kv;          // valid K columns (multiple of 4, <= 64)
  int bk0;         // first K index of W
};
struct RgArgs {
  alignas(64) CUtensorMap tm[2];  // the two K segments of A as [M rows x K_seg] fp32 tensors, box 128 x 64
  RgChunkDesc ch[kRgMaxChunks];
  int nchunks, stages, has_a1;
  const float* B;
  int ldb, btrans;
  float* C;
  int ldc;
  int64_t M;
  int N, nhalves, nc;  // nc = accumulator columns per CTA (multiple of 16)
  const float* bias;
  const float* aux;
  int ldaux;
  const float* r1s;
  const float* r1w;
  const float* a_amax;
  const float* a1_amax;
  float* c_amax;
  const uint32_t* mask_in;  // TCG_MASKBITS: ReLU bit masks written by an earlier TCG_BIAS_RELU call (same M, N)
  uint32_t* mask_out;       // TCG_BIAS_RELU: receives the bits [C > 0] (nullable)
};

template <int EPI>
__global__ void __launch_bounds__(kRgThreads, 1) tcg_rows_kernel(const __grid_constant__ RgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();
  const uint32_t sB = sbase, sA = sbase + uint32_t(a.nchunks) * kRgChunk, bars = sbase + kRgData;
  volatile uint32_t* bar_words = reinterpret_cast<volatile uint32_t*>(smem_raw + kRgData);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = int(blockIdx.x) % a.nhalves, n0 = half * 128;
  const int64_t tiles = ceil_div(a.M, 128);
  const int64_t tstride = int64_t(gridDim.x) / a.nhalves, tfirst = int64_t(blockIdx.x) / a.nhalves;
  const int64_t my_tiles = tiles > tfirst ? (tiles - tfirst + tstride - 1) / tstride : 0;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(bars + kRgAFull + 8 * s, kRgProducers / 32);
      mbar_init(bars + kRgAEmpty + 8 * s, 1);
      mbar_init(bars + kRgTmaFull + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars + kRgAccFull + 8 * b, 1);
      mbar_init(bars + kRgAccEmpty + 8 * b, 8);
    }
    bar_words[kRgBMax / 4] = 0u;
    fence_barrier_init();
  }
  if (warp == 16) {
    tmem_alloc(bars + kRgTmemSlot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bar_words[kRgTmemSlot / 4];

  // ---- resident W operand of this column half.  ONE pass over global memory (every CTA reads the same W, so the
  // pass is bound by the L2's broadcast bandwidth: the earlier max-then-convert version read W twice, ~3.7 us per
  // 64-row chunk): each chunk lands as raw fp32 in its own 32 KB of shared memory ([64 k][128 n] for W[k][n],
  // [128 n][64 k] for Wt[n][k]; zero outside the matrix) while max |w| is taken, then it is split in place into
  // the hi / lo blocks (all of a chunk's items are read into registers before any is written).
  const int bt = a.btrans;
  const int ncols = a.N - n0 < 128 ? a.N - n0 : 128;
  float* rawf = reinterpret_cast<float*>(smem_raw);
  {
    float mx = 0.0f;
    for (int c = 0; c < a.nchunks; ++c) {
      const RgChunkDesc cd = a.ch[c];
      float4* raw4 = reinterpret_cast<float4*>(rawf + c * (kRgChunk / 4));
      if (!bt) {  // rows k = bk0 .. bk0 + kv - 1, 128 consecutive n from n0
        const float* base = a.B + int64_t(cd.bk0) * a.ldb + n0;
#pragma unroll 4
        for (int i = tid; i < 64 * 32; i += kRgThreads) {
          const int k = i >> 5, q = i & 31;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < cd.kv && 4 * q < ncols) v = __ldg(reinterpret_cast<const float4*>(base + int64_t(k) * a.ldb) + q);
          mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
          raw4[i] = v;
        }
      } else {    // rows n = n0 .. n0 + ncols - 1, kv consecutive k from bk0
        const float* base = a.B + int64_t(n0) * a.ldb + cd.bk0;
#pragma unroll 4
        for (int i = tid; i < 128 * 16; i += kRgThreads) {
          const int n = i >> 4, q = i & 15;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n < ncols && 4 * q < cd.kv) v = __ldg(reinterpret_cast<const float4*>(base + int64_t(n) * a.ldb) + q);
          mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
          raw4[i] = v;
        }
      }
    }
    mx = warp_max(mx);
    if (lane == 0) atomicMax(const_cast<uint32_t*>(bar_words + kRgBMax / 4), __float_as_uint(mx));
  }
  __syncthreads();
  const float sb = pow2_scale(__uint_as_float(bar_words[kRgBMax / 4]));
  for (int c = 0; c < a.nchunks; ++c) {
    // item -> (row n, 8-column group g8); 1024 items per chunk, at most two per thread
    const float* raw = rawf + c * (kRgChunk / 4);
    float v[2][8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int item = tid + h * kRgThreads;
      if (item < 1024) {
        const int n = bt ? item >> 3 : item & 127, g8 = bt ? item & 7 : item >> 7;
        if (bt) {
          const float4 x0 = *reinterpret_cast<const float4*>(raw + n * 64 + g8 * 8);
          const float4 x1 = *reinterpret_cast<const float4*>(raw + n * 64 + g8 * 8 + 4);
          v[h][0] = x0.x; v[h][1] = x0.y; v[h][2] = x0.z; v[h][3] = x0.w;
          v[h][4] = x1.x; v[h][5] = x1.y; v[h][6] = x1.z; v[h][7] = x1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[h][j] = raw[(g8 * 8 + j) * 128 + n];
        }
      }
    }
    __syncthreads();  // every raw value of this chunk is in registers
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int item = tid + h * kRgThreads;
      if (item < 1024) {
        const int n = bt ? item >> 3 : item & 127, g8 = bt ? item & 7 : item >> 7;
        const uint32_t off = uint32_t(n) * 128u + (uint32_t((g8 ^ (n & 7)) & 7) << 4);
        split8_store(make_float4(v[h][0], v[h][1], v[h][2], v[h][3]), make_float4(v[h][4], v[h][5], v[h][6], v[h][7]), sb,
                     sB + c * kRgChunk + off, sB + c * kRgChunk + kRgBlock + off);
      }
    }
  }
  fence_proxy_async_smem();
  __syncthreads();

  // one scale for both K segments (they share the accumulator): from the larger of the two maxima; a segment
  // without an amax counts as O(1) once the other one has one
  float sa = 1.0f;
  if (a.a_amax != nullptr || a.a1_amax != nullptr) {
    const float m0 = a.a_amax ? __ldg(a.a_amax) : 1.0f;
    const float m1 = a.a1_amax ? __ldg(a.a1_amax) : (a.has_a1 ? 1.0f : 0.0f);
    sa = pow2_scale(fmaxf(m0, m1));
  }

  if (warp == 17) {
    // ===== TMA issuer: one 128 x 64 fp32 box of A per ring slot, as soon as the MMAs that read the slot retired.
    // (Loading A through registers was bounded by the L1's outstanding-request capacity: issuing the next chunk's
    // four 256-bit loads stalled ~2 k clk per chunk, i.e. ~3 TB/s, however deep the register ring.)
    if (lane == 0) {
      uint32_t stage = 0, phases = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int32_t row = int32_t((tfirst + it * tstride) * 128);
        for (int c = 0; c < a.nchunks; ++c) {
          mbar_wait(bars + kRgAEmpty + 8 * stage, ((phases >> stage) & 1u) ^ 1u);
          phases ^= 1u << stage;
          mbar_arrive_expect_tx(bars + kRgTmaFull + 8 * stage, kRgChunk);
          tma_load_2d(sA + stage * kRgChunk, &a.tm[a.ch[c].seg], a.ch[c].k0, row, bars + kRgTmaFull + 8 * stage);
          if (++stage == uint32_t(a.stages)) stage = 0;
        }
      }
    }
  } else if (warp >= 8 && warp < 16) {
    // ===== converters: the slot arrives as raw fp32 rows (256 B each); every thread reads its four 32-byte pieces
    // (row, 8 columns), all 256 threads meet, then the hi / lo fp16 operand blocks are written IN PLACE.
    const int pt = tid - 256;
    const int g8 = pt & 7, r0 = pt >> 3;  // rows r0 + 32 j
    const uint32_t raw0 = uint32_t(r0) * 256u + uint32_t(g8) * 32u;
    const uint32_t off0 = uint32_t(r0) * 128u + (uint32_t((g8 ^ (r0 & 7)) & 7) << 4);  // rows r0 + 32 j share r & 7
    const int64_t total = my_tiles * a.nchunks;
    uint32_t stage = 0, phases = 0;
    for (int64_t q = 0; q < total; ++q) {
      mbar_wait(bars + kRgTmaFull + 8 * stage, (phases >> stage) & 1u);
      phases ^= 1u << stage;
      const uint32_t base = sA + stage * kRgChunk;
      float4 v[8];
      // The eight lanes of a row read 16-byte units {0,2,4,6,9,11,13,15} first and {1,3,5,7,8,10,12,14} second:
      // each instruction covers all 32 banks once (reading both halves in piece order is a 2-way conflict, and
      // shared-memory bandwidth is this kernel's limiter)
      const uint32_t hi_half = g8 >= 4 ? 16u : 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 p0 = lds_f4(base + raw0 + j * 8192u + hi_half);
        const float4 p1 = lds_f4(base + raw0 + j * 8192u + (16u - hi_half));
        v[2 * j] = hi_half ? p1 : p0;
        v[2 * j + 1] = hi_half ? p0 : p1;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");  // every raw piece is in registers before any is overwritten
#pragma unroll
      for (int j = 0; j < 4; ++j)
        split8_store(v[2 * j], v[2 * j + 1], sa, base + off0 + j * 4096u, base + kRgBlock + off0 + j * 4096u);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + kRgAFull + 8 * stage);
      if (++stage == uint32_t(a.stages)) stage = 0;
    }
  } else if (warp == 16) {
    // ===== MMA issuer
    const uint32_t idesc = idesc_f16(128, a.nc), idesc256 = idesc_f16(128, 256);
    uint32_t stage = 0, phases = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const uint32_t buf = uint32_t(it & 1);
      mbar_wait(bars + kRgAccEmpty + 8 * buf, (uint32_t(it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_main = tmem + buf * 256u, d_corr = d_main + 128u;
      for (int c = 0; c < a.nchunks; ++c) {
        mbar_wait(bars + kRgAFull + 8 * stage, (phases >> stage) & 1u);
        phases ^= 1u << stage;
        tc_fence_after();
        if (elect_one_sync()) {
          const int ksteps = (a.ch[c].kv + 15) >> 4;
          const uint32_t ah = (((sA + stage * kRgChunk) & 0x3FFFFu) >> 4) | (1u << 16), al = ah + (kRgBlock >> 4);
          const uint32_t bh = (((sB + uint32_t(c) * kRgChunk) & 0x3FFFFu) >> 4) | (1u << 16);  // hi block, lo block behind it
          for (int k = 0; k < ksteps; ++k) {  // 32 bytes (16 fp16) of K per step
            const uint32_t acc = (c | k) ? 1u : 0u;
            // [main | corr] (+)= A_hi [B_hi ; B_lo]^T as ONE N = 256 MMA (the lo block follows the hi block in
            // shared memory, the corr columns follow the main columns in TMEM): A_hi is read once instead of twice
            umma_f16_lohi(d_main, ah + 2 * k, bh + 2 * k, kDescHiSw128, idesc256, acc);
            umma_f16_lohi(d_corr, al + 2 * k, bh + 2 * k, kDescHiSw128, idesc, 1u);
          }
          umma_commit(bars + kRgAEmpty + 8 * stage);
          if (c == a.nchunks - 1) umma_commit(bars + kRgAccFull + 8 * buf);
        }
        __syncwarp();
        if (++stage == uint32_t(a.stages)) stage = 0;
      }
    }
  } else if (warp < 8) {
    // ===== epilogue (eight warps: TMEM lane quadrant warp & 3, accumulator column half warp >> 2: one warp per
    // scheduler needs ~7 k clk per tile, a single dependent instruction chain, against 3.4 k clk of MMAs).  tcgen05.ld hands lane r the 32 columns of tile row r: storing that directly makes every
    // warp store touch 32 different lines, and the LSU (not HBM) bounds the kernel (measured: 1.8 TB/s).  So
    // each 8-lane group first transposes its 8 rows x 8 four-column granules through shuffles: lane (a, b) =
    // (lane / 8, lane % 8) ends up with granule b of rows 8a .. 8a+7, and store instruction i writes rows
    // {8a + i}: eight lanes cover 128 contiguous bytes of a row.  The mask rows are read the same way.
    const float inv = 1.0f / (sa * sb);
    float amax = 0.0f;
    const int la = lane >> 3, lb = lane & 7;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const uint32_t buf = uint32_t(it & 1);
      const int64_t row0 = (tfirst + it * tstride) * 128 + (warp & 3) * 32 + la * 8;  // rows row0 + i, i = 0..7
      float r1[8];
      if (EPI == TCG_RANK1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r1[i] = row0 + i < a.M ? __ldg(a.r1s + row0 + i) : 0.0f;
      }
      // columns of this warp: 64 per warp quad-pair member, 32 when the half has <= 64 columns (the 64-wide
      // layers of the hash-grid heads: otherwise warps 4-7 would idle and one warp per scheduler drains the tile)
      const int cw = a.nc <= 64 ? 32 : 64;
      const int c_beg = (warp >> 2) * cw, c_end = a.nc < c_beg + cw ? a.nc : c_beg + cw;
      // bit masks: one word per lane and 32 x 32 block = [rows row0 .. row0+7] x [columns n .. n+3], bit 4 i + e;
      // words exist for the rows below ceil(M / 32) * 32.  Fetched BEFORE the wait on the accumulator (the load
      // latency would otherwise sit in the drain chain of every tile).
      const bool mrow_ok = row0 < ((a.M + 31) & ~int64_t(31));
      const int64_t mword0 = ((row0 >> 5) * ((a.N + 31) >> 5) + ((n0 + c_beg) >> 5)) * 32 + lane;
      uint32_t mpre0 = 0, mpre1 = 0;
      if (EPI == TCG_MASKBITS && mrow_ok) {
        if (c_beg < c_end) mpre0 = __ldg(a.mask_in + mword0);
        if (c_beg + 32 < c_end) mpre1 = __ldg(a.mask_in + mword0 + 32);
      }
      mbar_wait(bars + kRgAccFull + 8 * buf, uint32_t(it >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_main = tmem + (uint32_t((warp & 3) * 32) << 16) + buf * 256u;
      for (int c0 = c_beg; c0 < c_end; c0 += 32) {
        uint32_t vm[32], vc[32];
        tmem_ld32(t_main + c0, vm);
        tmem_ld32(t_main + 128 + c0, vc);
        const int n = n0 + c0 + 4 * lb;               // this lane's four output columns
        const bool col_ok = c0 + 4 * lb < a.nc && n < a.N;
        float4 mk[8];  // EPI_MASK: the mask values of this lane's eight outputs, all loads in flight at once
        float4 bw = make_float4(0.f, 0.f, 0.f, 0.f);  // bias (or the rank-1 row vector) of the four columns
        const int64_t mword = mword0 + (c0 - c_beg);
        uint32_t mbits = EPI == TCG_MASKBITS ? (c0 == c_beg ? mpre0 : mpre1) : 0u;
        if (EPI == TCG_MASK) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            mk[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col_ok && row0 + i < a.M) mk[i] = __ldg(reinterpret_cast<const float4*>(a.aux + (row0 + i) * a.ldaux + n));
          }
        } else if (EPI == TCG_BIAS_RELU || EPI == TCG_BIAS) {
          if (col_ok) bw = __ldg(reinterpret_cast<const float4*>(a.bias + n));
        } else if (EPI == TCG_RANK1) {
          if (col_ok) bw = __ldg(reinterpret_cast<const float4*>(a.r1w + n));
        }
        tmem_wait_ld();
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = (__uint_as_float(vm[j]) + __uint_as_float(vc[j])) * inv;
        // 8 x 8 transpose of granules (x[4g .. 4g+3] = granule g of this lane's row) inside each 8-lane group
#pragma unroll
        for (int s2 = 1; s2 < 8; s2 <<= 1) {
          const bool up = (lb & s2) != 0;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (g & s2) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lo_v = x[4 * g + e], hi_v = x[4 * (g + s2) + e];
              const float recv = __shfl_xor_sync(0xffffffffu, up ? lo_v : hi_v, s2);
              x[4 * g + e] = up ? recv : lo_v;
              x[4 * (g + s2) + e] = up ? hi_v : recv;
            }
          }
        }
        // x[4i .. 4i+3] = granule lb (columns n .. n+3) of row row0 + i
        if (col_ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (row0 + i >= a.M) continue;
            float v[4] = {x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]};
            if (EPI == TCG_BIAS_RELU || EPI == TCG_BIAS) {
              v[0] += bw.x; v[1] += bw.y; v[2] += bw.z; v[3] += bw.w;
              if (EPI == TCG_BIAS_RELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  v[e] = fmaxf(v[e], 0.0f);
                  mbits |= (v[e] > 0.0f ? 1u : 0u) << (4 * i + e);
                }
              }
            } else if (EPI == TCG_MASKBITS) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = (mbits >> (4 * i + e)) & 1u ? v[e] : 0.0f;
            } else if (EPI == TCG_MASK) {
              v[0] = mk[i].x > 0.f ? v[0] : 0.f; v[1] = mk[i].y > 0.f ? v[1] : 0.f;
              v[2] = mk[i].z > 0.f ? v[2] : 0.f; v[3] = mk[i].w > 0.f ? v[3] : 0.f;
            } else if (EPI == TCG_RANK1) {
              v[0] = fmaf(r1[i], bw.x, v[0]); v[1] = fmaf(r1[i], bw.y, v[1]);
              v[2] = fmaf(r1[i], bw.z, v[2]); v[3] = fmaf(r1[i], bw.w, v[3]);
            }
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))));
            *reinterpret_cast<float4*>(a.C + (row0 + i) * a.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
        if (EPI == TCG_BIAS_RELU && a.mask_out != nullptr && mrow_ok) a.mask_out[mword] = mbits;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + kRgAccEmpty + 8 * buf);
    }
    if (a.c_amax) {
      amax = warp_max(amax);
      if (lane == 0 && amax > 0.0f && amax < 3.0e38f) atomicMax(reinterpret_cast<uint32_t*>(a.c_amax), __float_as_uint(amax));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

// ================================================================ TN GEMM (K = samples)
// 16 producer warps: with 8 (two per scheduler, each thread holding a whole 24-load chunk) the producers' own
// instruction stream (~1.2 k instructions per warp and chunk) and the exposed latency of their loads bounded the
// kernel (0.39 ms per fine-level 256 x 256 dW, of which 0.15 ms remained with loads, conversion and MMAs all
// removed; a CTA pair with cta_group::2 MMAs that loads B once instead of twice ran at the same 0.39 ms).
// 16 warps: 0.345 ms (fp32 NeRF step -3 %, Ref-NeRF -1.3 % on the same box).
// (The 64-wide SMALL variant below keeps 8: its chunks are 192 samples and 17 warps cap it at 96 registers, which it spills.)
constexpr int kTnProdBig = 16, kTnProdSmall = 8;  // producer warps (0-3 also drain the accumulator); + the MMA issuer warp
constexpr uint32_t kTnHalf = 64 * 128;      // one [64 samples x 64 features] fp16 block
constexpr uint32_t kTnAPart = 2 * kTnHalf;  // A hi (or lo): features 0..127 of the CTA's M block
constexpr uint32_t kTnBPart = 4 * kTnHalf;  // B hi (or lo): up to 256 columns
constexpr uint32_t kTnStage = 2 * kTnAPart + 2 * kTnBPart;  // 96 KB
constexpr int kTnStages = 2;
constexpr uint32_t kTnSmem = kTnStages * kTnStage + 256;
constexpr uint32_t kTnFull = 0, kTnEmpty = 16, kTnDone = 32, kTnTmemSlot = 40;
// SMALL variant (M <= 64 and N <= 64: the 64-wide Instant-NGP heads): a chunk of 64 samples would be 32 KB, and
// with one chunk of loads in flight per CTA the kernel sat at 2.3 TB/s.  There a chunk is 192 samples of ONE
// 64-feature block per operand (4 x 24 KB = the same 96 KB stage, the same 24 loads per thread); the M = 128 MMA
// reads the rows 64.. of the A part as "features 64..127": finite values into accumulator lanes nobody reads.
constexpr int kTnSmallRows = 192;
constexpr uint32_t kTnSmallPart = kTnSmallRows * 128;
static_assert(4 * kTnSmallPart == kTnStage, "small-shape stage layout");

struct TnArgs {
  const float* At;
  int lda, M;
  const float* B;
  int ldb, N;
  int64_t K;
  float* C;
  int ldc;
  float* db;
  const float* a_amax;
  const float* b_amax;
  int mblocks, nb;  // nb = 64-column blocks of B (N padded)
  int64_t chunks_per_cta, chunks;
};

template <bool SMALL, int kTnProd>
__global__ void __launch_bounds__(32 * kTnProd + 32, 1) tcg_tn_kernel(const __grid_constant__ TnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();
  const uint32_t bars = sbase + kTnStages * kTnStage;
  volatile uint32_t* bar_words = reinterpret_cast<volatile uint32_t*>(smem_raw + kTnStages * kTnStage);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mb = int(blockIdx.x) % a.mblocks;
  const int64_t ks = int64_t(blockIdx.x) / a.mblocks;
  const int64_t q0 = ks * a.chunks_per_cta;
  const int64_t q1 = q0 + a.chunks_per_cta < a.chunks ? q0 + a.chunks_per_cta : a.chunks;
  const int64_t nq = q1 > q0 ? q1 - q0 : 0;
  const int Npad = a.nb * 64;

  if (tid == 0) {
    for (int s = 0; s < kTnStages; ++s) {
      mbar_init(bars + kTnFull + 8 * s, kTnProd);
      mbar_init(bars + kTnEmpty + 8 * s, 1);
    }
    mbar_init(bars + kTnDone, 1);
    fence_barrier_init();
  }
  if (warp == kTnProd) {
    tmem_alloc(bars + kTnTmemSlot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bar_words[kTnTmemSlot / 4];
  const float sa = a.a_amax ? pow2_scale(__ldg(a.a_amax)) : 1.0f;
  const float sb = a.b_amax ? pow2_scale(__ldg(a.b_amax)) : 1.0f;

  if (warp < kTnProd) {
    // ===== producers
    const int ag8 = tid & 15, ar0 = tid >> 4;  // A pieces: features mb 128 + 8 ag8 ..
    const int bg8 = tid & 31, br0 = tid >> 5;  // B pieces: columns 8 bg8 ..  (fixed columns per thread)
    const int af = mb * 128 + ag8 * 8, bc = bg8 * 8;
    const bool b_active = bg8 < a.nb * 8;
    float cs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = 0.0f;
    const bool do_db = a.db != nullptr && mb == 0;
    uint32_t stage = 0, phases = 0;
    if constexpr (SMALL) {
      // pieces of 8 floats: p = tid + 32 kTnProd j -> sample p / 8, features / columns 8 (tid & 7) .. (fixed per thread)
      const int g8 = tid & 7, f0 = g8 * 8;
      for (int64_t q = 0; q < nq; ++q) {
        const int64_t k0 = (q0 + q) * kTnSmallRows;
        constexpr int kPieces = 48 / kTnProd;
        float4 va[2 * kPieces], vb[2 * kPieces];
#pragma unroll
        for (int j = 0; j < kPieces; ++j) {
          const int64_t row = k0 + (tid >> 3) + 4 * kTnProd * j;
          va[2 * j] = va[2 * j + 1] = vb[2 * j] = vb[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < a.K) {
            const float* pa = a.At + row * a.lda + f0;
            if (f0 < a.M) va[2 * j] = __ldg(reinterpret_cast<const float4*>(pa));
            if (f0 + 4 < a.M) va[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(pa + 4));
            const float* pb = a.B + row * a.ldb + f0;
            if (f0 < a.N) vb[2 * j] = __ldg(reinterpret_cast<const float4*>(pb));
            if (f0 + 4 < a.N) vb[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(pb + 4));
          }
        }
        mbar_wait(bars + kTnEmpty + 8 * stage, ((phases >> stage) & 1u) ^ 1u);
        phases ^= 1u << stage;
        const uint32_t s_ah = sbase + stage * kTnStage, s_al = s_ah + kTnSmallPart;
        const uint32_t s_bh = s_al + kTnSmallPart, s_bl = s_bh + kTnSmallPart;
#pragma unroll
        for (int j = 0; j < kPieces; ++j) {
          const int r = (tid >> 3) + 4 * kTnProd * j;
          const uint32_t off = uint32_t(r) * 128u + (uint32_t((g8 ^ (r & 7)) & 7) << 4);
          split8_store(va[2 * j], va[2 * j + 1], sa, s_ah + off, s_al + off);
          split8_store(vb[2 * j], vb[2 * j + 1], sb, s_bh + off, s_bl + off);
          if (do_db) {
            cs[0] += vb[2 * j].x; cs[1] += vb[2 * j].y; cs[2] += vb[2 * j].z; cs[3] += vb[2 * j].w;
            cs[4] += vb[2 * j + 1].x; cs[5] += vb[2 * j + 1].y; cs[6] += vb[2 * j + 1].z; cs[7] += vb[2 * j + 1].w;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + kTnFull + 8 * stage);
        if (++stage == kTnStages) stage = 0;
      }
    } else
    // One chunk = 4 + 8 sixteen-byte loads per thread, all issued before the first use (96 KB in flight per SM).
    for (int64_t q = 0; q < nq; ++q) {
      const int64_t k0 = (q0 + q) * 64;
      constexpr int kPa = 32 / kTnProd, kPb = 64 / kTnProd;  // pieces per thread: A rows ar0 + 2 kTnProd j, B rows br0 + kTnProd j
      float4 va[2 * kPa], vb[2 * kPb];
#pragma unroll
      for (int j = 0; j < kPa; ++j) {
        const int64_t row = k0 + ar0 + 2 * kTnProd * j;
        va[2 * j] = make_float4(0.f, 0.f, 0.f, 0.f);
        va[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < a.K) {
          const float* p = a.At + row * a.lda + af;
          if (af < a.M) va[2 * j] = __ldg(reinterpret_cast<const float4*>(p));
          if (af + 4 < a.M) va[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(p + 4));
        }
      }
#pragma unroll
      for (int j = 0; j < kPb; ++j) {
        const int64_t row = k0 + br0 + kTnProd * j;
        vb[2 * j] = make_float4(0.f, 0.f, 0.f, 0.f);
        vb[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b_active && row < a.K) {
          const float* p = a.B + row * a.ldb + bc;
          if (bc < a.N) vb[2 * j] = __ldg(reinterpret_cast<const float4*>(p));
          if (bc + 4 < a.N) vb[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(p + 4));
        }
      }
      mbar_wait(bars + kTnEmpty + 8 * stage, ((phases >> stage) & 1u) ^ 1u);
      phases ^= 1u << stage;
      const uint32_t s_ah = sbase + stage * kTnStage, s_al = s_ah + kTnAPart;
      const uint32_t s_bh = s_al + kTnAPart, s_bl = s_bh + kTnBPart;
#pragma unroll
      for (int j = 0; j < kPa; ++j) {
        const int r = ar0 + 2 * kTnProd * j;
        const uint32_t off = uint32_t(ag8 >> 3) * kTnHalf + uint32_t(r) * 128u + (uint32_t(((ag8 & 7) ^ (r & 7)) & 7) << 4);
        split8_store(va[2 * j], va[2 * j + 1], sa, s_ah + off, s_al + off);
      }
      if (b_active) {
#pragma unroll
        for (int j = 0; j < kPb; ++j) {
          const int r = br0 + kTnProd * j;
          const uint32_t off = uint32_t(bg8 >> 3) * kTnHalf + uint32_t(r) * 128u + (uint32_t(((bg8 & 7) ^ (r & 7)) & 7) << 4);
          split8_store(vb[2 * j], vb[2 * j + 1], sb, s_bh + off, s_bl + off);
          if (do_db) {
            cs[0] += vb[2 * j].x; cs[1] += vb[2 * j].y; cs[2] += vb[2 * j].z; cs[3] += vb[2 * j].w;
            cs[4] += vb[2 * j + 1].x; cs[5] += vb[2 * j + 1].y; cs[6] += vb[2 * j + 1].z; cs[7] += vb[2 * j + 1].w;
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + kTnFull + 8 * stage);
      if (++stage == kTnStages) stage = 0;
    }
    // wait for the last MMAs: shared memory and the accumulator are then free / complete
    mbar_wait(bars + kTnDone, 0);
    tc_fence_after();
    if (do_db) {  // fold the row groups of a column (tid >> 5 resp. tid >> 3) through shared memory, one atomic per column
      float* s_cs = reinterpret_cast<float*>(smem_raw);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kTnProd) : "memory");
      if constexpr (SMALL) {  // 4 kTnProd row groups x 64 columns
#pragma unroll
        for (int j = 0; j < 8; ++j) s_cs[(tid >> 3) * 64 + (tid & 7) * 8 + j] = cs[j];
      } else if (b_active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s_cs[br0 * 256 + bc + j] = cs[j];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kTnProd) : "memory");
      if (tid < a.N) {
        float t = 0.0f;
        if constexpr (SMALL) {
#pragma unroll 8
          for (int g = 0; g < 4 * kTnProd; ++g) t += s_cs[g * 64 + tid];
        } else {
#pragma unroll
          for (int g = 0; g < kTnProd; ++g) t += s_cs[g * 256 + tid];
        }
        atomicAdd(a.db + tid, t);
      }
    }
    if (warp < 4 && nq > 0) {
      const float inv = 1.0f / (sa * sb);
      const int row = mb * 128 + tid;
      const uint32_t t0 = tmem + (uint32_t(warp * 32) << 16);
      for (int c0 = 0; c0 < Npad; c0 += 32) {
        uint32_t vm[32], vc[32];
        tmem_ld32(t0 + c0, vm);
        tmem_ld32(t0 + 256 + c0, vc);
        tmem_wait_ld();
        if (row < a.M) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (c0 + j < a.N)
              red_add_v4(a.C + int64_t(row) * a.ldc + c0 + j,
                         (__uint_as_float(vm[j]) + __uint_as_float(vc[j])) * inv,
                         (__uint_as_float(vm[j + 1]) + __uint_as_float(vc[j + 1])) * inv,
                         (__uint_as_float(vm[j + 2]) + __uint_as_float(vc[j + 2])) * inv,
                         (__uint_as_float(vm[j + 3]) + __uint_as_float(vc[j + 3])) * inv);
          }
        }
      }
    }
  } else {
    // ===== MMA issuer: main += Ah^T Bh, corr += Ah^T Bl + Al^T Bh; operands MN-major, K = 16 samples per MMA
    const uint32_t idesc = idesc_f16_mn(128, Npad);
    uint32_t stage = 0, phases = 0;
    for (int64_t q = 0; q < nq; ++q) {
      mbar_wait(bars + kTnFull + 8 * stage, (phases >> stage) & 1u);
      phases ^= 1u << stage;
      tc_fence_after();
      if (elect_one_sync()) {
        // MN-major SW128: 64-feature blocks kTnHalf bytes apart (LBO), 16 samples = 2048 bytes per K step
        const uint32_t lbo = (kTnHalf >> 4) << 16;
        constexpr uint32_t kAP = SMALL ? kTnSmallPart : kTnAPart, kBP = SMALL ? kTnSmallPart : kTnBPart;
        const uint32_t ah = (((sbase + stage * kTnStage) & 0x3FFFFu) >> 4) | lbo, al = ah + (kAP >> 4);
        const uint32_t bh = al + (kAP >> 4), bl = bh + (kBP >> 4);
#pragma unroll
        for (int k = 0; k < (SMALL ? kTnSmallRows / 16 : 4); ++k) {
          const uint32_t acc = (q | k) ? 1u : 0u;
          umma_f16_lohi(tmem, ah + 128 * k, bh + 128 * k, kDescHiSw128, idesc, acc);
          umma_f16_lohi(tmem + 256, ah + 128 * k, bl + 128 * k, kDescHiSw128, idesc, acc);
          umma_f16_lohi(tmem + 256, al + 128 * k, bh + 128 * k, kDescHiSw128, idesc, 1u);
        }
        umma_commit(bars + kTnEmpty + 8 * stage);
        if (q == nq - 1) umma_commit(bars + kTnDone);
      }
      __syncwarp();
      if (++stage == kTnStages) stage = 0;
    }
    if (nq == 0 && elect_one_sync()) mbar_arrive(bars + kTnDone);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTnProd) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(256) tcg_amax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ amax) {
  float m = 0.0f;
  const int64_t n4 = n >> 2;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = fmaxf(m, fabsf(__ldg(x + (n4 << 2) + threadIdx.x)));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.0f && m < 3.0e38f) atomicMax(reinterpret_cast<uint32_t*>(amax), __float_as_uint(m));
}

// tensor map of a row-major fp32 [M x K] matrix with leading dimension lda, box = 128 rows x 64 columns
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;  // resolved once in init_gemm_tc (driver entry point: no libcuda link)
int make_tmap(CUtensorMap* tm, const float* A, int64_t M, int K, int lda) {
  LNRF_REQUIRE(g_encode_tiled != nullptr, LNRF_E_INVALID, "tcg_rows: call lnrf_init first");
  LNRF_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0, LNRF_E_INVALID, "tcg_rows: operand not 16-byte aligned");
  const cuuint64_t dims[2] = {cuuint64_t(K), cuuint64_t(M)};
  const cuuint64_t strides[1] = {cuuint64_t(lda) * 4u};
  const cuuint32_t box[2] = {64u, 128u};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LNRF_REQUIRE(r == CUDA_SUCCESS, LNRF_E_INVALID, "tcg_rows: cuTensorMapEncodeTiled failed (%d) M=%lld K=%d lda=%d", int(r),
               (long long)M, K, lda);
  return LNRF_OK;
}

template <int EPI>
int launch_rows(const RgArgs& a, unsigned grid, cudaStream_t st) {
  tcg_rows_kernel<EPI><<<grid, kRgThreads, kRgSmem, st>>>(a);
  LNRF_LAUNCH_CHECK("tcg_rows_kernel");
  return LNRF_OK;
}

}  // namespace

bool tcg_supported(int N, int K0, int K1) {
  if (N <= 0 || N > 256 || N % 4) return false;
  if (K0 <= 0 || K0 % 4 || K1 < 0 || K1 % 4) return false;
  return ceil_div(K0, 64) + ceil_div(K1, 64) <= kRgMaxChunks;
}

int tcg_rows(cudaStream_t st, int epi, bool btrans, int64_t M, int N, const float* A0, int lda0, int K0,
             const float* A1, int lda1, int K1, const float* B, int ldb, float* C, int ldc, const float* bias,
             const float* aux, int ldaux, const float* r1s, const float* r1w, const float* a_amax, const float* a1_amax, float* c_amax,
             const uint32_t* mask_in, uint32_t* mask_out) {
  LNRF_REQUIRE(tcg_supported(N, K0, K1), LNRF_E_UNSUPPORTED, "tcg_rows: N=%d K0=%d K1=%d", N, K0, K1);
  LNRF_REQUIRE(lda0 % 4 == 0 && lda1 % 4 == 0 && ldc % 4 == 0 && ldaux % 4 == 0, LNRF_E_UNSUPPORTED,
               "tcg_rows: leading dimensions must be multiples of 4");
  if (M <= 0) return LNRF_OK;
  RgArgs a{};
  int nch = 0;
  for (int k = 0; k < K0; k += 64) a.ch[nch++] = RgChunkDesc{0, k, K0 - k < 64 ? K0 - k : 64, k};
  for (int k = 0; k < K1; k += 64) a.ch[nch++] = RgChunkDesc{1, k, K1 - k < 64 ? K1 - k : 64, K0 + k};
  {
    int rc = make_tmap(&a.tm[0], A0, M, K0, lda0);
    if (rc) return rc;
    if (K1 > 0 && (rc = make_tmap(&a.tm[1], A1, M, K1, lda1))) return rc;
  }
  a.nchunks = nch;
  a.stages = 7 - nch > 4 ? 4 : 7 - nch;
  a.B = B; a.ldb = ldb; a.btrans = btrans ? 1 : 0;
  a.C = C; a.ldc = ldc; a.M = M; a.N = N;
  a.nhalves = N > 128 ? 2 : 1;
  const int ncols = a.nhalves == 2 ? 128 : N;
  a.nc = int(align_up(ncols, 16));
  a.bias = bias; a.aux = aux; a.ldaux = ldaux; a.r1s = r1s; a.r1w = r1w;
  a.a_amax = a_amax; a.a1_amax = K1 > 0 ? a1_amax : nullptr; a.c_amax = c_amax; a.has_a1 = K1 > 0 ? 1 : 0;
  a.mask_in = mask_in; a.mask_out = mask_out;
  LNRF_REQUIRE(epi != TCG_MASKBITS || mask_in != nullptr, LNRF_E_INVALID, "tcg_rows: TCG_MASKBITS without mask_in");
  const int64_t tiles = ceil_div(M, 128);
  int64_t grid = sm_count() / a.nhalves;
  if (grid > tiles) grid = tiles;
  grid *= a.nhalves;
  switch (epi) {
    case TCG_BIAS_RELU: return launch_rows<TCG_BIAS_RELU>(a, unsigned(grid), st);
    case TCG_BIAS: return launch_rows<TCG_BIAS>(a, unsigned(grid), st);
    case TCG_MASK: return launch_rows<TCG_MASK>(a, unsigned(grid), st);
    case TCG_RANK1: return launch_rows<TCG_RANK1>(a, unsigned(grid), st);
    case TCG_STORE: return launch_rows<TCG_STORE>(a, unsigned(grid), st);
    case TCG_MASKBITS: return launch_rows<TCG_MASKBITS>(a, unsigned(grid), st);
  }
  LNRF_REQUIRE(false, LNRF_E_INVALID, "tcg_rows: unknown epilogue %d", epi);
}

int tcg_tn_acc(cudaStream_t st, int M, int N, const float* At, int lda, const float* B, int ldb, int64_t K, float* C,
               int ldc, float* db, const float* a_amax, const float* b_amax) {
  LNRF_REQUIRE(M > 0 && M <= 256 && M % 4 == 0 && N > 0 && N <= 256 && N % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 &&
                   ldc % 4 == 0,
               LNRF_E_UNSUPPORTED, "tcg_tn_acc: M=%d N=%d lda=%d ldb=%d ldc=%d", M, N, lda, ldb, ldc);
  if (K <= 0) return LNRF_OK;
  TnArgs a{};
  a.At = At; a.lda = lda; a.M = M; a.B = B; a.ldb = ldb; a.N = N; a.K = K; a.C = C; a.ldc = ldc; a.db = db;
  a.a_amax = a_amax; a.b_amax = b_amax;
  a.mblocks = int(ceil_div(M, 128));
  a.nb = int(ceil_div(N, 64));
  const bool small = M <= 64 && N <= 64;
  a.chunks = ceil_div(K, small ? kTnSmallRows : 64);
  int64_t splits = sm_count() / a.mblocks;
  if (splits < 1) splits = 1;
  a.chunks_per_cta = ceil_div(a.chunks, splits);
  if (a.chunks_per_cta < 4) a.chunks_per_cta = 4;
  splits = ceil_div(a.chunks, a.chunks_per_cta);
  if (small) tcg_tn_kernel<true, kTnProdSmall><<<unsigned(splits * a.mblocks), 32 * kTnProdSmall + 32, kTnSmem, st>>>(a);
  else tcg_tn_kernel<false, kTnProdBig><<<unsigned(splits * a.mblocks), 32 * kTnProdBig + 32, kTnSmem, st>>>(a);
  LNRF_LAUNCH_CHECK("tcg_tn_kernel");
  return LNRF_OK;
}

int tcg_amax(cudaStream_t st, const float* x, int64_t n, float* amax) {
  if (n <= 0) return LNRF_OK;
  int64_t blocks = ceil_div(n, 256 * 16);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  tcg_amax_kernel<<<unsigned(blocks), 256, 0, st>>>(x, n, amax);
  LNRF_LAUNCH_CHECK("tcg_amax_kernel");
  return LNRF_OK;
}

int init_gemm_tc() {
  if (g_encode_tiled == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    LNRF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    LNRF_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, LNRF_E_UNSUPPORTED,
                 "lnrf_init: the driver does not export cuTensorMapEncodeTiled");
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  }
  LNRF_CUDA(cudaFuncSetAttribute(tcg_rows_kernel<TCG_BIAS_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRgSmem));
  LNRF_CUDA(cudaFuncSetAttribute(tcg_rows_kernel<TCG_BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRgSmem));
  LNRF_CUDA(cudaFuncSetAttribute(tcg_rows_kernel<TCG_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRgSmem));
  LNRF_CUDA(cudaFuncSetAttribute(tcg_rows_kernel<TCG_RANK1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRgSmem));
  LNRF_CUDA(cudaFuncSetAttribute(tcg_rows_kernel<TCG_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRgSmem));
  LNRF_CUDA(cudaFuncSetAttribute(tcg_rows_kernel<TCG_MASKBITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRgSmem));
  LNRF_CUDA(cudaFuncSetAttribute(tcg_tn_kernel<false, kTnProdBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTnSmem));
  LNRF_CUDA(cudaFuncSetAttribute(tcg_tn_kernel<true, kTnProdSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTnSmem));
  return LNRF_OK;
}

}  // namespace lnrf

// Diagnostic entry (tests): the GEMM engine on its own.  mode 0: rows NN, 1: rows NT, 2: TN accumulate
// (C[K0, N] += A0[M, K0]^T B[M, N]: M = samples), 3: c_amax = max |A0[0..M)|.
extern "C" int lnrf_tcgemm(int mode, int epi, int64_t M, int N, const float* A0, int lda0, int K0, const float* A1,
                           int lda1, int K1, const float* B, int ldb, float* C, int ldc, const float* bias,
                           const float* aux, int ldaux, const float* r1s, const float* r1w, float* db,
                           const float* a_amax, const float* b_amax, float* c_amax, const uint32_t* mask_in,
                           uint32_t* mask_out, lnrf_stream_t stream) {
  using namespace lnrf;
  cudaStream_t st = as_stream(stream);
  if (mode == 0 || mode == 1)
    return tcg_rows(st, epi, mode == 1, M, N, A0, lda0, K0, A1, lda1, K1, B, ldb, C, ldc, bias, aux, ldaux, r1s, r1w,
                    a_amax, nullptr, c_amax, mask_in, mask_out);
  if (mode == 2) return tcg_tn_acc(st, K0, N, A0, lda0, B, ldb, M, C, ldc, db, a_amax, b_amax);
  if (mode == 3) return tcg_amax(st, A0, M, c_amax);
  set_error("lnrf_tcgemm: unknown mode %d", mode);
  return LNRF_E_INVALID;
}
