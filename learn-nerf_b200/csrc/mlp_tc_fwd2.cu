// K2 on the tensor cores, pair kernel: the whole NeRFModel forward (model.py:42-62) fused
// into one tcgen05/TMEM kernel, two 128-sample tiles per CTA (see mlp_tc_pair.cuh).
//
// Per tile the 128 x 256 bf16 activation matrix lives in shared memory as the UMMA A operand
// (four K-major SW128 blocks + a fifth block holding the positional encoding, which is
// computed in registers and never touches HBM).  The epilogue warps read the fp32 accumulator
// back with tcgen05.ld (software-pipelined, 32 columns per load), add the bias, apply ReLU,
// convert to bf16 and write the next layer's A operand in place.  The density head rides as
// output column 128 of the colour-layer GEMM; the 128 -> 3 rgb head runs in fp32 FMAs.
// With SAVE the activation tile images and row-major 1-bit ReLU masks are streamed to the
// stash with bulk stores for the backward kernels.
#include <stdlib.h>

#include "mlp_tc_pair.cuh"

namespace lnrf {

using namespace ptx;

// biases and the rgb head of the model being evaluated (copied before every launch, stream-ordered)
static __constant__ SmallParams c_small;

struct TcFwdArgs2 {
  const uint8_t* packed;
  const float* x;
  const float* d;
  const float* rays;
  const float* ts;
  int T;
  int64_t m;
  float* dens;
  float* rgb;
  TcStash stash;
  int debug;  // profiling ablations: 8 = skip the stash stores, 16 = alias all stash tiles onto the first 64
};

// sin/cos of a = x * 2^f with an explicit two-step Cody-Waite reduction to [-pi, pi] followed by
// the MUFU approximations (abs error ~5e-7, far below the bf16 rounding of the result).
__device__ __forceinline__ void fast_sincos(float a, float* s, float* c) {
  const float k = rintf(a * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, a);
  r = fmaf(k, 1.7484555e-7f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}

// bias + (ReLU) + bf16 pack of 32 accumulator columns -> four 16-byte row chunks of the A tile.
// C0 is compile-time and TL warp-uniform, so every bias is a constant-bank operand reached through a
// uniform register; the layers T0..T7 share ONE copy of this code (a runtime loop): the fully
// unrolled ten-layer epilogue was ~120 KB of SASS and lost 15 % of its issue slots to instruction
// fetch (ncu no_inst).
// SAVE: also shifts the 32 "pre-activation > 0" bits into `mword` (column c0+j -> bit 31-j).
template <bool RELU, int C0, bool SAVE>  // RELU = TL < 8: Dense_8's output feeds the heads raw (model.py:53-58)
__device__ __forceinline__ void epi_store32(const uint32_t (&v)[32], uint32_t sA, int r, uint32_t& mword, int TL) {
  uint32_t pk[16];
  uint32_t signs = 0;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const float f0 = __uint_as_float(v[j]) + c_small.b[TL][C0 + j];
    const float f1 = __uint_as_float(v[j + 1]) + c_small.b[TL][C0 + j + 1];
    pk[j / 2] = RELU ? pack_bf16x2_relu(f0, f1) : pack_bf16x2(f0, f1);
    if (SAVE && RELU) {  // collect sign bits: (signs << 1) | sign(f)
      signs = __funnelshift_l(__float_as_uint(f0), signs, 1);
      signs = __funnelshift_l(__float_as_uint(f1), signs, 1);
    }
  }
  mword = ~signs;
  const uint32_t blk = sA + (C0 >> 6) * kABlockBytes;
  constexpr int cbase = (C0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    store_row_chunk(blk, r, cbase + q, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
}

// One N half (128 accumulator columns) of a hidden layer's epilogue.  For half 1 the
// "accumulator drained" barrier is signalled as soon as the last TMEM load has landed.
template <bool RELU, int H, bool SAVE>
__device__ __forceinline__ void epi_half(uint32_t tm_lane, uint32_t sA, int r, uint32_t bar_drained,
                                         uint32_t (&mw)[4], int TL) {
  constexpr int B = H * 128;
  uint32_t va[32], vb[32];
  tmem_ld32(tm_lane + B, va);
  tmem_wait_ld_dep(va);
  tmem_ld32(tm_lane + B + 32, vb);
  epi_store32<RELU, B, SAVE>(va, sA, r, mw[0], TL);
  tmem_wait_ld_dep(vb);
  tmem_ld32(tm_lane + B + 64, va);
  epi_store32<RELU, B + 32, SAVE>(vb, sA, r, mw[1], TL);
  tmem_wait_ld_dep(va);
  tmem_ld32(tm_lane + B + 96, vb);
  epi_store32<RELU, B + 64, SAVE>(va, sA, r, mw[2], TL);
  tmem_wait_ld_dep(vb);
  if (H == 1) {
    tc_fence_before();
    mbar_arrive(bar_drained);
  }
  epi_store32<RELU, B + 96, SAVE>(vb, sA, r, mw[3], TL);
}

template <bool LAST, bool SAVE>  // LAST: TL == 8 (no ReLU, the d_emb block is written); else TL = 0..7
__device__ __forceinline__ void epi_layer(const TcFwdArgs2& args, int TL, int X, int r, bool leader, bool tile_ok,
                                          int64_t tile, uint32_t tm_lane, uint32_t sA, uint32_t bars,
                                          uint4* mask_row, const uint32_t (&de)[12]) {
  const uint32_t par = TL & 1;  // ten layers per tile pair: the phase parity of layer TL is fixed
  uint32_t mw[4];
  if (args.debug & 8) tile_ok = false;
  if (args.debug & 16) tile &= 63;
  // ---- half 0: columns 0..127 -> A blocks 0,1
  mbar_wait(bars + PairSmem::acc0 + 8 * X, par);
  tc_fence_after();
  // The stash stores are issued per half (blocks 0,1 / blocks 2,3) as separate bulk groups, so
  // before overwriting a half only the groups older than the most recent one must have left smem.
  if (SAVE) {
    if (leader) bulk_wait_read1();
    pair_bar(X);
  }
  epi_half<!LAST, 0, SAVE>(tm_lane, sA, r, 0u, mw, TL);
  if (SAVE && !LAST && mask_row) mask_row[TL * 256] = make_uint4(mw[0], mw[1], mw[2], mw[3]);
  fence_proxy_async_smem();
  if (SAVE) {
    pair_bar(X);
    if (leader && tile_ok) bulk_s2g(args.stash.H[TL] + tile * kTileBytes, sA, 2 * kABlockBytes);
    if (leader) bulk_commit();
  }
  tc_fence_before();
  mbar_arrive(bars + PairSmem::a_ready0 + 8 * X);
  // ---- half 1: columns 128..255 -> A blocks 2,3
  mbar_wait(bars + PairSmem::acc1 + 8 * X, par);
  tc_fence_after();
  if (SAVE) {
    if (leader) bulk_wait_read1();
    pair_bar(X);
  }
  epi_half<!LAST, 1, SAVE>(tm_lane, sA, r, bars + PairSmem::drained1 + 8 * X, mw, TL);
  if (SAVE && !LAST && mask_row) mask_row[TL * 256 + 1] = make_uint4(mw[0], mw[1], mw[2], mw[3]);
  if (LAST) {  // x_emb is dead after T5: block 4 now carries d_emb (24 cols) + zeros
    const uint32_t blk = sA + 4 * kABlockBytes;
    store_row_chunk(blk, r, 0, de[0], de[1], de[2], de[3]);
    store_row_chunk(blk, r, 1, de[4], de[5], de[6], de[7]);
    store_row_chunk(blk, r, 2, de[8], de[9], de[10], de[11]);
#pragma unroll
    for (int c = 3; c < 8; ++c) store_row_chunk(blk, r, c, 0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  if (SAVE) {
    pair_bar(X);
    if (leader && tile_ok) {
      bulk_s2g(args.stash.H[TL] + tile * kTileBytes + 2 * kABlockBytes, sA + 2 * kABlockBytes, 2 * kABlockBytes);
      if (LAST) bulk_s2g(args.stash.DE + tile * kABlockBytes, sA + 4 * kABlockBytes, kABlockBytes);
    }
    if (leader) bulk_commit();
  }
  tc_fence_before();
  mbar_arrive(bars + PairSmem::a_ready1 + 8 * X);
}

// FULLN = lockstep schedule on full-N weight chunks (faster without the stash), else the N-half
// pipelined schedule.
template <bool SAVE, bool FULLN>
__global__ void __launch_bounds__(kPairThreads, 1)
nerf_fwd_pair_kernel(const __grid_constant__ TcFwdArgs2 args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) {  // SW128 operands need 1024-byte aligned blocks
    if (threadIdx.x == 0) printf("lnrf: dynamic smem base 0x%x not 1024-aligned\n", smem_base);
    __trap();
  }
  const uint32_t sA0 = smem_base + PairSmem::a_off;
  const uint32_t sW = smem_base + PairSmem::w_off;
  const uint32_t bars = smem_base + PairSmem::bar_off;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + PairSmem::bar_off + PairSmem::tmem_slot);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (args.m + 127) / 128;
  const int64_t pairs = (tiles + 1) / 2;
  const int64_t my_pairs = (pairs > blockIdx.x) ? (pairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < PairCfg<FULLN>::stages; ++s) {
      mbar_init(bars + PairSmem::full + 8 * s, 1);
      mbar_init(bars + PairSmem::empty + 8 * s, 1);
    }
    for (int X = 0; X < 2; ++X) {
      mbar_init(bars + PairSmem::a_ready0 + 8 * X, 128);
      mbar_init(bars + PairSmem::a_ready1 + 8 * X, 128);
      mbar_init(bars + PairSmem::drained1 + 8 * X, 128);
      mbar_init(bars + PairSmem::acc0 + 8 * X, 1);
      mbar_init(bars + PairSmem::acc1 + 8 * X, 1);
    }
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(bars + PairSmem::tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 8) {
    if (lane == 0) {
      if (FULLN) pair_producer<true>(args.packed, c_chunks.f, kTcChunks, my_pairs, sW, bars);
      else pair_producer<false>(args.packed, c_chunks.f2, kF2Chunks, my_pairs, sW, bars);
    }
  } else if (warp == 9) {
    if (FULLN) pair_mma<true>(c_pair_meta.f_full, kTcChunks, my_pairs, sA0, sW, bars, tmem);
    else pair_mma<false>(c_pair_meta.f2, kF2Chunks, my_pairs, sA0, sW, bars, tmem);
  } else {
    // ===== epilogue group X: thread r owns row r of tile X
    const int X = warp >> 2;
    const int r = tid & 127;
    const bool leader = r == 0;
    const uint32_t sA = sA0 + X * kPairTileBytes;
    const uint32_t tm_lane = tmem + (uint32_t((warp & 3) * 32) << 16) + X * 256;
    for (int64_t t = 0; t < my_pairs; ++t) {
      const int64_t tile = 2 * (blockIdx.x + t * gridDim.x) + X;
      const bool tile_ok = tile < tiles;  // the last pair may have no tile B
      const int64_t s = tile * 128 + r;
      const bool valid = tile_ok && s < args.m;
      uint4* mask_row = (SAVE && tile_ok) ? reinterpret_cast<uint4*>(args.stash.MASK + ((tile * 9) * 128 + r) * 8)
                                          : nullptr;
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
        if (leader) bulk_wait_read0();
        pair_bar(X);
      }
      // ---- sinusoidal_emb(x, 10) -> A block 4 (cols dim*20 + [sin f | cos f]), cols 60..63 = 0
      {
        uint32_t pk[32];
#pragma unroll
        for (int dim = 0; dim < 3; ++dim) {
          float sn[kXFreqs], cs[kXFreqs];
#pragma unroll
          for (int f = 0; f < kXFreqs; ++f) fast_sincos(px[dim] * float(1 << f), &sn[f], &cs[f]);
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
        for (int f = 0; f < kDFreqs; ++f) fast_sincos(dv[dim] * float(1 << f), &sn[f], &cs[f]);
        de[dim * 4 + 0] = pack_bf16x2(sn[0], sn[1]);
        de[dim * 4 + 1] = pack_bf16x2(sn[2], sn[3]);
        de[dim * 4 + 2] = pack_bf16x2(cs[0], cs[1]);
        de[dim * 4 + 3] = pack_bf16x2(cs[2], cs[3]);
      }
      fence_proxy_async_smem();
      if (SAVE) {
        pair_bar(X);
        if (leader && tile_ok) {
          bulk_s2g(args.stash.XE + tile * kABlockBytes, sA + 4 * kABlockBytes, kABlockBytes);
          bulk_commit();
        }
      }
      // T0 may start for this tile: embedding written, both accumulator halves drained (the
      // previous pair's head epilogue finished in program order)
      tc_fence_before();
      mbar_arrive(bars + PairSmem::a_ready0 + 8 * X);
      mbar_arrive(bars + PairSmem::a_ready1 + 8 * X);
      mbar_arrive(bars + PairSmem::drained1 + 8 * X);
      // ---- hidden layers T0..T8
      if (SAVE) {
        // with the stash the eight ReLU layers share one copy of the epilogue code (runtime loop):
        // 1.25 -> 1.19 ms on the fine level (fewer instruction-cache misses)
#pragma unroll 1
        for (int TL = 0; TL < 8; ++TL)
          epi_layer<false, SAVE>(args, TL, X, r, leader, tile_ok, tile, tm_lane, sA, bars, mask_row, de);
      } else {
        // rendering: fully unrolled, every bias an immediate constant-bank operand (the loop form
        // costs 0.85 -> 1.09 ms here: the uniform-register bias loads sit on the critical path)
#pragma unroll
        for (int TL = 0; TL < 8; ++TL)
          epi_layer<false, SAVE>(args, TL, X, r, leader, tile_ok, tile, tm_lane, sA, bars, mask_row, de);
      }
      epi_layer<true, SAVE>(args, 8, X, r, leader, tile_ok, tile, tm_lane, sA, bars, mask_row, de);
      // ---- T9: colour layer (half 0) + density column (half 1) and the fp32 rgb head
      mbar_wait(bars + PairSmem::acc0 + 8 * X, 1u);  // layer 9: parity 1
      tc_fence_after();
      if (SAVE) {  // blocks 0,1 get the colour-hidden image once the z8 image has left smem
        if (leader) bulk_wait_read0();
        pair_bar(X);
      }
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      uint32_t mwc[4];
#pragma unroll
      for (int c0 = 0; c0 < kHC; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tm_lane + c0, v);
        tmem_wait_ld_dep(v);
        uint32_t pk[16];
        uint32_t signs = 0;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float p0 = __uint_as_float(v[j]) + c_small.b10[c0 + j];
          const float p1 = __uint_as_float(v[j + 1]) + c_small.b10[c0 + j + 1];
          const float h0 = fmaxf(p0, 0.0f), h1 = fmaxf(p1, 0.0f);  // model.py:59
          o0 = fmaf(h0, c_small.w11[(c0 + j) * 3 + 0], o0);
          o1 = fmaf(h0, c_small.w11[(c0 + j) * 3 + 1], o1);
          o2 = fmaf(h0, c_small.w11[(c0 + j) * 3 + 2], o2);
          o0 = fmaf(h1, c_small.w11[(c0 + j) * 3 + 3], o0);
          o1 = fmaf(h1, c_small.w11[(c0 + j) * 3 + 4], o1);
          o2 = fmaf(h1, c_small.w11[(c0 + j) * 3 + 5], o2);
          if (SAVE) {
            pk[j / 2] = pack_bf16x2(h0, h1);
            signs = __funnelshift_l(__float_as_uint(p0), signs, 1);
            signs = __funnelshift_l(__float_as_uint(p1), signs, 1);
          }
        }
        if (SAVE) {
          mwc[c0 >> 5] = ~signs;
          const uint32_t blk = sA + (c0 >> 6) * kABlockBytes;
          const int cbase = (c0 & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            store_row_chunk(blk, r, cbase + q, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
      }
      if (SAVE && mask_row) mask_row[8 * 256] = make_uint4(mwc[0], mwc[1], mwc[2], mwc[3]);
      mbar_wait(bars + PairSmem::acc1 + 8 * X, 1u);
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld32(tm_lane + kHC, v);  // column 128 = Dense_9 pre-activation
        tmem_wait_ld_dep(v);
        if (valid) {
          args.dens[s] = softplus_f(__uint_as_float(v[0]) + c_small.b9);  // model.py:57
          args.rgb[s * 3 + 0] = tanhf(o0 + c_small.b11[0]);  // model.py:60
          args.rgb[s * 3 + 1] = tanhf(o1 + c_small.b11[1]);
          args.rgb[s * 3 + 2] = tanhf(o2 + c_small.b11[2]);
        }
      }
      if (SAVE) {
        fence_proxy_async_smem();
        pair_bar(X);
        if (leader && tile_ok) {
          bulk_s2g(args.stash.C + tile * 2 * kABlockBytes, sA, 2 * kABlockBytes);
          bulk_commit();
        }
      }
      tc_fence_before();  // orders these TMEM reads before the next pair's arrivals
    }
    if (SAVE && leader) bulk_wait0();  // all stash stores complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

// profiling ablations (see TcFwdArgs2::debug): read once from LNRF_DEBUG_FLAGS in lnrf_init
static int g_fwd_debug = 0;

int init_mlp_tc_fwd2() {
  if (const char* e = getenv("LNRF_DEBUG_FLAGS")) g_fwd_debug = atoi(e) & (8 | 16 | 64);
  int rc = upload_tc_tables();
  if (rc) return rc;
  if ((rc = upload_pair_meta())) return rc;
  LNRF_CUDA(cudaFuncSetAttribute(nerf_fwd_pair_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)PairSmem::total));
  LNRF_CUDA(cudaFuncSetAttribute(nerf_fwd_pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)PairSmem::total));
  LNRF_CUDA(cudaFuncSetAttribute(nerf_fwd_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)PairSmem::total));
  return LNRF_OK;
}

int nerf_fwd_pair(const void* packed, const float* x, const float* d, const float* rays, const float* ts,
                  int64_t m, int T, bool save, const TcStash& stash, float* dens, float* rgb,
                  cudaStream_t st) {
  // this model's biases / rgb head -> constant bank (11 KB device-to-device, stream-ordered)
  LNRF_CUDA(cudaMemcpyToSymbolAsync(c_small, reinterpret_cast<const uint8_t*>(packed) + kSmallOffset,
                                    sizeof(SmallParams), 0, cudaMemcpyDeviceToDevice, st));
  TcFwdArgs2 a{reinterpret_cast<const uint8_t*>(packed), x, d, rays, ts, T, m, dens, rgb, stash, g_fwd_debug};
  const int64_t pairs = (ceil_div(m, 128) + 1) / 2;
  int64_t grid = sm_count();
  if (grid > pairs) grid = pairs;
  if (save) nerf_fwd_pair_kernel<true, false><<<(unsigned)grid, kPairThreads, PairSmem::total, st>>>(a);
  else if (g_fwd_debug & 64) nerf_fwd_pair_kernel<false, false><<<(unsigned)grid, kPairThreads, PairSmem::total, st>>>(a);
  else nerf_fwd_pair_kernel<false, true><<<(unsigned)grid, kPairThreads, PairSmem::total, st>>>(a);
  LNRF_LAUNCH_CHECK("nerf_fwd_pair_kernel");
  return LNRF_OK;
}

}  // namespace lnrf
