// K6 dX chain, pair kernel: the input-gradient chain of the NeRF MLP backward (what jax.grad
// derives from model.py:42-62 at train.py:90) with the forward pair kernel's structure
// (mlp_tc_pair.cuh): one CTA per SM owns TWO 128-sample tiles, every streamed transposed-weight
// chunk ([128 n x 64 k] bf16, 4-slot ring) feeds one MMA group per tile, the two N halves of a
// layer's accumulator complete at different times so that the epilogue of one half overlaps the
// MMAs of the other.
//
// Per tile: dc = (dpre @ W11^T) * (c > 0) is formed on the CUDA cores (W11 / w9 are constant-bank
// operands) -> B0: g8 = dc @ W10[:256]^T + spre (x) w9 -> B1..B8: g_{l-1} = (g_l @ W_l[:256]^T)
// * (h_{l-1} > 0).  Every g tile leaves as two bulk stores (blocks 0,1 / blocks 2,3) of the exact
// shared-memory image, which the dW kernel reads back as an MN-major UMMA operand.
//
// The single-tile kernel it replaces (mlp_tc_bwd.cu, kept for A/B) had ONE 32 KB weight slot per
// CTA, so every chunk paid a full L2 round trip before its four MMAs: 1.48 ms on the fine level.
#include "mlp_tc_pair.cuh"

namespace lnrf {

using namespace ptx;

// head weights of the model being differentiated (copied before every launch, stream-ordered)
struct BwdSmall {
  float w9[256];      // Dense_9 kernel [256,1]
  float w11[128 * 3]; // Dense_11 kernel [128,3]
};
static __constant__ BwdSmall c_bsmall;

// 32 accumulator columns -> four 16-byte row chunks of the g tile.
// FIRST: g8 = acc + spre * w9 (no mask, model.py:57); else g = acc where the forward activation
// was positive (mask bit 31-j of `mwd` <-> column C0+j).
template <int C0, bool FIRST>
__device__ __forceinline__ void bwd_store32(const uint32_t (&v)[32], uint32_t sA, int r, uint32_t mwd,
                                            float spre) {
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float f0, f1;
    if (FIRST) {
      f0 = fmaf(spre, c_bsmall.w9[C0 + j], __uint_as_float(v[j]));
      f1 = fmaf(spre, c_bsmall.w9[C0 + j + 1], __uint_as_float(v[j + 1]));
    } else {
      f0 = (mwd & (0x80000000u >> j)) ? __uint_as_float(v[j]) : 0.0f;
      f1 = (mwd & (0x80000000u >> (j + 1))) ? __uint_as_float(v[j + 1]) : 0.0f;
    }
    pk[j / 2] = pack_bf16x2(f0, f1);
  }
  const uint32_t blk = sA + (C0 >> 6) * kABlockBytes;
  constexpr int cbase = (C0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    store_row_chunk(blk, r, cbase + q, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
}

// One N half (128 accumulator columns); for half 1 "accumulator drained" is signalled as soon as
// the last TMEM load has landed (bar_drained = 0: nobody waits for it).
template <int H, bool FIRST>
__device__ __forceinline__ void bwd_epi_half(uint32_t tm_lane, uint32_t sA, int r, uint32_t bar_drained,
                                             const uint32_t (&mw)[4], float spre) {
  constexpr int B = H * 128;
  uint32_t va[32], vb[32];
  tmem_ld32(tm_lane + B, va);
  tmem_wait_ld_dep(va);
  tmem_ld32(tm_lane + B + 32, vb);
  bwd_store32<B, FIRST>(va, sA, r, mw[0], spre);
  tmem_wait_ld_dep(vb);
  tmem_ld32(tm_lane + B + 64, va);
  bwd_store32<B + 32, FIRST>(vb, sA, r, mw[1], spre);
  tmem_wait_ld_dep(va);
  tmem_ld32(tm_lane + B + 96, vb);
  bwd_store32<B + 64, FIRST>(va, sA, r, mw[2], spre);
  tmem_wait_ld_dep(vb);
  if (H == 1 && bar_drained) {
    tc_fence_before();
    mbar_arrive(bar_drained);
  }
  bwd_store32<B + 96, FIRST>(vb, sA, r, mw[3], spre);
}

// One tensor layer's epilogue for one tile.  `last`: g0 feeds no further GEMM, so nothing is
// signalled to the MMA issuer (the next pair's prologue re-arms the three barriers).
template <bool FIRST>
__device__ __forceinline__ void bwd_epi_layer(const TcBwdArgs& args, int X, int r, bool leader, bool tile_ok,
                                              int64_t tile, int out_layer, bool last, uint32_t par,
                                              uint32_t tm_lane, uint32_t sA, uint32_t bars,
                                              const uint4* mask_row, float spre) {
  // the layer's ReLU masks: issued before the accumulator wait so that the latency is hidden
  uint4 ma = make_uint4(0u, 0u, 0u, 0u), mb = ma;
  if (!FIRST && tile_ok) {
    ma = __ldg(mask_row + out_layer * 256);
    mb = __ldg(mask_row + out_layer * 256 + 1);
  }
  uint8_t* gdst = args.stash.G[out_layer] + tile * kTileBytes;
  // ---- half 0: columns 0..127 -> blocks 0,1
  mbar_wait(bars + PairSmem::acc0 + 8 * X, par);
  // B0 has K = 128: BOTH N halves read blocks 0,1 (dc), and its half-0 commit precedes the half-1
  // MMAs, so blocks 0,1 may only be overwritten once half 1 is complete as well.
  if (FIRST) mbar_wait(bars + PairSmem::acc1 + 8 * X, par);
  tc_fence_after();
  if (leader) {
    // B0 overwrites blocks 0,1 while the dc image (the most recent bulk group) may still be read
    if (FIRST) bulk_wait_read0();
    else bulk_wait_read1();
  }
  pair_bar(X);
  {
    const uint32_t mw[4] = {ma.x, ma.y, ma.z, ma.w};
    bwd_epi_half<0, FIRST>(tm_lane, sA, r, 0u, mw, spre);
  }
  fence_proxy_async_smem();
  pair_bar(X);
  if (leader) {
    if (tile_ok) bulk_s2g(gdst, sA, 2 * kABlockBytes);
    bulk_commit();
  }
  tc_fence_before();
  if (!last) mbar_arrive(bars + PairSmem::a_ready0 + 8 * X);
  // ---- half 1: columns 128..255 -> blocks 2,3
  mbar_wait(bars + PairSmem::acc1 + 8 * X, par);
  tc_fence_after();
  if (leader) bulk_wait_read1();
  pair_bar(X);
  {
    const uint32_t mw[4] = {mb.x, mb.y, mb.z, mb.w};
    bwd_epi_half<1, FIRST>(tm_lane, sA, r, last ? 0u : bars + PairSmem::drained1 + 8 * X, mw, spre);
  }
  fence_proxy_async_smem();
  pair_bar(X);
  if (leader) {
    if (tile_ok) bulk_s2g(gdst + 2 * kABlockBytes, sA + 2 * kABlockBytes, 2 * kABlockBytes);
    bulk_commit();
  }
  tc_fence_before();
  if (!last) mbar_arrive(bars + PairSmem::a_ready1 + 8 * X);
}

__global__ void __launch_bounds__(kPairThreads, 1)
nerf_bwd_dx_pair_kernel(const __grid_constant__ TcBwdArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();  // SW128 operands need 1024-byte aligned blocks
  const uint32_t sA0 = smem_base + PairSmem::a_off;
  const uint32_t sW = smem_base + PairSmem::w_off;
  const uint32_t bars = smem_base + PairSmem::bar_off;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + PairSmem::bar_off + PairSmem::tmem_slot);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (args.m + 127) / 128;
  const int64_t pairs = (tiles + 1) / 2;
  const int64_t my_pairs = (pairs > blockIdx.x) ? (pairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < PairCfg<false>::stages; ++s) {
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
    if (lane == 0) pair_producer<false>(args.packed, c_chunks.b2, kB2Chunks, my_pairs, sW, bars);
  } else if (warp == 9) {
    pair_mma<false>(c_pair_meta.b2, kB2Chunks, my_pairs, sA0, sW, bars, tmem);
  } else {
    // ===== epilogue group X: thread r owns row r of tile X
    const int X = warp >> 2;
    const int r = tid & 127;
    const bool leader = r == 0;
    const uint32_t sA = sA0 + X * kPairTileBytes;
    const uint32_t tm_lane = tmem + (uint32_t((warp & 3) * 32) << 16) + X * 256;
    float acc_db9 = 0.f, acc_db11[3] = {0.f, 0.f, 0.f};
    uint32_t par = 0;  // nine layers per pair: the barrier phase parity keeps alternating
    for (int64_t t = 0; t < my_pairs; ++t) {
      const int64_t tile = 2 * (blockIdx.x + t * gridDim.x) + X;
      const bool tile_ok = tile < tiles;  // the last pair may have no tile B
      const int64_t s = tile * 128 + r;
      const bool valid = tile_ok && s < args.m;
      // row-major ReLU masks written by the forward: [tile][layer 9][row 128][8 words]
      const uint4* mask_row = reinterpret_cast<const uint4*>(args.stash.MASK + ((tile * 9) * 128 + r) * 8);
      // ---- head gradients (model.py:57,60): softplus' = sigmoid(pre) = 1 - exp(-density)
      float spre = 0.f, dp[3] = {0.f, 0.f, 0.f};
      uint4 mc4 = make_uint4(0u, 0u, 0u, 0u);
      if (valid) {
        spre = __ldg(args.d_dens + s) * (-expm1f(-__ldg(args.dens + s)));
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float y = __ldg(args.rgb + s * 3 + j);
          dp[j] = __ldg(args.d_rgb + s * 3 + j) * (1.0f - y * y);
        }
      }
      if (tile_ok) {
        mc4 = __ldg(mask_row + 8 * 256);
        args.stash.SPRE[s] = spre;
        reinterpret_cast<float4*>(args.stash.DPRE)[s] = make_float4(dp[0], dp[1], dp[2], 0.f);
      }
      acc_db9 += spre;
      acc_db11[0] += dp[0]; acc_db11[1] += dp[1]; acc_db11[2] += dp[2];
      // ---- dc = (dpre @ W11^T) * (c > 0) -> blocks 0,1 (and the DC stash image)
      if (leader) bulk_wait_read0();  // the previous tile's g0 image has left smem
      pair_bar(X);
      {
        const uint32_t mc[4] = {mc4.x, mc4.y, mc4.z, mc4.w};
#pragma unroll
        for (int c0 = 0; c0 < kHC; c0 += 32) {
          uint32_t pk[16];
          const uint32_t mwd = mc[c0 >> 5];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float v0 = dp[0] * c_bsmall.w11[(c0 + j) * 3 + 0] + dp[1] * c_bsmall.w11[(c0 + j) * 3 + 1] +
                       dp[2] * c_bsmall.w11[(c0 + j) * 3 + 2];
            float v1 = dp[0] * c_bsmall.w11[(c0 + j) * 3 + 3] + dp[1] * c_bsmall.w11[(c0 + j) * 3 + 4] +
                       dp[2] * c_bsmall.w11[(c0 + j) * 3 + 5];
            v0 = (mwd & (0x80000000u >> j)) ? v0 : 0.0f;
            v1 = (mwd & (0x80000000u >> (j + 1))) ? v1 : 0.0f;
            pk[j / 2] = pack_bf16x2(v0, v1);
          }
          const uint32_t blk = sA + (c0 >> 6) * kABlockBytes;
          const int cbase = (c0 & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            store_row_chunk(blk, r, cbase + q, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      pair_bar(X);
      if (leader) {
        if (tile_ok) bulk_s2g(args.stash.DC + tile * 2 * kABlockBytes, sA, 2 * kABlockBytes);
        bulk_commit();
      }
      // B0 may start for this tile: dc written, both accumulator halves drained (the previous
      // pair's last epilogue finished in program order)
      tc_fence_before();
      mbar_arrive(bars + PairSmem::a_ready0 + 8 * X);
      mbar_arrive(bars + PairSmem::a_ready1 + 8 * X);
      mbar_arrive(bars + PairSmem::drained1 + 8 * X);
      // ---- B0: g8 = acc + spre * w9;  B1..B8: g_{l-1} = acc * (h_{l-1} > 0)
      bwd_epi_layer<true>(args, X, r, leader, tile_ok, tile, 8, false, par, tm_lane, sA, bars, mask_row, spre);
      par ^= 1;
#pragma unroll 1
      for (int tl = 1; tl < kBwLayers; ++tl) {
        bwd_epi_layer<false>(args, X, r, leader, tile_ok, tile, 8 - tl, tl == kBwLayers - 1, par, tm_lane, sA,
                             bars, mask_row, 0.0f);
        par ^= 1;
      }
      tc_fence_before();  // orders these TMEM reads before the next pair's arrivals
    }
    if (leader) bulk_wait0();  // all stash stores complete before the CTA exits
    // bias gradients of the two heads: db9 = sum spre, db11 = sum dpre
    acc_db9 = warp_sum(acc_db9);
#pragma unroll
    for (int j = 0; j < 3; ++j) acc_db11[j] = warp_sum(acc_db11[j]);
    if (lane == 0) {
      atomicAdd(args.G + c_nerf.b[9], acc_db9);
#pragma unroll
      for (int j = 0; j < 3; ++j) atomicAdd(args.G + c_nerf.b[11] + j, acc_db11[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

int init_mlp_tc_bwd2() {
  int rc = upload_tc_tables();
  if (rc) return rc;
  if ((rc = upload_pair_meta())) return rc;
  LNRF_CUDA(cudaFuncSetAttribute(nerf_bwd_dx_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)PairSmem::total));
  return LNRF_OK;
}

int nerf_bwd_dx_pair(const TcBwdArgs& a, cudaStream_t st) {
  // head kernels -> constant bank (2.5 KB device-to-device, stream-ordered)
  LNRF_CUDA(cudaMemcpyToSymbolAsync(c_bsmall, a.P + kNerf.w[9], 256 * sizeof(float), offsetof(BwdSmall, w9),
                                    cudaMemcpyDeviceToDevice, st));
  LNRF_CUDA(cudaMemcpyToSymbolAsync(c_bsmall, a.P + kNerf.w[11], 384 * sizeof(float), offsetof(BwdSmall, w11),
                                    cudaMemcpyDeviceToDevice, st));
  const int64_t pairs = (ceil_div(a.m, 128) + 1) / 2;
  int64_t grid = sm_count();
  if (grid > pairs) grid = pairs;
  nerf_bwd_dx_pair_kernel<<<(unsigned)grid, kPairThreads, PairSmem::total, st>>>(a);
  LNRF_LAUNCH_CHECK("nerf_bwd_dx_pair_kernel");
  return LNRF_OK;
}

}  // namespace lnrf
