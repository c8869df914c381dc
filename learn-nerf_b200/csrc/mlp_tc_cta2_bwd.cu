// K6 dX chain, CTA-pair kernel: the input-gradient chain of the NeRF MLP backward (what jax.grad
// derives from model.py:42-62 at train.py:90) as cta_group::2 MMAs over a cluster of two CTAs
// (structure and cross-CTA protocol: mlp_tc_cta2.cuh).
//
// Per tile: dc = (dpre @ W11^T) * (c > 0) is formed on the CUDA cores (W11 / w9 are constant-bank
// operands) -> B0: g8 = dc @ W10[:256]^T + spre (x) w9 -> B1..B8: g_{l-1} = (g_l @ W_l[:256]^T) *
// (h_{l-1} > 0).  Every g tile leaves as two bulk stores (blocks 0,1 by team 0 / blocks 2,3 by team 1)
// of the exact shared-memory image, which the dW kernel reads back as an MN-major UMMA operand.
#include <stdlib.h>


#include "mlp_tc_cta2.cuh"

namespace lnrf {

using namespace ptx;

struct C2BwdArgs {
  TcBwdArgs a;
  int off_b9, off_b11;  // float offsets of the two head biases inside the flat gradient
  int off_w9, off_w11;  // float offsets of the two head kernels inside the flat parameters
  C2Sched sched;
};

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// 32 accumulator columns [C0, C0 + 32) of the team's half -> four 16-byte row chunks of the g tile.
// FIRST: g8 = acc + spre * w9 (no mask, model.py:57), w9 staged in shared memory (`sw9`) by the team; else
// g = acc where the forward activation was positive (mask bit 31-j of `mwd` <-> column C0+j of the half).
template <int C0, bool FIRST>
__device__ __forceinline__ void c2_bwd_store32(const uint32_t (&v)[32], uint32_t blk0, int r, uint32_t mwd, float spre,
                                               uint32_t sw9) {
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float f0, f1, f2, f3;
    if (FIRST) {
      const float4 w = lds_f4(sw9 + (C0 + j) * 4);
      f0 = fmaf(spre, w.x, __uint_as_float(v[j]));
      f1 = fmaf(spre, w.y, __uint_as_float(v[j + 1]));
      f2 = fmaf(spre, w.z, __uint_as_float(v[j + 2]));
      f3 = fmaf(spre, w.w, __uint_as_float(v[j + 3]));
    } else {
      f0 = (mwd & (0x80000000u >> j)) ? __uint_as_float(v[j]) : 0.0f;
      f1 = (mwd & (0x80000000u >> (j + 1))) ? __uint_as_float(v[j + 1]) : 0.0f;
      f2 = (mwd & (0x80000000u >> (j + 2))) ? __uint_as_float(v[j + 2]) : 0.0f;
      f3 = (mwd & (0x80000000u >> (j + 3))) ? __uint_as_float(v[j + 3]) : 0.0f;
    }
    pk[j / 2] = pack_bf16x2(f0, f1);
    pk[j / 2 + 1] = pack_bf16x2(f2, f3);
  }
  const uint32_t blk = blk0 + (C0 >> 6) * kABlockBytes;
  constexpr int cbase = (C0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    store_row_chunk(blk, r, cbase + q, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
}

template <bool FIRST>
__device__ __forceinline__ void c2_bwd_epi_half(uint32_t tm, uint32_t blk0, int r, const uint4& m4, float spre,
                                                uint32_t sw9) {
  uint32_t va[32], vb[32];
  tmem_ld32(tm, va);
  tmem_wait_ld_dep(va);
  tmem_ld32(tm + 32, vb);
  c2_bwd_store32<0, FIRST>(va, blk0, r, m4.x, spre, sw9);
  tmem_wait_ld_dep(vb);
  tmem_ld32(tm + 64, va);
  c2_bwd_store32<32, FIRST>(vb, blk0, r, m4.y, spre, sw9);
  tmem_wait_ld_dep(va);
  tmem_ld32(tm + 96, vb);
  c2_bwd_store32<64, FIRST>(va, blk0, r, m4.z, spre, sw9);
  tmem_wait_ld_dep(vb);
  c2_bwd_store32<96, FIRST>(vb, blk0, r, m4.w, spre, sw9);
}

// One epilogue team: group g, column half H (compile-time: see c2_fwd_team).
template <int H>
__device__ __forceinline__ void c2_bwd_team(const C2BwdArgs& cargs, const C2Ctx& cx, int g, int64_t tiles) {
  const TcBwdArgs& args = cargs.a;
  constexpr int h = H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int team = 2 * g + H;
    // ===== epilogue team (g, h): thread r owns row r of group g's tile, columns 128 h .. 128 h + 127
    const int r = tid & 127;
    const bool leader = r == 0;
    const uint32_t sA = cx.sA0 + g * kPairTileBytes;
    const uint32_t blk0 = sA + 2 * h * kABlockBytes;
    const uint32_t tm = cx.tmem + (uint32_t((warp & 3) * 32) << 16) + g * 256 + h * 128;
    const uint32_t bar_a = cx.bars + C2Smem::a_ready + 8 * g, bar_acc = cx.bars + C2Smem::acc_full + 8 * g;
    const int64_t cid = cluster_id_x(), ncl = nclusters_x();
    const uint32_t sw9 = cx.bars + C2Smem::bias + uint32_t(g * 256 + H * 128) * 4u;
    const float4* w11v = reinterpret_cast<const float4*>(args.P + cargs.off_w11);  // Dense_11 kernel [128,3]
    const float w9_mine = __ldg(args.P + cargs.off_w9 + H * 128 + r);              // Dense_9 kernel [256,1]
    float acc_db9 = 0.f, acc_db11[3] = {0.f, 0.f, 0.f};
    uint32_t par = 0;  // nine layers per tile: the barrier phase parity keeps alternating
    for (int64_t t = 0; t < cx.my_iters; ++t) {
      const int64_t tile = ((cid + t * ncl) * 2 + g) * 2 + int64_t(cx.rank);
      const bool tile_ok = tile < tiles;
      const int64_t s = tile * 128 + r;
      const bool valid = tile_ok && s < args.m;
      // row-major ReLU masks written by the forward: [tile][layer 9][row 128][8 words]
      const uint4* mask_row = reinterpret_cast<const uint4*>(args.stash.MASK + ((tile * 9) * 128 + r) * 8);
      // ---- head gradients (model.py:57,60): softplus' = sigmoid(pre) = 1 - exp(-density)
      float spre = 0.f, dp[3] = {0.f, 0.f, 0.f};
      if (valid) {
        spre = __ldg(args.d_dens + s) * (-expm1f(-__ldg(args.dens + s)));
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float y = __ldg(args.rgb + s * 3 + j);
          dp[j] = __ldg(args.d_rgb + s * 3 + j) * (1.0f - y * y);
        }
      }
      if (h == 0) {
        // ---- dc = (dpre @ W11^T) * (c > 0) -> blocks 0,1 (and the DC stash image)
        uint4 mc4 = make_uint4(0u, 0u, 0u, 0u);
        if (tile_ok) mc4 = __ldg(mask_row + 8 * 256);
        if (leader) bulk_wait_read0();  // the previous tile's g0 image (blocks 0,1) has left smem
        team_bar(team);
        const uint32_t mc[4] = {mc4.x, mc4.y, mc4.z, mc4.w};
#pragma unroll
        for (int c0 = 0; c0 < kHC; c0 += 32) {
          uint32_t pk[16];
          const uint32_t mwd = mc[c0 >> 5];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {  // 4 columns = 12 head weights = three 16-byte read-only loads (L1 broadcast)
            const float4 wa = __ldg(w11v + (c0 + j) * 3 / 4), wb = __ldg(w11v + (c0 + j) * 3 / 4 + 1),
                         wc = __ldg(w11v + (c0 + j) * 3 / 4 + 2);
            float v0 = dp[0] * wa.x + dp[1] * wa.y + dp[2] * wa.z;
            float v1 = dp[0] * wa.w + dp[1] * wb.x + dp[2] * wb.y;
            float v2 = dp[0] * wb.z + dp[1] * wb.w + dp[2] * wc.x;
            float v3 = dp[0] * wc.y + dp[1] * wc.z + dp[2] * wc.w;
            v0 = (mwd & (0x80000000u >> j)) ? v0 : 0.0f;
            v1 = (mwd & (0x80000000u >> (j + 1))) ? v1 : 0.0f;
            v2 = (mwd & (0x80000000u >> (j + 2))) ? v2 : 0.0f;
            v3 = (mwd & (0x80000000u >> (j + 3))) ? v3 : 0.0f;
            pk[j / 2] = pack_bf16x2(v0, v1);
            pk[j / 2 + 1] = pack_bf16x2(v2, v3);
          }
          const uint32_t blk = sA + (c0 >> 6) * kABlockBytes;
          const int cbase = (c0 & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            store_row_chunk(blk, r, cbase + q, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
        fence_proxy_async_smem();
        team_bar(team);
        if (leader) {
          if (tile_ok) bulk_s2g(args.stash.DC + tile * 2 * kABlockBytes, sA, 2 * kABlockBytes);
          bulk_commit();
        }
      } else {
        // team 1: the per-row head gradients for the dW kernel and the two head-bias gradients
        if (tile_ok) {
          args.stash.SPRE[s] = spre;
          reinterpret_cast<float4*>(args.stash.DPRE)[s] = make_float4(dp[0], dp[1], dp[2], 0.f);
        }
        acc_db9 += spre;
        acc_db11[0] += dp[0]; acc_db11[1] += dp[1]; acc_db11[2] += dp[2];
      }
      // B0 may start for this tile: dc written, the whole accumulator drained (this thread's reads of the
      // previous tile's last layer are complete in program order)
      c2_arrive_a(bar_a, cx.rank);
      // ---- B0: g8 = acc + spre * w9;  B1..B8: g_{l-1} = acc * (h_{l-1} > 0)
#pragma unroll 1
      for (int tl = 0; tl < kBwLayers; ++tl) {
        const int out_layer = 8 - tl;
        uint4 m4 = make_uint4(0u, 0u, 0u, 0u);
        if (tl > 0 && tile_ok) m4 = __ldg(mask_row + out_layer * 256 + h);  // issued before the wait: latency hidden
        mbar_wait(bar_acc, par);
        par ^= 1;
        tc_fence_after();
        if (leader) bulk_wait_read0();  // this team's previous image (the same two blocks) has left smem
        if (tl == 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sw9 + r * 4), "f"(w9_mine) : "memory");
        team_bar(team);
        if (tl == 0) c2_bwd_epi_half<true>(tm, blk0, r, m4, spre, sw9);
        else c2_bwd_epi_half<false>(tm, blk0, r, m4, 0.0f, sw9);
        fence_proxy_async_smem();
        team_bar(team);
        if (leader) {
          if (tile_ok)
            bulk_s2g(args.stash.G[out_layer] + tile * kTileBytes + 2 * h * kABlockBytes, blk0, 2 * kABlockBytes);
          bulk_commit();
        }
        if (tl + 1 < kBwLayers) {  // g0 feeds no further GEMM: the next tile's prologue re-arms the barrier
          c2_arrive_a(bar_a, cx.rank);
        }
      }
    }
    if (leader) bulk_wait0();  // all stash stores complete before the CTA exits
    if (h == 1) {  // bias gradients of the two heads: db9 = sum spre, db11 = sum dpre
      acc_db9 = warp_sum(acc_db9);
#pragma unroll
      for (int j = 0; j < 3; ++j) acc_db11[j] = warp_sum(acc_db11[j]);
      if (lane == 0) {
        atomicAdd(args.G + cargs.off_b9, acc_db9);
#pragma unroll
        for (int j = 0; j < 3; ++j) atomicAdd(args.G + cargs.off_b11 + j, acc_db11[j]);
      }
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kC2Threads, 1)
nerf_bwd_dx_cta2_kernel(const __grid_constant__ C2BwdArgs cargs) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int64_t tiles = (cargs.a.m + 127) / 128;
  const int64_t quads = (tiles + 3) / 4;
  const C2Ctx cx = c2_setup(smem_raw, quads);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 16) {
    if (lane == 0) c2_producer(cargs.a.packed, cargs.sched, cx.my_iters, cx.rank, cx.sW, cx.bars);
  } else if (warp == 17) {
    if (cx.rank == 0) c2_mma(cargs.sched, cx.my_iters, cx.sA0, cx.sW, cx.bars, cx.tmem);
    else c2_relay(cargs.sched, cx.my_iters, cx.bars);
  } else if (warp & 4) {
    c2_bwd_team<1>(cargs, cx, warp >> 3, tiles);
  } else {
    c2_bwd_team<0>(cargs, cx, warp >> 3, tiles);
  }
  c2_teardown(cx);
}

// ---------------------------------------------------------------- host side
C2Sched c2_make_sched(const ChunkInfo* tab, int n, int layers);  // mlp_tc_cta2_fwd.cu
int c2_max_clusters();

static C2Sched g_bwd_sched;
int init_mlp_tc_cta2_bwd() {
  const ChunkTable t = build_chunk_table();
  g_bwd_sched = c2_make_sched(t.b, kBwChunks, kBwLayers);
  LNRF_CUDA(cudaFuncSetAttribute(nerf_bwd_dx_cta2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2Smem::total));
  return LNRF_OK;
}

int nerf_bwd_dx_cta2(const TcBwdArgs& a, cudaStream_t st) {
  C2BwdArgs ca{a, int(kNerf.b[9]), int(kNerf.b[11]), int(kNerf.w[9]), int(kNerf.w[11]), g_bwd_sched};
  const int64_t quads = (ceil_div(a.m, 128) + 3) / 4;
  int64_t clusters = c2_max_clusters();
  if (clusters > quads) clusters = quads;
  nerf_bwd_dx_cta2_kernel<<<unsigned(clusters * 2), kC2Threads, C2Smem::total, st>>>(ca);
  LNRF_LAUNCH_CHECK("nerf_bwd_dx_cta2_kernel");
  return LNRF_OK;
}

}  // namespace lnrf
