// K6 on the tensor cores: backward of the NeRF MLP (what jax.grad derives from
// model.py:42-62 at train.py:90) as two tcgen05/TMEM kernels over the forward's stash.
//
//  dX chain: mlp_tc_cta2_bwd.cu (per tile: head gradients, then g_{l-1} = (g_l @ W_l^T) * relu' for
//      l = 8..1 on the tensor cores, every g_l tile image streamed to the stash).
//  nerf_bwd_dw_kernel  "dW": dW_l = act_{l-1}^T @ g_l summed over all samples.  Both operands
//      are the stashed tile images read as MN-major UMMA operands (K = samples); each CTA
//      owns one (layer, tile-range) job, accumulates the full 256x256 fp32 dW in TMEM (all
//      512 columns) across its tiles and adds it to global memory once.  The otherwise idle
//      warps form the bias gradients (column sums of g) and the two tiny head gradients.
#include <stdlib.h>

#include "tc_common.cuh"

namespace lnrf {

using namespace ptx;

bool tc_ready();
int nerf_bwd_dx_cta2(const TcBwdArgs& a, cudaStream_t st);  // mlp_tc_cta2_bwd.cu
int64_t tc_workspace_bytes(int64_t m, bool save);

// ================================================================ dW
constexpr int kDwMaxJobs = 16;
constexpr int kDwStages = 3;
constexpr uint32_t kHalfBlock = 8192;              // 64 samples x 128 B
constexpr uint32_t kDwStageBytes = 8 * kHalfBlock;  // A: 4 half-blocks, B: 4 half-blocks
constexpr int kDwThreads = 192;                    // warp 0 producer, warp 1 MMA, warps 2-5 workers
constexpr uint32_t kDwSmemBytes = kDwStages * kDwStageBytes + 128 + 2 * 64 * 16;  // ring + barriers + row strip

struct DwJob {
  const uint8_t* A;   // stash image, a_blocks x 16 KB per tile
  const uint8_t* B;
  int a_blocks;       // 4, 2 or 1 (1: the second 64-feature block of the M=128 half is zero)
  int b_blocks;       // 4 (N = 256) or 2 (N = 128); 0 = no MMA (CUDA-core job only)
  int m_halves;       // M = 128 * m_halves
  int rows_valid;     // dW rows actually written
  int ld;             // leading dimension of dW (= out features)
  int extra;          // 0: none, 1: dW9 += z8^T spre (A = z8 image), 2: dW11 += c^T dpre (A = c image)
  float* dW;          // fp32 [rows, ld], accumulated
  float* db;          // column sums of the B image, accumulated (nullable)
  float* extra_out;
  int cta_begin, cta_count;
};
struct DwArgs {
  DwJob jobs[kDwMaxJobs];
  int n_jobs;
  int64_t tiles;
  const float* SPRE;
  const float* DPRE;
  int debug;  // ablation flags (tests/bench only): 1 = skip loads, 2 = skip MMAs, 4 = skip worker math
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(kDwThreads, 1)
nerf_bwd_dw_kernel(DwArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t bars = smem_base + kDwStages * kDwStageBytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kDwStages, bar_done = bars + 16 * kDwStages;
  const uint32_t tmem_slot = bar_done + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // which job / which tile range
  int ji = 0;
  while (ji + 1 < args.n_jobs && int(blockIdx.x) >= args.jobs[ji].cta_begin + args.jobs[ji].cta_count) ++ji;
  const DwJob job = args.jobs[ji];
  const int part = blockIdx.x - job.cta_begin;
  const int64_t t_begin = args.tiles * part / job.cta_count;
  const int64_t t_end = args.tiles * (part + 1) / job.cta_count;
  const int64_t n_iters = (t_end - t_begin) * 2;  // half tiles (64 samples) per ring slot
  const int N = job.b_blocks * 64;

  if (tid == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1 + 4);  // MMA commit + one arrive per worker warp
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (job.a_blocks == 1) {  // zero the second 64-feature half-block of every ring slot once
    for (int s = 0; s < kDwStages; ++s)
      for (uint32_t i = tid * 16; i < kHalfBlock; i += kDwThreads * 16)
        st_shared_v4(smem_base + s * kDwStageBytes + kHalfBlock + i, 0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {  // ===== producer: one half tile of A and B per ring slot
      uint32_t stage = 0, phase = 0;
      for (int64_t it = 0; it < n_iters; ++it) {
        const int64_t tile = t_begin + (it >> 1);
        const uint32_t half = uint32_t(it & 1) * kHalfBlock;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (args.debug & 1) {
          mbar_arrive(bar_full + 8 * stage);
          if (++stage == kDwStages) { stage = 0; phase ^= 1; }
          continue;
        }
        mbar_arrive_expect_tx(bar_full + 8 * stage, uint32_t(job.a_blocks + job.b_blocks) * kHalfBlock);
        const uint32_t sa = smem_base + stage * kDwStageBytes, sb = sa + 4 * kHalfBlock;
        // (an L2 evict_first hint on these read-once loads was measured: no gain, 6.10 vs 6.04 ms/step)
        for (int b = 0; b < job.a_blocks; ++b)
          bulk_g2s(sa + b * kHalfBlock, job.A + (tile * job.a_blocks + b) * int64_t(kABlockBytes) + half,
                   kHalfBlock, bar_full + 8 * stage);
        for (int b = 0; b < job.b_blocks; ++b)
          bulk_g2s(sb + b * kHalfBlock, job.B + (tile * job.b_blocks + b) * int64_t(kABlockBytes) + half,
                   kHalfBlock, bar_full + 8 * stage);
        if (++stage == kDwStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer: D[mh] += A_mh^T B, both operands MN-major, K = 64 samples
      uint32_t stage = 0, phase = 0;
      const uint32_t idesc = umma_idesc_bf16_mn(128, N > 0 ? N : 64);
      for (int64_t it = 0; it < n_iters; ++it) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (job.b_blocks > 0 && !(args.debug & 2)) {
          const uint32_t sa = smem_base + stage * kDwStageBytes, sb = sa + 4 * kHalfBlock;
          for (int mh = 0; mh < job.m_halves; ++mh) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem + mh * N, umma_desc_sw128_mnmajor(sa + mh * 2 * kHalfBlock + k * 2048, kHalfBlock),
                        umma_desc_sw128_mnmajor(sb + k * 2048, kHalfBlock), idesc, (it | k) ? 1u : 0u);
          }
        }
        umma_commit(bar_empty + 8 * stage);
        if (++stage == kDwStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(bar_done);
    }
  } else {
    // ===== workers (128 threads): column sums / head gradients per slot, then the TMEM drain
    const int wt = tid - 64;  // 0..127
    uint32_t stage = 0, phase = 0;
    float s0 = 0.f, s1 = 0.f;                  // db of columns 2wt, 2wt+1
    float e0 = 0.f, e1 = 0.f, e2 = 0.f;        // extra accumulators
    // Per-row head gradients (spre / dpre) of the next slot are fetched one iteration ahead
    // with one coalesced load and parked in a double-buffered smem strip, so the inner
    // loops below never wait on DRAM.
    float4* s_rows = reinterpret_cast<float4*>(smem_raw + kDwStages * kDwStageBytes + 128);  // [2][64]
    float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch_rows = [&](int64_t it) {
      const int64_t base = (t_begin + (it >> 1)) * 128 + int(it & 1) * 64 + (wt & 63);
      if (job.extra == 1) pre.x = __ldg(args.SPRE + base);
      else pre = __ldg(reinterpret_cast<const float4*>(args.DPRE) + base);
    };
    if (job.extra != 0 && wt < 64 && n_iters > 0) fetch_rows(0);
    for (int64_t it = 0; it < n_iters; ++it) {
      if (job.extra != 0) {
        if (wt < 64) {
          s_rows[(it & 1) * 64 + wt] = pre;
          if (it + 1 < n_iters) fetch_rows(it + 1);
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
      }
      const float4* rows = s_rows + (it & 1) * 64;
      mbar_wait(bar_full + 8 * stage, phase);
      const uint32_t sa = smem_base + stage * kDwStageBytes, sb = sa + 4 * kHalfBlock;
      if (job.db != nullptr && 2 * wt < N && !(args.debug & 4)) {
        const int c = 2 * wt;
        const uint32_t blk = sb + (c >> 6) * kHalfBlock;
#pragma unroll 16
        for (int rr = 0; rr < 64; ++rr) {
          uint32_t u;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(blk + sw128_offset(rr, c & 63)));
          s0 += __uint_as_float(u << 16);
          s1 += __uint_as_float(u & 0xffff0000u);
        }
      }
      if (args.debug & 4) {
      } else if (job.extra == 1) {  // dW9[f] += z8[row, f] * spre[row], f = 2wt, 2wt+1
        const int c = 2 * wt;
        const uint32_t blk = sa + (c >> 6) * kHalfBlock;
#pragma unroll 16
        for (int rr = 0; rr < 64; ++rr) {
          uint32_t u;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(blk + sw128_offset(rr, c & 63)));
          const float w = rows[rr].x;
          e0 = fmaf(__uint_as_float(u << 16), w, e0);
          e1 = fmaf(__uint_as_float(u & 0xffff0000u), w, e1);
        }
      } else if (job.extra == 2) {  // dW11[k, :] += c[row, k] * dpre[row, :], k = wt
        const uint32_t blk = sa + (wt >> 6) * kHalfBlock;
#pragma unroll 16
        for (int rr = 0; rr < 64; ++rr) {
          uint32_t u;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(blk + sw128_offset(rr, (wt & 63) & ~1)));
          const float cv = (wt & 1) ? __uint_as_float(u & 0xffff0000u) : __uint_as_float(u << 16);
          const float4 d4 = rows[rr];
          e0 = fmaf(cv, d4.x, e0);
          e1 = fmaf(cv, d4.y, e1);
          e2 = fmaf(cv, d4.z, e2);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
      if (++stage == kDwStages) { stage = 0; phase ^= 1; }
    }
    if (job.db != nullptr && 2 * wt < N) {
      atomicAdd(job.db + 2 * wt, s0);
      atomicAdd(job.db + 2 * wt + 1, s1);
    }
    if (job.extra == 1) {
      atomicAdd(job.extra_out + 2 * wt, e0);
      atomicAdd(job.extra_out + 2 * wt + 1, e1);
    } else if (job.extra == 2) {
      atomicAdd(job.extra_out + wt * 3 + 0, e0);
      atomicAdd(job.extra_out + wt * 3 + 1, e1);
      atomicAdd(job.extra_out + wt * 3 + 2, e2);
    }
    // ---- drain: TMEM lane quadrant q = warp % 4 holds dW rows 32q .. 32q+31 of each M half
    mbar_wait(bar_done, 0);
    tc_fence_after();
    if (job.b_blocks > 0 && n_iters > 0) {
      const int q = warp & 3;
      for (int mh = 0; mh < job.m_halves; ++mh) {
        const int row = mh * 128 + q * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem + (uint32_t(q * 32) << 16) + mh * N + c0, v);
          tmem_wait_ld();
          if (row < job.rows_valid) {
            float* dst = job.dW + int64_t(row) * job.ld + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                         __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ================================================================ host side
// CTAs per SM over the kernel's lifetime.  With one wave a job gets 12 or 13 of the 148 CTAs (6 %
// imbalance, the slowest CTA sets the kernel time); three waves let the hardware scheduler balance
// 444 smaller units: 5.48 -> 5.22 ms/step (2 waves 5.29, 4: 5.22, 6: 5.26).
constexpr int kDwWaves = 3;
// relative cost of the two jobs that also form a head gradient on the CUDA cores (tuned, round 1)
constexpr double kDwHeadJobWeight = 8.0;
// Ablation switches for profiling (bit flags 1 / 2 / 4 = dW kernel without loads / MMAs / worker
// math: results are WRONG): read ONCE from the environment (LNRF_DEBUG_FLAGS) in lnrf_init, never
// changed afterwards -- there is no run-time switch in the ABI.
static int g_dw_debug = 0;

int init_mlp_tc_bwd() {
  if (const char* e = getenv("LNRF_DEBUG_FLAGS")) g_dw_debug = atoi(e) & 7;
  int rc = upload_tc_tables();
  if (rc) return rc;
  LNRF_CUDA(cudaFuncSetAttribute(nerf_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(kDwSmemBytes)));
  LNRF_CUDA(cudaFuncSetAttribute(nerf_bwd_dw_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
  return LNRF_OK;
}

int nerf_bwd_tc(const float* P, const void* packed, int64_t m, void* ws, int64_t ws_bytes,
                const float* dens, const float* rgb, const float* d_dens, const float* d_rgb, float* G,
                cudaStream_t st) {
  LNRF_REQUIRE(tc_ready(), LNRF_E_INVALID, "lnrf_nerf_mlp_bwd(bf16): call lnrf_init first");
  LNRF_REQUIRE(ws && (uintptr_t)ws % 1024 == 0 && ws_bytes >= tc_workspace_bytes(m, true),
               LNRF_E_WORKSPACE, "lnrf_nerf_mlp_bwd(bf16): workspace %lld < %lld bytes or misaligned",
               (long long)ws_bytes, (long long)tc_workspace_bytes(m, true));
  LNRF_REQUIRE((uintptr_t)G % 16 == 0, LNRF_E_INVALID, "lnrf_nerf_mlp_bwd(bf16): d_params not 16-byte aligned");
  const TcStash s = carve_stash(ws, m);
  const int64_t tiles = ceil_div(m, 128);

  TcBwdArgs a{reinterpret_cast<const uint8_t*>(packed), P, dens, rgb, d_dens, d_rgb, m, s, G};
  {
    const int rc = nerf_bwd_dx_cta2(a, st);
    if (rc) return rc;
  }

  // ---- dW jobs; CTAs are shared out in proportion to the bytes each job streams per tile
  DwArgs d{};
  d.tiles = tiles;
  d.SPRE = s.SPRE;
  d.DPRE = s.DPRE;
  d.debug = g_dw_debug;
  int nj = 0;
  auto add = [&](const uint8_t* A, int ab, const uint8_t* B, int bb, int mh, int rows, int ld, float* dW,
                 float* db, int extra, float* extra_out) {
    d.jobs[nj++] = DwJob{A, B, ab, bb, mh, rows, ld, extra, dW, db, extra_out, 0, 0};
  };
  for (int l = 1; l <= 8; ++l)  // dW_l = h_{l-1}^T g_l (rows 0..255 of Dense_5 for l = 5)
    add(s.H[l - 1], 4, s.G[l], 4, 2, kH, kH, G + kNerf.w[l], G + kNerf.b[l], 0, nullptr);
  add(s.XE, 1, s.G[0], 4, 1, kXE, kH, G + kNerf.w[0], G + kNerf.b[0], 0, nullptr);               // dW0
  add(s.XE, 1, s.G[5], 4, 1, kXE, kH, G + kNerf.w[5] + int64_t(kH) * kH, nullptr, 0, nullptr);   // dW5 skip rows
  add(s.H[8], 4, s.DC, 2, 2, kH, kHC, G + kNerf.w[10], G + kNerf.b[10], 1, G + kNerf.w[9]);      // dW10[:256], dW9
  add(s.DE, 1, s.DC, 2, 1, kDE, kHC, G + kNerf.w[10] + int64_t(kH) * kHC, nullptr, 0, nullptr);  // dW10[256:]
  add(s.C, 2, nullptr, 0, 0, 0, 3, nullptr, nullptr, 2, G + kNerf.w[11]);                        // dW11
  d.n_jobs = nj;
  const int total_ctas = sm_count() * kDwWaves;
  double wsum = 0.0;
  double wj[kDwMaxJobs];
  for (int j = 0; j < nj; ++j) {
    wj[j] = d.jobs[j].a_blocks + d.jobs[j].b_blocks;
    // the two jobs that also form a head gradient on the CUDA cores are bounded by that row
    // loop, not by the bytes they stream
    if (d.jobs[j].extra != 0) wj[j] = kDwHeadJobWeight;
    wsum += wj[j];
  }
  // largest-remainder apportionment so that every SM gets a CTA
  int cnt[kDwMaxJobs];
  double frac[kDwMaxJobs];
  int used = 0;
  for (int j = 0; j < nj; ++j) {
    const double exact = total_ctas * wj[j] / wsum;
    cnt[j] = int(exact);
    if (cnt[j] < 1) cnt[j] = 1;
    frac[j] = exact - cnt[j];
    used += cnt[j];
  }
  while (used < total_ctas) {
    int best = 0;
    for (int j = 1; j < nj; ++j)
      if (frac[j] > frac[best]) best = j;
    ++cnt[best];
    frac[best] -= 1.0;
    ++used;
  }
  int begin = 0;
  for (int j = 0; j < nj; ++j) {
    int c = cnt[j];
    if (int64_t(c) > tiles) c = int(tiles);
    d.jobs[j].cta_begin = begin;
    d.jobs[j].cta_count = c;
    begin += c;
  }
  nerf_bwd_dw_kernel<<<(unsigned)begin, kDwThreads, kDwSmemBytes, st>>>(d);
  LNRF_LAUNCH_CHECK("nerf_bwd_dw_kernel");
  return LNRF_OK;
}

}  // namespace lnrf
