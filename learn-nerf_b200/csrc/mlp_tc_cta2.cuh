// Shared scaffolding of the CTA-PAIR tcgen05 kernels (forward: mlp_tc_cta2_fwd.cu, dX chain:
// mlp_tc_cta2_bwd.cu): `tcgen05.mma.cta_group::2` over a cluster of two CTAs on one TPC.
//
// Each CTA keeps TWO resident 128-sample tiles ("groups" g = 0, 1).  A group's MMAs are M = 256:
// rows 0..127 are this CTA's tile, rows 128..255 the peer CTA's tile of the same group, and each CTA
// holds only HALF of every weight chunk (N/2 rows of the B operand), so a weight chunk is fetched
// from L2 once per 256 samples while each SM reads just 64 B/clk of operands out of its shared
// memory (an M=128 N=256 cta_group::1 MMA needs 96 B/clk of the 128 B/clk there is).  That headroom
// is what lets the two groups run half a period apart: while the tensor pipe works on layer L of
// group 0, all epilogue warps of group 1 convert ITS layer-(L-1) accumulator, and vice versa, at
// WHOLE-layer granularity (no N split, no lockstep stall).
//
// 576 threads per CTA: warps 0-15 = four epilogue teams (group g, column half h) of 4 warps each
// (team = 4 (2 g + h); thread r of a team <-> tile row r <-> TMEM lane r, columns 128 h .. 128 h + 127),
// warp 16 = weight producer (bulk copies of this CTA's chunk halves), warp 17 = MMA issuer in the
// leader CTA (cluster rank 0) and barrier relay in the peer CTA.
//
// Cross-CTA protocol (all barriers are mbarriers at the same shared-memory offset in both CTAs):
//   full[s]     this CTA's half of ring slot s has landed.  The leader's copy counts 2 arrivals: its own
//               producer's expect_tx and the peer relay's remote arrive after the PEER's half landed.
//   empty[s]    multicast tcgen05.commit (both CTAs): the MMAs reading slot s have retired.
//   a_ready[g]  (leader's copy only) A operand of group g written + accumulator drained in BOTH CTAs: 256
//               local epilogue threads + one releasing remote arrive per peer epilogue warp (8).
//   acc_full[g] multicast tcgen05.commit: the layer's accumulator of group g is complete.
// The MMA issuer, both producers and the relay walk the SAME static order (tile quad, layer, group,
// chunk), so no barrier can be waited on out of order.
#pragma once
#include "tc_common.cuh"

namespace lnrf {

constexpr uint32_t kPairTileBytes = 5 * kABlockBytes;  // one group's A tile: 4 activation blocks + embedding block

// tcgen05.wait::ld that also carries a register dependency on the loaded values, so the
// compiler cannot schedule their consumers above the wait.
__device__ __forceinline__ void tmem_wait_ld_dep(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]),
                 "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]),
                 "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                 "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

constexpr int kC2Threads = 576;
constexpr int kC2Stages = 4;

struct C2Smem {
  static constexpr uint32_t a_off = 0;                               // group 0 tile, then group 1 tile (5 blocks each)
  static constexpr uint32_t w_off = 2 * kPairTileBytes;              // 163,840: ring of 4 x 16 KB chunk halves
  static constexpr uint32_t bar_off = w_off + kC2Stages * kChunkBytes128;  // 229,376
  static constexpr uint32_t total = bar_off + 128 + 2048;
  static constexpr uint32_t bias = 128;     // 2 groups x 256 floats: the current layer's biases (forward) / w9 (dX)
  static constexpr uint32_t full = 0;       // [4]
  static constexpr uint32_t empty = 32;     // [4]
  static constexpr uint32_t a_ready = 64;   // [2]
  static constexpr uint32_t acc_full = 80;  // [2]
  static constexpr uint32_t tmem_slot = 96;
};
static_assert(C2Smem::total <= 232448, "CTA-pair kernel exceeds 227 KB of shared memory");

// The static schedule of one tile quad: one step = the four K=16 MMAs of one (group, weight chunk).
// A chunk is a [n rows x 64 k] bf16 K-major SW128 image at `offset` in the packed buffer; rank r of the
// pair loads rows [r n/2, (r+1) n/2) into ring slot `slot`.  Layers of at most kC2Stages chunks are
// SHARED by the two groups: group 0's steps load the chunks (S_LOAD) and leave them in the ring, group
// 1's steps reuse them and release the slots (S_RELEASE) -- the weight traffic L2 -> SM halves and a
// chunk of the next layer can be fetched a whole burst (2 k clk) before its first use, which is what
// the stream needs when the stash stores load the L2 (measured: 1.8 - 4.3 k clk of "weights not there
// yet" per group-layer without the sharing).  Longer layers (5 chunks: the skip and colour layers)
// stream their chunks once per group.
constexpr uint8_t S_LOAD = 1;     // producer fills `slot` (issuer / relay wait for it)
constexpr uint8_t S_RELEASE = 2;  // issuer commits empty[slot] behind these MMAs
constexpr uint8_t S_FIRST = 4;    // first chunk of the group-layer: wait a_ready[g], overwrite the accumulator
constexpr uint8_t S_LAST = 8;     // last chunk: commit acc_full[g]
constexpr uint8_t S_G1 = 16;      // group 1
constexpr int kC2MaxSteps = 80;
struct C2Step {
  uint32_t offset;
  uint16_t n;       // 256, or 144 for the colour layer (+ density column)
  uint8_t ablock;   // A block the MMA reads: 0..3 activations, 4 = embedding block
  uint8_t flags;
  uint8_t slot;
  uint8_t pad[3];
};
struct C2Sched {
  C2Step step[kC2MaxSteps];
  int steps;   // per tile quad
  int layers;  // group-layers per tile and group (10 forward, 9 backward)
};

#ifdef LNRF_C2_TRACE
// profiling build only (-DLNRF_C2_TRACE): clock64 stamps of cluster 0's leader CTA
static __device__ unsigned long long g_c2_trace[8192];
#define C2_TRACE(idx, val) do { if ((idx) < 8192) g_c2_trace[(idx)] = (val); } while (0)
#endif

namespace ptx {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t nclusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
// all threads of both CTAs
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}
// same without ordering: pure "an asynchronous copy into MY shared memory has landed" notifications
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc2(uint32_t result_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(result_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= [A_cta0; A_cta1] * [B_cta0; B_cta1]^T, M = 256; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma2_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on `bar` in BOTH CTAs once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(uint16_t(3))
               : "memory");
}

}  // namespace ptx

// named barrier of one epilogue team (128 threads); ids 1..4
__device__ __forceinline__ void team_bar(int team) { asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory"); }

// ---- weight producer (lane 0 of warp 16, BOTH CTAs): this CTA's half of every S_LOAD step, static order
__device__ __forceinline__ void c2_producer(const uint8_t* packed, const C2Sched& sc, int64_t my_iters, uint32_t rank,
                                            uint32_t sW, uint32_t bars) {
  using namespace ptx;
  uint32_t ph = 0;  // bit s = phase of empty[s] the next fill of slot s waits for
  for (int64_t t = 0; t < my_iters; ++t)
    for (int i = 0; i < sc.steps; ++i) {
      const C2Step st = sc.step[i];
      if (!(st.flags & S_LOAD)) continue;
      const uint32_t half = uint32_t(st.n) * 64u;  // (n / 2) rows x 128 B
      const uint32_t s = st.slot;
      mbar_wait(bars + C2Smem::empty + 8 * s, ((ph >> s) & 1u) ^ 1u);
      ph ^= 1u << s;
      mbar_arrive_expect_tx(bars + C2Smem::full + 8 * s, half);
      bulk_g2s(sW + s * kChunkBytes128, packed + st.offset + rank * half, half, bars + C2Smem::full + 8 * s);
    }
}

// ---- barrier relay (warp 17 of the PEER CTA): forwards "my chunk half landed" to the leader's full
// barriers in ring order.  Relaxed arrives: the data was written by the bulk-copy engine and is only
// ever read by the tensor core; a releasing arrive costs ~450 clk each, serialised in this one thread
// (measured: the issuer then waits ~1800 clk per group-layer for weights that are long in place).
__device__ __forceinline__ void c2_relay(const C2Sched& sc, int64_t my_iters, uint32_t bars) {
  using namespace ptx;
  uint32_t ph = 0;  // bit s = phase of full[s] that completes next
  for (int64_t t = 0; t < my_iters; ++t)
    for (int i = 0; i < sc.steps; ++i) {
      const C2Step st = sc.step[i];
      if (!(st.flags & S_LOAD)) continue;
      const uint32_t s = st.slot;
      mbar_wait(bars + C2Smem::full + 8 * s, (ph >> s) & 1u);
      ph ^= 1u << s;
      if (elect_one()) mbar_arrive_remote_relaxed(bars + C2Smem::full + 8 * s, 0);
      __syncwarp();
    }
}

// "A operand of my tile written, my accumulator reads done" -> the LEADER's a_ready barrier.  Leader CTA:
// every epilogue thread arrives locally.  Peer CTA: one remote arrive per warp, issued by the epilogue
// warps themselves (a relay thread serialises them: measured +2 k clk per layer).  The remote arrive is
// RELAXED: what it publishes is (a) shared-memory writes, already pushed to the async proxy by each
// writer's fence.proxy.async and ordered before lane 0 by __syncwarp, and (b) completed tcgen05.ld reads
// (tcgen05.fence::before_thread_sync).  A releasing cluster-scope arrive would also wait for the warp's
// outstanding GLOBAL stores (the ReLU masks of the stash): measured ~4 k clk per layer on the peer.
__device__ __forceinline__ void c2_arrive_a(uint32_t bar_a, uint32_t rank) {
  using namespace ptx;
  tc_fence_before();
  if (rank == 0) {
    mbar_arrive(bar_a);
  } else {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_remote_relaxed(bar_a, 0);
  }
}

// ---- MMA issuer (warp 17 of the LEADER CTA; the whole warp runs the loop, one elected lane issues)
__device__ __forceinline__ void c2_mma(const C2Sched& sc, int64_t my_iters, uint32_t sA, uint32_t sW, uint32_t bars,
                                       uint32_t tmem) {
  using namespace ptx;
  const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B
  const uint32_t a_lo0 = ((sA & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t b_lo0 = ((sW & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t idesc256 = umma_idesc_bf16(256, 256), idesc144 = umma_idesc_bf16(256, kNColor);
  uint32_t ph = 0;          // bit s = phase of full[s] that completes next
  uint32_t ka[2] = {0, 0};  // group-layers started per group -> parity of a_ready[g]
#ifdef LNRF_C2_TRACE
  const bool tr = cluster_id_x() == 0 && (threadIdx.x & 31) == 0;
  long long T0 = 0, T1 = 0, fw = 0;
  int gl = 0;
#endif
  for (int64_t t = 0; t < my_iters; ++t)
    for (int i = 0; i < sc.steps; ++i) {
      const C2Step st = sc.step[i];
      const uint32_t g = (st.flags & S_G1) ? 1u : 0u;
      const uint32_t s = st.slot;
      if (st.flags & S_FIRST) {
#ifdef LNRF_C2_TRACE
        T0 = clock64();
        fw = 0;
#endif
        mbar_wait(bars + C2Smem::a_ready + 8 * g, ka[g] & 1u);  // both CTAs' tiles (256 local + 8 peer-warp arrives)
        ++ka[g];
        tc_fence_after();
#ifdef LNRF_C2_TRACE
        T1 = clock64();
#endif
      }
      if (st.flags & S_LOAD) {
#ifdef LNRF_C2_TRACE
        const long long F0 = clock64();
#endif
        mbar_wait(bars + C2Smem::full + 8 * s, (ph >> s) & 1u);  // both halves landed
        ph ^= 1u << s;
        tc_fence_after();
#ifdef LNRF_C2_TRACE
        fw += clock64() - F0;
#endif
      }
      if (elect_one()) {
        const uint32_t idesc = st.n == 256 ? idesc256 : idesc144;
        const uint32_t ax = a_lo0 + (g * kPairTileBytes + uint32_t(st.ablock) * kABlockBytes) / 16u;
        const uint32_t bx = b_lo0 + s * (kChunkBytes128 >> 4);
        const uint32_t d = tmem + g * 256u;
        umma2_bf16_lohi(d, ax, bx, desc_hi, idesc, (st.flags & S_FIRST) ? 0u : 1u);
        umma2_bf16_lohi(d, ax + 2, bx + 2, desc_hi, idesc, 1u);
        umma2_bf16_lohi(d, ax + 4, bx + 4, desc_hi, idesc, 1u);
        umma2_bf16_lohi(d, ax + 6, bx + 6, desc_hi, idesc, 1u);
        if (st.flags & S_RELEASE) umma2_commit_mc(bars + C2Smem::empty + 8 * s);
        if (st.flags & S_LAST) umma2_commit_mc(bars + C2Smem::acc_full + 8 * g);
      }
      __syncwarp();
#ifdef LNRF_C2_TRACE
      if ((st.flags & S_LAST) && tr && t < 3) {
        const int base = gl * 4;
        C2_TRACE(base, T0); C2_TRACE(base + 1, T1); C2_TRACE(base + 2, clock64()); C2_TRACE(base + 3, fw);
        ++gl;
      }
#endif
    }
}

// ---- common prologue / epilogue of both kernels
struct C2Ctx {
  uint32_t sA0, sW, bars, tmem, rank;
  int64_t my_iters;
};

__device__ __forceinline__ C2Ctx c2_setup(uint8_t* smem_raw, int64_t quads) {
  using namespace ptx;
  C2Ctx c;
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) {  // SW128 operands need 1024-byte aligned blocks
    if (threadIdx.x == 0) printf("lnrf: dynamic smem base 0x%x not 1024-aligned\n", smem_base);
    __trap();
  }
  c.sA0 = smem_base + C2Smem::a_off;
  c.sW = smem_base + C2Smem::w_off;
  c.bars = smem_base + C2Smem::bar_off;
  c.rank = cluster_ctarank();
  const int64_t cid = cluster_id_x(), ncl = nclusters_x();
  c.my_iters = quads > cid ? (quads - cid + ncl - 1) / ncl : 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    const bool leader = c.rank == 0;
    for (int s = 0; s < kC2Stages; ++s) {
      mbar_init(c.bars + C2Smem::full + 8 * s, leader ? 2 : 1);
      mbar_init(c.bars + C2Smem::empty + 8 * s, 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(c.bars + C2Smem::a_ready + 8 * g, 256 + 8);  // leader's copy only: 256 local threads + 8 peer warps
      mbar_init(c.bars + C2Smem::acc_full + 8 * g, 1);
    }
    fence_barrier_init();
  }
  if (warp == 16) {
    tmem_alloc2(c.bars + C2Smem::tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits + TMEM allocation of BOTH CTAs visible before any cross-CTA signal
  tc_fence_after();
  c.tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + C2Smem::bar_off + C2Smem::tmem_slot);
  return c;
}

__device__ __forceinline__ void c2_teardown(const C2Ctx& c) {
  using namespace ptx;
  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still signal it / read its smem
  if ((threadIdx.x >> 5) == 16) tmem_dealloc2(c.tmem, 512);
}

}  // namespace lnrf
