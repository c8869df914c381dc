// K2 on the tensor cores, CTA-pair kernel: the whole NeRFModel forward (model.py:42-62) fused into one
// tcgen05/TMEM kernel issued as cta_group::2 MMAs (M = 256 over two CTAs, see mlp_tc_cta2.cuh).
//
// Per tile the 128 x 256 bf16 activation matrix lives in shared memory as the UMMA A operand (four
// K-major SW128 blocks + a fifth block holding the positional encoding, computed in registers, never
// in HBM).  Two epilogue teams per tile (column halves) read the fp32 accumulator back with
// tcgen05.ld, add the bias, apply ReLU, convert to bf16 and write the next layer's A operand in place.
// The density head rides as output column 128 of the colour-layer GEMM; the 128 -> 3 rgb head runs in
// fp32 FMAs.  With SAVE the activation tile images and row-major 1-bit ReLU masks are streamed to the
// stash (same format as before: the dX / dW kernels are unchanged consumers).
#include <stdlib.h>

#include <type_traits>

#include "mlp_tc_cta2.cuh"

namespace lnrf {

using namespace ptx;

struct C2FwdArgs {
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
  C2Sched sched;
};

__device__ __forceinline__ void c2_fast_sincos(float a, float* s, float* c) {
  const float k = rintf(a * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, a);
  r = fmaf(k, 1.7484555e-7f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}

// packed fp32x2 add (FADD2 on sm_100): {a0, a1} += {b0, b1}
__device__ __forceinline__ void fadd2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 x, y;\n\t"
      "mov.b64 x, {%0, %1};\n\t"
      "mov.b64 y, {%2, %3};\n\t"
      "add.rn.f32x2 x, x, y;\n\t"
      "mov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "f"(b0), "f"(b1));
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// bias + (ReLU) + bf16 pack of 32 accumulator columns [C0, C0 + 32) of the team's half -> four 16-byte row
// chunks of the A tile.  The layer's 128 biases of this half sit in shared memory (`sbias`, staged by the
// team itself, see c2_fwd_team): broadcast 16-byte loads + packed fp32x2 adds -- no constant bank, so the
// ten layers share ONE copy of this code without paying an indexed constant load per bias, and nothing
// per-model lives in process-wide memory (re-entrant).  SAVE: also collects the 32 "pre-activation > 0"
// bits in `mword` (column j -> bit 31-j).
template <bool RELU, int C0, bool SAVE>
__device__ __forceinline__ void c2_epi_store32(const uint32_t (&v)[32], uint32_t blk0, int r, uint32_t& mword,
                                               uint32_t sbias) {
  uint32_t pk[16];
  uint32_t signs = 0;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 bv = lds_f4(sbias + (C0 + j) * 4);
    float f0 = __uint_as_float(v[j]), f1 = __uint_as_float(v[j + 1]);
    float f2 = __uint_as_float(v[j + 2]), f3 = __uint_as_float(v[j + 3]);
    fadd2(f0, f1, bv.x, bv.y);
    fadd2(f2, f3, bv.z, bv.w);
    pk[j / 2] = RELU ? pack_bf16x2_relu(f0, f1) : pack_bf16x2(f0, f1);
    pk[j / 2 + 1] = RELU ? pack_bf16x2_relu(f2, f3) : pack_bf16x2(f2, f3);
    if (SAVE && RELU) {
      signs = __funnelshift_l(__float_as_uint(f0), signs, 1);
      signs = __funnelshift_l(__float_as_uint(f1), signs, 1);
      signs = __funnelshift_l(__float_as_uint(f2), signs, 1);
      signs = __funnelshift_l(__float_as_uint(f3), signs, 1);
    }
  }
  mword = ~signs;
  const uint32_t blk = blk0 + (C0 >> 6) * kABlockBytes;  // blk0 = first of the team's two A blocks
  constexpr int cbase = (C0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    store_row_chunk(blk, r, cbase + q, pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
}

// the team's 128 accumulator columns of a hidden layer -> its two A blocks
template <bool RELU, bool SAVE>
__device__ __forceinline__ void c2_epi_half(uint32_t tm, uint32_t blk0, int r, uint32_t (&mw)[4], uint32_t sbias) {
  uint32_t va[32], vb[32];
  tmem_ld32(tm, va);
  tmem_wait_ld_dep(va);
  tmem_ld32(tm + 32, vb);
  c2_epi_store32<RELU, 0, SAVE>(va, blk0, r, mw[0], sbias);
  tmem_wait_ld_dep(vb);
  tmem_ld32(tm + 64, va);
  c2_epi_store32<RELU, 32, SAVE>(vb, blk0, r, mw[1], sbias);
  tmem_wait_ld_dep(va);
  tmem_ld32(tm + 96, vb);
  c2_epi_store32<RELU, 64, SAVE>(va, blk0, r, mw[2], sbias);
  tmem_wait_ld_dep(vb);
  c2_epi_store32<RELU, 96, SAVE>(vb, blk0, r, mw[3], sbias);
}

// One epilogue team: group g (0/1), column half H.  Per layer the team stages its 128 biases in shared
// memory: every thread fetches ONE value from the packed buffer's SmallParams image before it waits for
// the accumulator (latency hidden), stores it after the wait, and the team barrier publishes it.  No
// barrier is needed before the next layer overwrites the buffer: that layer's accumulator only
// completes after every thread of the team has arrived on a_ready, i.e. finished reading this layer's.
template <bool SAVE, int H>
__device__ __forceinline__ void c2_fwd_team(const C2FwdArgs& args, const C2Ctx& cx, int g, int64_t tiles) {
  constexpr int h = H;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int team = 2 * g + H;
    // ===== epilogue team (g, h): thread r owns row r of group g's tile, columns 128 h .. 128 h + 127
    const int r = tid & 127;
    const bool leader = r == 0;
    const uint32_t sA = cx.sA0 + g * kPairTileBytes;
    const uint32_t blk0 = sA + 2 * h * kABlockBytes;
    const uint32_t tm = cx.tmem + (uint32_t((warp & 3) * 32) << 16) + g * 256 + h * 128;
    const uint32_t bar_a = cx.bars + C2Smem::a_ready + 8 * g, bar_acc = cx.bars + C2Smem::acc_full + 8 * g;
    const int64_t cid = cluster_id_x(), ncl = nclusters_x();
    const SmallParams* small = reinterpret_cast<const SmallParams*>(args.packed + kSmallOffset);
    const uint32_t sbias = cx.bars + C2Smem::bias + uint32_t(g * 256 + H * 128) * 4u;
    uint32_t de[12];  // d_emb of the current tile (team 1 only)

    // inputs + sinusoidal_emb of tile `tile` -> A block 4 (x_emb) and `de` (d_emb); team h == 1 only
    auto prologue = [&](int64_t tile) {
      const bool tile_ok = tile < tiles;
      const int64_t s = tile * 128 + r;
      float px[3] = {0.f, 0.f, 0.f}, dv[3] = {0.f, 0.f, 0.f};
      if (tile_ok && s < args.m) {
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
      if (SAVE) {  // block 4 may still be read by the previous tile's d_emb bulk store (this team's leader issued it)
        if (leader) bulk_wait_read0();
        team_bar(team);
      }
      {
        uint32_t pk[32];
#pragma unroll
        for (int dim = 0; dim < 3; ++dim) {
          float sn[kXFreqs], cs[kXFreqs];
#pragma unroll
          for (int f = 0; f < kXFreqs; ++f) c2_fast_sincos(px[dim] * float(1 << f), &sn[f], &cs[f]);
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
#pragma unroll
      for (int dim = 0; dim < 3; ++dim) {
        float sn[kDFreqs], cs[kDFreqs];
#pragma unroll
        for (int f = 0; f < kDFreqs; ++f) c2_fast_sincos(dv[dim] * float(1 << f), &sn[f], &cs[f]);
        de[dim * 4 + 0] = pack_bf16x2(sn[0], sn[1]);
        de[dim * 4 + 1] = pack_bf16x2(sn[2], sn[3]);
        de[dim * 4 + 2] = pack_bf16x2(cs[0], cs[1]);
        de[dim * 4 + 3] = pack_bf16x2(cs[2], cs[3]);
      }
      fence_proxy_async_smem();
      if (SAVE) {
        team_bar(team);
        if (leader) {
          if (tile_ok) bulk_s2g(args.stash.XE + tile * kABlockBytes, sA + 4 * kABlockBytes, kABlockBytes);
          bulk_commit();
        }
      }
    };
    auto tile_of = [&](int64_t t) { return ((cid + t * ncl) * 2 + g) * 2 + int64_t(cx.rank); };

    if (cx.my_iters > 0) {
      if (h == 1) prologue(tile_of(0));
      c2_arrive_a(bar_a, cx.rank);  // T0 of the first tile may start
    }
    for (int64_t t = 0; t < cx.my_iters; ++t) {
      const int64_t tile = tile_of(t);
      const bool tile_ok = tile < tiles;
      const int64_t s = tile * 128 + r;
      const bool valid = tile_ok && s < args.m;
      uint4* mask_row = (SAVE && tile_ok) ? reinterpret_cast<uint4*>(args.stash.MASK + ((tile * 9) * 128 + r) * 8) + h
                                          : nullptr;
      // ---- hidden layers T0..T8 (ten layers per tile: the barrier parity of layer TL is TL & 1).  T0..T7 share
      // one copy of the code (runtime TL); T8 (no ReLU, d_emb) is its own instance.
      auto layer = [&](int TL, auto last_c) {
        constexpr bool LAST = decltype(last_c)::value;
        const float bval = __ldg(&small->b[TL][H * 128 + r]);  // in flight while we wait for the accumulator
#ifdef LNRF_C2_TRACE
        const long long E0 = clock64();
#endif
        mbar_wait(bar_acc, TL & 1);
        tc_fence_after();
#ifdef LNRF_C2_TRACE
        const long long E1 = clock64();
#endif
        if (SAVE && leader) bulk_wait_read0();  // this team's previous image (same two blocks) has left smem
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + r * 4), "f"(bval) : "memory");
        team_bar(team);
#ifdef LNRF_C2_TRACE
        const long long Ea = clock64();
#endif
        uint32_t mw[4];
        c2_epi_half<!LAST, SAVE>(tm, blk0, r, mw, sbias);  // Dense_8 feeds the heads raw (model.py:53-58)
#ifdef LNRF_C2_TRACE
        const long long Eb = clock64();
#endif
        if (SAVE && !LAST && mask_row) mask_row[TL * 256] = make_uint4(mw[0], mw[1], mw[2], mw[3]);
        if (LAST && h == 1) {  // x_emb is dead after T5: block 4 now carries d_emb (24 cols) + zeros
          const uint32_t blk = sA + 4 * kABlockBytes;
          store_row_chunk(blk, r, 0, de[0], de[1], de[2], de[3]);
          store_row_chunk(blk, r, 1, de[4], de[5], de[6], de[7]);
          store_row_chunk(blk, r, 2, de[8], de[9], de[10], de[11]);
#pragma unroll
          for (int c = 3; c < 8; ++c) store_row_chunk(blk, r, c, 0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
        if (SAVE) {
          team_bar(team);
          if (leader) {
            if (tile_ok) {
              bulk_s2g(args.stash.H[TL] + tile * kTileBytes + 2 * h * kABlockBytes, blk0, 2 * kABlockBytes);
              if (LAST && h == 1) bulk_s2g(args.stash.DE + tile * kABlockBytes, sA + 4 * kABlockBytes, kABlockBytes);
            }
            bulk_commit();
          }
        }
#ifdef LNRF_C2_TRACE
        const long long Ed = clock64();
#endif
        c2_arrive_a(bar_a, cx.rank);
#ifdef LNRF_C2_TRACE
        if (leader && cid == 0 && cx.rank == 0 && t < 3) {
          const int base = 1024 + int(((t * 10 + TL) * 4 + team) * 6);
          C2_TRACE(base, E0); C2_TRACE(base + 1, E1); C2_TRACE(base + 2, Ea); C2_TRACE(base + 3, Eb);
          C2_TRACE(base + 4, Ed); C2_TRACE(base + 5, clock64());
        }
#endif
      };
#pragma unroll 1
      for (int TL = 0; TL < 8; ++TL) layer(TL, std::false_type{});
      layer(8, std::true_type{});
      // ---- T9: colour layer (columns 0..127, team 0) + density column 128 (team 1) and the fp32 rgb head.
      // b10 is staged like a hidden layer's bias; the 128 x 3 rgb head weights are read with 16-byte
      // read-only loads (every thread the same address: an L1 broadcast, once per tile).
      const float bval9 = h == 0 ? __ldg(&small->b10[r]) : __ldg(&small->b9);
      mbar_wait(bar_acc, 1u);
      tc_fence_after();
      if (h == 0) {
        if (SAVE && leader) bulk_wait_read0();  // blocks 0,1 get the colour-hidden image once the z8 image has left smem
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + r * 4), "f"(bval9) : "memory");
        team_bar(team);
        const float4* w11v = reinterpret_cast<const float4*>(small->w11);
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        uint32_t mwc[4];
#pragma unroll
        for (int c0 = 0; c0 < kHC; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tm + c0, v);
          tmem_wait_ld_dep(v);
          uint32_t pk[16];
          uint32_t signs = 0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {  // 4 columns: 4 biases (one 16-byte smem load), 12 head weights (three 16-byte loads)
            const float4 bv = lds_f4(sbias + (c0 + j) * 4);
            const float4 wa = __ldg(w11v + (c0 + j) * 3 / 4), wb = __ldg(w11v + (c0 + j) * 3 / 4 + 1),
                         wc = __ldg(w11v + (c0 + j) * 3 / 4 + 2);
            const float p0 = __uint_as_float(v[j]) + bv.x, p1 = __uint_as_float(v[j + 1]) + bv.y;
            const float p2 = __uint_as_float(v[j + 2]) + bv.z, p3 = __uint_as_float(v[j + 3]) + bv.w;
            const float h0 = fmaxf(p0, 0.0f), h1 = fmaxf(p1, 0.0f), h2 = fmaxf(p2, 0.0f), h3 = fmaxf(p3, 0.0f);  // model.py:59
            o0 = fmaf(h0, wa.x, o0); o1 = fmaf(h0, wa.y, o1); o2 = fmaf(h0, wa.z, o2);
            o0 = fmaf(h1, wa.w, o0); o1 = fmaf(h1, wb.x, o1); o2 = fmaf(h1, wb.y, o2);
            o0 = fmaf(h2, wb.z, o0); o1 = fmaf(h2, wb.w, o1); o2 = fmaf(h2, wc.x, o2);
            o0 = fmaf(h3, wc.y, o0); o1 = fmaf(h3, wc.z, o1); o2 = fmaf(h3, wc.w, o2);
            if (SAVE) {
              pk[j / 2] = pack_bf16x2(h0, h1);
              pk[j / 2 + 1] = pack_bf16x2(h2, h3);
              signs = __funnelshift_l(__float_as_uint(p0), signs, 1);
              signs = __funnelshift_l(__float_as_uint(p1), signs, 1);
              signs = __funnelshift_l(__float_as_uint(p2), signs, 1);
              signs = __funnelshift_l(__float_as_uint(p3), signs, 1);
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
        if (valid) {
          args.rgb[s * 3 + 0] = tanhf(o0 + __ldg(&small->b11[0]));  // model.py:60
          args.rgb[s * 3 + 1] = tanhf(o1 + __ldg(&small->b11[1]));
          args.rgb[s * 3 + 2] = tanhf(o2 + __ldg(&small->b11[2]));
        }
        if (SAVE) {
          fence_proxy_async_smem();
          team_bar(team);
          if (leader) {
            if (tile_ok) bulk_s2g(args.stash.C + tile * 2 * kABlockBytes, sA, 2 * kABlockBytes);
            bulk_commit();
          }
        }
      } else {
        uint32_t v[32];
        tmem_ld32(tm, v);  // team 1's first column = accumulator column 128 = Dense_9 pre-activation
        tmem_wait_ld_dep(v);
        if (valid) args.dens[s] = softplus_f(__uint_as_float(v[0]) + bval9);  // model.py:57
        if (t + 1 < cx.my_iters) prologue(tile_of(t + 1));  // T9's MMAs are complete: block 4 is free
      }
      if (t + 1 < cx.my_iters) {
        c2_arrive_a(bar_a, cx.rank);  // (its tcgen05 fence orders these TMEM reads before the next tile's T0 MMAs)
      }
    }
    if (SAVE && leader) bulk_wait0();  // all stash stores complete before the CTA exits
}

template <bool SAVE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kC2Threads, 1)
nerf_fwd_cta2_kernel(const __grid_constant__ C2FwdArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int64_t tiles = (args.m + 127) / 128;
  const int64_t quads = (tiles + 3) / 4;  // 2 groups x 2 CTAs
  const C2Ctx cx = c2_setup(smem_raw, quads);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 16) {
    if (lane == 0) c2_producer(args.packed, args.sched, cx.my_iters, cx.rank, cx.sW, cx.bars);
  } else if (warp == 17) {
    if (cx.rank == 0) c2_mma(args.sched, cx.my_iters, cx.sA0, cx.sW, cx.bars, cx.tmem);
    else c2_relay(args.sched, cx.my_iters, cx.bars);
  } else if (warp & 4) {
    c2_fwd_team<SAVE, 1>(args, cx, warp >> 3, tiles);
  } else {
    c2_fwd_team<SAVE, 0>(args, cx, warp >> 3, tiles);
  }
  c2_teardown(cx);
}

// ---------------------------------------------------------------- host side
static C2Sched g_fwd_sched;
static int g_max_clusters = 0;

C2Sched c2_make_sched(const ChunkInfo* tab, int n, int layers) {
  C2Sched s{};
  s.layers = layers;
  int loads = 0;  // ring position of the next fill
  int i = 0;
  for (int L = 0; L < layers; ++L) {
    int j = i;
    while (j < n && tab[j].tlayer == L) ++j;
    const int cnt = j - i;
    const bool shared = cnt <= kC2Stages;
    int slot_of[8];
    for (int gi = 0; gi < 2; ++gi)
      for (int c = 0; c < cnt; ++c) {
        const int g = gi;
        C2Step st{};
        st.offset = tab[i + c].offset;
        st.n = uint16_t(tab[i + c].n);
        st.ablock = uint8_t(tab[i + c].ablock);
        st.flags = uint8_t((g ? S_G1 : 0) | (c == 0 ? S_FIRST : 0) | (c == cnt - 1 ? S_LAST : 0));
        if (!shared || gi == 0) {
          st.flags |= S_LOAD;
          slot_of[c] = loads++ % kC2Stages;
        }
        if (!shared || gi == 1) st.flags |= S_RELEASE;
        st.slot = uint8_t(slot_of[c]);
        s.step[s.steps++] = st;
      }
    i = j;
  }
  return s;
}

int c2_max_clusters() { return g_max_clusters; }

int init_mlp_tc_cta2_fwd() {
  const ChunkTable t = build_chunk_table();
  g_fwd_sched = c2_make_sched(t.f, kTcChunks, kTcLayers);
  LNRF_CUDA(cudaFuncSetAttribute(nerf_fwd_cta2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2Smem::total));
  LNRF_CUDA(cudaFuncSetAttribute(nerf_fwd_cta2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2Smem::total));
  // how many CTA pairs the device can hold at once (one per TPC with this much shared memory)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(sm_count()), 1, 1);
  cfg.blockDim = dim3(kC2Threads, 1, 1);
  cfg.dynamicSmemBytes = C2Smem::total;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, nerf_fwd_cta2_kernel<true>, &cfg);
  if (e != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    n = sm_count() / 2;
  }
  g_max_clusters = n;
  if (getenv("LNRF_VERBOSE"))
    fprintf(stderr, "lnrf: CTA-pair kernels: %d co-resident clusters of 2 (occupancy query: %s), %d SMs\n", n,
            e == cudaSuccess ? "ok" : cudaGetErrorName(e), sm_count());
  return LNRF_OK;
}

int nerf_fwd_cta2(const void* packed, const float* x, const float* d, const float* rays, const float* ts, int64_t m,
                  int T, bool save, const TcStash& stash, float* dens, float* rgb, cudaStream_t st) {
  C2FwdArgs a{reinterpret_cast<const uint8_t*>(packed), x, d, rays, ts, T, m, dens, rgb, stash, g_fwd_sched};
  const int64_t quads = (ceil_div(m, 128) + 3) / 4;
  int64_t clusters = g_max_clusters;
  if (clusters > quads) clusters = quads;
  const unsigned grid = unsigned(clusters * 2);
  if (save) nerf_fwd_cta2_kernel<true><<<grid, kC2Threads, C2Smem::total, st>>>(a);
  else nerf_fwd_cta2_kernel<false><<<grid, kC2Threads, C2Smem::total, st>>>(a);
  LNRF_LAUNCH_CHECK("nerf_fwd_cta2_kernel");
  return LNRF_OK;
}

}  // namespace lnrf

#ifdef LNRF_C2_TRACE
extern "C" int lnrf_debug_c2_trace(unsigned long long* out_host, int count) {
  return (int)cudaMemcpyFromSymbol(out_host, lnrf::g_c2_trace, sizeof(unsigned long long) * count);
}
#endif
