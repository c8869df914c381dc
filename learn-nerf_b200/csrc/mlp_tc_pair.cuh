// Shared scaffolding of the "pair" tcgen05 kernels (forward and dX chain).
//
// One CTA per SM owns TWO 128-sample tiles (A and B).  Every streamed weight chunk
// ([128 n x 64 k] bf16, 16 KB, 4-slot ring filled by the bulk-copy engine) feeds one MMA group
// per tile, so the L2 -> SM weight traffic per sample is half of a one-tile CTA, and the ring
// prefetches three chunks (1.5 K MMA cycles) ahead of the tensor pipe.  Each tile has its own
// 256-column fp32 accumulator in TMEM (512 columns in total) and its own group of four
// epilogue warps, so the two epilogues run side by side while the first MMAs of the next layer
// are already issued for whichever tile finished first.
//
// 320 threads: warps 0-3 = epilogue of tile A (thread r <-> row r <-> TMEM lane r),
// warps 4-7 = epilogue of tile B, warp 8 = weight producer, warp 9 = MMA issuer.
#pragma once
#include "tc_common.cuh"

namespace lnrf {

constexpr int kPairThreads = 320;
constexpr uint32_t kPairTileBytes = 5 * kABlockBytes;  // 4 activation blocks + embedding block

struct PairSmem {
  static constexpr uint32_t a_off = 0;                                     // tile A, then tile B
  static constexpr uint32_t w_off = 2 * kPairTileBytes;                    // 163,840
  static constexpr uint32_t bar_off = w_off + 4 * kChunkBytes128;          // 229,376 (64 KB ring)
  static constexpr uint32_t total = bar_off + 256;
  // barrier map (8 B each, relative to bar_off)
  static constexpr uint32_t full = 0;                     // [<= 4 stages]
  static constexpr uint32_t empty = 32;                   // [<= 4 stages]
  // per tile X (8 X bytes further): epilogue -> MMA (128 arrivals each)
  static constexpr uint32_t a_ready0 = 64;                 // A blocks 0,1 written, accumulator half 0 drained
  static constexpr uint32_t a_ready1 = a_ready0 + 16;     // A blocks 2,3 (+ embedding block) written
  static constexpr uint32_t drained1 = a_ready1 + 16;     // accumulator half 1 read back
  // MMA -> epilogue (tcgen05.commit)
  static constexpr uint32_t acc0 = drained1 + 16;         // accumulator half 0 complete
  static constexpr uint32_t acc1 = acc0 + 16;             // accumulator half 1 complete
  static constexpr uint32_t tmem_slot = acc1 + 16;
};
static_assert(PairSmem::total <= 232448, "pair kernel exceeds 227 KB of shared memory");

// barrier among the 128 epilogue threads of one tile (named barriers 1 and 2)
__device__ __forceinline__ void bulk_wait_read1() {  // all but the most recent bulk group have left smem
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void pair_bar(int X) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + X) : "memory");
}

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

// Producer (one thread): stream `nchunks` weight chunks per tile pair through the ring.
template <bool FULLN>
__device__ __forceinline__ void pair_producer(const uint8_t* packed, const ChunkInfo* chunks, int nchunks,
                                              int64_t my_pairs, uint32_t sW, uint32_t bars) {
  constexpr int kPairStages = PairCfg<FULLN>::stages;
  constexpr uint32_t kPairSlotBytes = PairCfg<FULLN>::slot_bytes;
  uint32_t stage = 0, phase = 0;
  for (int64_t t = 0; t < my_pairs; ++t) {
    for (int ci = 0; ci < nchunks; ++ci) {
      const uint32_t bytes = uint32_t(chunks[ci].n) * 128u;
      ptx::mbar_wait(bars + PairSmem::empty + 8 * stage, phase ^ 1);
      ptx::mbar_arrive_expect_tx(bars + PairSmem::full + 8 * stage, bytes);
      ptx::bulk_g2s(sW + stage * kPairSlotBytes, packed + chunks[ci].offset, bytes,
                    bars + PairSmem::full + 8 * stage);
      if (++stage == kPairStages) { stage = 0; phase ^= 1; }
    }
  }
}

// Per-chunk control words of the MMA issuer in their own small constant array (one
// constant-bank load per chunk; bit meanings: PM_* in tc_common.cuh).
struct PairMeta {
  uint32_t f2[kF2Chunks];
  uint32_t b2[kB2Chunks];
  uint32_t f_full[kTcChunks];  // full-N schedule over the chunks of table `f`
};
static __constant__ PairMeta c_pair_meta;

static int upload_pair_meta() {
  const ChunkTable t = build_chunk_table();
  PairMeta m{};
  for (int i = 0; i < kF2Chunks; ++i) m.f2[i] = uint32_t(t.f2[i].last);
  for (int i = 0; i < kB2Chunks; ++i) m.b2[i] = uint32_t(t.b2[i].last);
  for (int i = 0; i < kTcChunks; ++i) m.f_full[i] = uint32_t(t.f[i].last);
  LNRF_CUDA(cudaMemcpyToSymbol(c_pair_meta, &m, sizeof(m)));
  return LNRF_OK;
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

// D[tmem] (+)= A * B with the descriptors given as (lo, hi) words
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
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

// MMA issuer (the whole warp runs the loop so that every operand stays in uniform registers;
// one elected lane issues): per chunk one group of four K=16 MMAs per tile, in the order and
// with the waits/commits its control word prescribes (tc_common.cuh, emit_layer).
template <bool FULLN>
__device__ __forceinline__ void pair_mma(const uint32_t* meta, int nchunks, int64_t my_pairs, uint32_t sA,
                                         uint32_t sW, uint32_t bars, uint32_t tmem) {
  using namespace ptx;
  constexpr int kPairStages = PairCfg<FULLN>::stages;
  constexpr uint32_t kPairSlotBytes = PairCfg<FULLN>::slot_bytes;
  // descriptor words (see umma_desc_sw128_kmajor): lo = addr >> 4 | LBO(1) << 16; hi is constant
  const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t a_lo0 = ((sA & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t b_lo0 = ((sW & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t idesc128 = umma_idesc_bf16(128, 128), idesc16 = umma_idesc_bf16(128, 16);
  const uint32_t idesc256 = umma_idesc_bf16(128, 256), idesc144 = umma_idesc_bf16(128, kNColor);
  uint32_t stage = 0, phase = 0, ev = 0;
  for (int64_t t = 0; t < my_pairs; ++t) {
    for (int ci = 0; ci < nchunks; ++ci) {
      const uint32_t mt = meta[ci];
      const uint32_t idesc = (mt & PM_FULL) ? ((mt & PM_SMALL) ? idesc144 : idesc256)
                                            : ((mt & PM_SMALL) ? idesc16 : idesc128);
      const uint32_t a_lo = a_lo0 + (mt & 7u) * (kABlockBytes >> 4);
      const uint32_t b_lo = b_lo0 + stage * (kPairSlotBytes >> 4);
      const uint32_t d0 = tmem + ((mt & PM_HALF) << 4);  // + 128 columns for the second N half
      const uint32_t acc = (mt & PM_OVERWRITE) ? 0u : 1u;
      if (mt & PM_W0) ev ^= 1;  // first chunk of a layer: the layer's phase parity
      const uint32_t par = ev ^ 1;
      mbar_wait(bars + PairSmem::full + 8 * stage, phase);
#pragma unroll
      for (int X = 0; X < 2; ++X) {
        if (mt & PM_W0) mbar_wait(bars + PairSmem::a_ready0 + 8 * X, par);
        if (mt & PM_W1) mbar_wait(bars + PairSmem::a_ready1 + 8 * X, par);
        if (mt & PM_WD) mbar_wait(bars + PairSmem::drained1 + 8 * X, par);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ax = a_lo + X * (kPairTileBytes >> 4);
          umma_bf16_lohi(d0 + X * 256, ax, b_lo, desc_hi, idesc, acc);
          umma_bf16_lohi(d0 + X * 256, ax + 2, b_lo + 2, desc_hi, idesc, 1u);
          umma_bf16_lohi(d0 + X * 256, ax + 4, b_lo + 4, desc_hi, idesc, 1u);
          umma_bf16_lohi(d0 + X * 256, ax + 6, b_lo + 6, desc_hi, idesc, 1u);
          if (mt & PM_C0) umma_commit(bars + PairSmem::acc0 + 8 * X);
          if (mt & PM_C1) umma_commit(bars + PairSmem::acc1 + 8 * X);
          if (X == 1) umma_commit(bars + PairSmem::empty + 8 * stage);  // slot free once both tiles' MMAs retire
        }
        __syncwarp();
      }
      if (++stage == kPairStages) { stage = 0; phase ^= 1; }
    }
  }
}

}  // namespace lnrf
