// Shared definitions of the bf16 tcgen05 NeRF-MLP kernels (forward, dX chain, dW).
#pragma once
#include <cuda_bf16.h>

#include "lnrf_common.cuh"
#include "lnrf_math.cuh"
#include "nerf_layout.cuh"
#include "sm100_ptx.cuh"

namespace lnrf {

// ---------------------------------------------------------------- packed weights
// B-operand images (bf16, K-major SW128, 64 K-columns per chunk) inside one buffer:
//   forward chunks  : B[n][k] = W_layer[k0 + k][n]   (n = output unit)
//   backward chunks : B[n][k] = W_layer[n][k0 + k]   (n = input unit; dX = g @ W^T)
// Forward tensor layers (T0..T9) and their chunks:
//   T0: Dense_0 (K = x_emb block)            T1..T4: Dense_1..4 (K = 4 activation blocks)
//   T5: Dense_5 (4 blocks + x_emb)           T6..T8: Dense_6..8
//   T9: Dense_10 with Dense_9 riding as output column 128 (4 blocks + d_emb), N = 144
// Backward tensor layers (B0..B8): B0: dc @ W10[:256]^T (K = 128), B1..B8: g_l @ W_l[:256]^T
// for l = 8..1 (K = 256).
constexpr int kTcLayers = 10;
constexpr int kTcChunks = 39;
constexpr int kBwLayers = 9;
constexpr int kBwChunks = 34;
constexpr int kNColor = 144;  // 128 colour units + density column + pad to a multiple of 16
constexpr uint32_t kChunkBytes256 = 256 * 128;
constexpr uint32_t kChunkBytes144 = kNColor * 128;
constexpr int64_t kFwdPackedBytes = 34 * int64_t(kChunkBytes256) + 5 * int64_t(kChunkBytes144);
// Pair kernels (two 128-sample tiles per CTA sharing every weight chunk) stream N-half chunks:
// [128 n x 64 k] = 16 KB, so a 64 KB ring holds four of them.
//   forward  f2: per tensor layer, per K chunk, per N half (T9: 128 colour columns, then a
//                16-row chunk carrying the density column)              -> 78 chunks
//   backward b2: B0 (K = 128) and B1..B8 (K = 256), per K chunk, per N half -> 68 chunks
// A second schedule ("full N") reuses the forward table `f` ([256 n x 64 k] chunks, 32 KB,
// 2-slot ring): both accumulator halves complete together (lockstep epilogue) but every MMA is
// N = 256, which needs 25 % less shared-memory operand bandwidth per FLOP than two N = 128
// MMAs.  It is the faster one when nothing is stashed (rendering).
constexpr int kF2Chunks = 78;
constexpr int kB2Chunks = 68;
constexpr uint32_t kChunkBytes128 = 128 * 128;
constexpr uint32_t kChunkBytes16 = 16 * 128;
template <bool FULLN>
struct PairCfg {
  static constexpr uint32_t slot_bytes = FULLN ? 2 * kChunkBytes128 : kChunkBytes128;
  static constexpr int stages = FULLN ? 2 : 4;
};
constexpr int64_t kF2PackedBytes = 73 * int64_t(kChunkBytes128) + 5 * int64_t(kChunkBytes16);
constexpr int64_t kB2PackedBytes = 68 * int64_t(kChunkBytes128);
// Small fp32 parameters the epilogues read from the constant bank (biases of the ten tensor
// layers, the rgb head): gathered by the pack kernel into the tail of the packed buffer and
// copied from there into constant memory right before each launch (one stream-ordered D2D copy).
struct SmallParams {
  float b[9][256];   // Dense_0..8 biases
  float b10[128];    // colour layer bias
  float w11[128 * 3];
  float b9, b11[3];
};
constexpr int64_t kSmallBytes = (int64_t(sizeof(SmallParams)) + 255) / 256 * 256;
constexpr int64_t kSmallOffset = kFwdPackedBytes + kBwChunks * int64_t(kChunkBytes256) + kF2PackedBytes +
                                 kB2PackedBytes;
constexpr int64_t kPackedBytes = kSmallOffset + kSmallBytes;

struct ChunkInfo {
  int layer;        // Dense index providing the weights
  int k0;           // first K index (fwd: kernel row; bwd: kernel column)
  int kvalid;       // K indices that exist (rest are zero padding)
  int n;            // B-operand rows
  int n0;           // first output unit (fwd) / input unit (bwd) of this chunk's rows
  int ablock;       // A block the MMA reads: 0..3 activations, 4 = embedding block
  int tlayer;       // tensor layer index
  int transposed;   // 0 = forward form, 1 = backward form
  int last;         // pair kernels: control word of the MMA issuer (PM_* bits)
  uint32_t offset;  // byte offset inside the packed buffer
};
// control word of a pair-kernel chunk: bits 0-2 A block, then
constexpr uint32_t PM_HALF = 8;        // accumulator columns 128.. (second N half)
constexpr uint32_t PM_SMALL = 16;      // 16-row chunk (N = 16: the density column)
constexpr uint32_t PM_C0 = 32;         // commit "accumulator half 0 complete" after this chunk
constexpr uint32_t PM_OVERWRITE = 64;  // first K chunk of its half: overwrite the accumulator
constexpr uint32_t PM_W0 = 128;        // wait: A blocks 0,1 written / accumulator half 0 drained
constexpr uint32_t PM_W1 = 256;        // wait: A blocks 2,3 and the embedding block written
constexpr uint32_t PM_WD = 512;        // wait: accumulator half 1 drained
constexpr uint32_t PM_C1 = 1024;       // commit "accumulator half 1 complete"
constexpr uint32_t PM_FULL = 2048;     // full-N chunk: N = 256 (with PM_SMALL: N = 144, the colour layer)
struct ChunkTable {
  ChunkInfo f[kTcChunks];
  ChunkInfo b[kBwChunks];
  ChunkInfo f2[kF2Chunks];
  ChunkInfo b2[kB2Chunks];
};
constexpr int kAllChunks = kTcChunks + kBwChunks + kF2Chunks + kB2Chunks;

// ---------------------------------------------------------------- tile geometry
constexpr int kTcThreads = 192;               // 4 epilogue warps + producer warp + MMA warp
constexpr uint32_t kABlockBytes = 128 * 128;  // one [128 samples x 64 features] bf16 SW128 block
constexpr uint32_t kTileBytes = 4 * kABlockBytes;

// ---------------------------------------------------------------- activation stash
// Saved by the forward (save_for_backward) and consumed by the backward kernels.  Every
// tile image is the exact shared-memory byte image (SW128 blocks), so it is written and
// re-read with plain bulk copies and doubles as a ready-made UMMA operand.
struct TcStash {
  uint8_t* H[9];    // h0..h7 (post-ReLU) and z8 (raw): tiles x 64 KB
  uint8_t* XE;      // x_emb block: tiles x 16 KB
  uint8_t* DE;      // d_emb block: tiles x 16 KB
  uint8_t* C;       // colour hidden (post-ReLU), 2 blocks: tiles x 32 KB
  uint32_t* MASK;   // ReLU masks: [tile][layer 0..8][warp 4][col 256] words, bit = row in warp
                    //   layers 0..7 = h_l > 0, layer 8 = c > 0 (cols 0..127)
  uint8_t* G[9];    // backward: g0..g8 (dL/d pre-activation), tiles x 64 KB
  uint8_t* DC;      // backward: dL/d colour pre-activation, 2 blocks: tiles x 32 KB
  float* SPRE;      // backward: dL/d density pre-activation [tiles*128]
  float* DPRE;      // backward: dL/d rgb pre-activation [tiles*128, 4] (3 used)
  int64_t bytes;
};

inline TcStash carve_stash(void* base, int64_t m) {
  TcStash s{};
  const int64_t tiles = (m + 127) / 128;
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    char* r = p + off;
    off += (bytes + 1023) / 1024 * 1024;
    return r;
  };
  for (int i = 0; i < 9; ++i) s.H[i] = reinterpret_cast<uint8_t*>(take(tiles * kTileBytes));
  s.XE = reinterpret_cast<uint8_t*>(take(tiles * kABlockBytes));
  s.DE = reinterpret_cast<uint8_t*>(take(tiles * kABlockBytes));
  s.C = reinterpret_cast<uint8_t*>(take(tiles * 2 * kABlockBytes));
  s.MASK = reinterpret_cast<uint32_t*>(take(tiles * 9 * 1024 * 4));
  for (int i = 0; i < 9; ++i) s.G[i] = reinterpret_cast<uint8_t*>(take(tiles * kTileBytes));
  s.DC = reinterpret_cast<uint8_t*>(take(tiles * 2 * kABlockBytes));
  s.SPRE = reinterpret_cast<float*>(take(tiles * 128 * 4));
  s.DPRE = reinterpret_cast<float*>(take(tiles * 128 * 16));
  s.bytes = off;
  return s;
}

// arguments of the dX-chain kernels (mlp_tc_bwd.cu, mlp_tc_bwd2.cu)
struct TcBwdArgs {
  const uint8_t* packed;
  const float* P;
  const float* dens;    // forward outputs [m], [m,3]
  const float* rgb;
  const float* d_dens;  // upstream gradients [m], [m,3]
  const float* d_rgb;
  int64_t m;
  TcStash stash;
  float* G;             // flat parameter gradient (for the two head biases)
};

// ---------------------------------------------------------------- device helpers
// 16-byte store of 8 bf16 into chunk `chunk` (0..7) of row `row` of an SW128 block
__device__ __forceinline__ void store_row_chunk(uint32_t block_base, int row, int chunk, uint32_t a,
                                                uint32_t b, uint32_t c, uint32_t d) {
  ptx::st_shared_v4(block_base + row * 128 + (((chunk ^ (row & 7)) & 7) << 4), a, b, c, d);
}
// barrier among the 128 epilogue threads only (named barrier 1)
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Each translation unit owns a private copy of the constant tables (no -rdc); it must call
// upload_tc_tables() once from its init function.
static __constant__ ChunkTable c_chunks;
static __constant__ NerfLayout c_nerf;

static ChunkTable build_chunk_table() {
  ChunkTable t{};
  int n = 0;
  uint32_t off = 0;
  auto addf = [&](int layer, int k0, int kvalid, int ncols, int ablock, int tlayer) {
    t.f[n++] = ChunkInfo{layer, k0, kvalid, ncols, 0, ablock, tlayer, 0, 0, off};
    off += uint32_t(ncols) * 128u;
  };
  addf(0, 0, kXE, 256, 4, 0);
  for (int l = 1; l <= 4; ++l)
    for (int b = 0; b < 4; ++b) addf(l, b * 64, 64, 256, b, l);
  for (int b = 0; b < 4; ++b) addf(5, b * 64, 64, 256, b, 5);
  addf(5, 256, kXE, 256, 4, 5);
  for (int l = 6; l <= 8; ++l)
    for (int b = 0; b < 4; ++b) addf(l, b * 64, 64, 256, b, l);
  for (int b = 0; b < 4; ++b) addf(10, b * 64, 64, kNColor, b, 9);
  addf(10, 256, kDE, kNColor, 4, 9);
  n = 0;
  auto addb = [&](int layer, int k0, int ablock, int tlayer) {
    t.b[n++] = ChunkInfo{layer, k0, 64, 256, 0, ablock, tlayer, 1, 0, off};
    off += kChunkBytes256;
  };
  for (int b = 0; b < 2; ++b) addb(10, b * 64, b, 0);  // dc (128 colour units) -> g8
  for (int l = 8; l >= 1; --l)
    for (int b = 0; b < 4; ++b) addb(l, b * 64, b, 9 - l);
  // ---- pair-kernel tables.  Per layer the chunks are ordered so that the two N halves of the
  // accumulator complete at different times and the A blocks are consumed pairwise:
  //   h0k0 h0k1 h1k0 h1k1 | h0k2 h0k3 (h0 emb) -> half 0 complete | h1k2 h1k3 (h1 emb) -> half 1
  // `last` carries the control word of the MMA issuer (see PM_* below).
  struct Item { int k0, kvalid, ablock; };
  // full-N control words for the chunks of an existing table (`f` / `b`), natural K order
  auto full_meta = [&](ChunkInfo* tab, int& cnt, const Item* items, int ni, int n_half1) {
    bool w1 = true;
    for (int i = 0; i < ni; ++i) {
      uint32_t m = uint32_t(items[i].ablock) | PM_FULL | (n_half1 == 16 ? PM_SMALL : 0u);
      if (i == 0) m |= PM_W0 | PM_WD | PM_OVERWRITE;
      if (w1 && (items[i].ablock >= 2 || i == ni - 1)) { m |= PM_W1; w1 = false; }
      if (i == ni - 1) m |= PM_C0 | PM_C1;
      tab[cnt++].last = int(m);
    }
  };
  auto emit_layer = [&](ChunkInfo* out, int& cnt, int layer, int tlayer, const Item* items, int ni, int transposed,
                        int n_half1) {
    // items: K chunks in order; the first min(2, ni) form group "lo" (blocks 0,1), the rest group "hi"
    const int nlo = ni < 2 ? ni : 2;
    bool w0 = true, w1 = true, wd = true;  // pending waits: A half 0, A half 1 (+emb), accumulator half 1 drained
    bool ow[2] = {true, true};
    auto push = [&](int h, const Item& it, bool needs_hi, bool commit) {
      uint32_t m = uint32_t(it.ablock) | (h ? PM_HALF : 0u) | ((h && n_half1 == 16) ? PM_SMALL : 0u);
      if (w0) { m |= PM_W0; w0 = false; }
      if (needs_hi && w1) { m |= PM_W1; w1 = false; }
      if (h && wd) { m |= PM_WD; wd = false; }
      if (ow[h]) { m |= PM_OVERWRITE; ow[h] = false; }
      if (commit) m |= h ? PM_C1 : PM_C0;
      const int rows = h ? n_half1 : 128;
      out[cnt++] = ChunkInfo{layer, it.k0, it.kvalid, rows, h * 128, it.ablock, tlayer, transposed, int(m), off};
      off += uint32_t(rows) * 128u;
    };
    const bool only_lo = ni <= nlo;  // every chunk of the layer is in the "lo" group
    for (int h = 0; h < 2; ++h)
      for (int i = 0; i < nlo; ++i) {
        const bool needs_hi = items[i].ablock >= 2;  // embedding block / blocks 2,3 are written by epilogue half 1
        push(h, items[i], needs_hi, only_lo && i == nlo - 1);
      }
    for (int h = 0; h < 2; ++h)
      for (int i = nlo; i < ni; ++i) push(h, items[i], true, i == ni - 1);
    // a layer must consume one phase of each barrier even if it does not read the blocks
    if (w1) {  // e.g. B0 of the backward (K = 128): park the wait on the first chunk
      for (int i = cnt - 1; i >= 0; --i)
        if (out[i].tlayer == tlayer && (out[i].last & PM_W0)) { out[i].last |= PM_W1; break; }
    }
  };
  n = 0;
  int nf = 0;
  {
    const Item emb_x{0, kXE, 4};
    emit_layer(t.f2, n, 0, 0, &emb_x, 1, 0, 128);
    full_meta(t.f, nf, &emb_x, 1, 128);
    const Item plain[4] = {{0, 64, 0}, {64, 64, 1}, {128, 64, 2}, {192, 64, 3}};
    for (int l = 1; l <= 4; ++l) { emit_layer(t.f2, n, l, l, plain, 4, 0, 128); full_meta(t.f, nf, plain, 4, 128); }
    const Item skip[5] = {{0, 64, 0}, {64, 64, 1}, {128, 64, 2}, {192, 64, 3}, {256, kXE, 4}};
    emit_layer(t.f2, n, 5, 5, skip, 5, 0, 128);
    full_meta(t.f, nf, skip, 5, 128);
    for (int l = 6; l <= 8; ++l) { emit_layer(t.f2, n, l, l, plain, 4, 0, 128); full_meta(t.f, nf, plain, 4, 128); }
    const Item colour[5] = {{0, 64, 0}, {64, 64, 1}, {128, 64, 2}, {192, 64, 3}, {256, kDE, 4}};
    emit_layer(t.f2, n, 10, 9, colour, 5, 0, 16);
    full_meta(t.f, nf, colour, 5, 16);
  }
  n = 0;
  {
    const Item dc[2] = {{0, 64, 0}, {64, 64, 1}};
    emit_layer(t.b2, n, 10, 0, dc, 2, 1, 128);
    const Item plain[4] = {{0, 64, 0}, {64, 64, 1}, {128, 64, 2}, {192, 64, 3}};
    for (int l = 8; l >= 1; --l) emit_layer(t.b2, n, l, 9 - l, plain, 4, 1, 128);
  }
  return t;
}

static int upload_tc_tables() {
  ChunkTable t = build_chunk_table();
  LNRF_CUDA(cudaMemcpyToSymbol(c_chunks, &t, sizeof(t)));
  NerfLayout lay = kNerf;
  LNRF_CUDA(cudaMemcpyToSymbol(c_nerf, &lay, sizeof(lay)));
  return LNRF_OK;
}

}  // namespace lnrf
