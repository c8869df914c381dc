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
constexpr uint32_t kChunkBytes128 = 128 * 128;  // one CTA's half of a [256 n x 64 k] chunk (CTA-pair kernels)
// Small fp32 parameters of the epilogues (biases of the ten tensor layers, the rgb head): gathered by
// the pack kernel into the tail of the packed buffer, from where the kernels stage them per layer in
// shared memory / read them with read-only loads (nothing per-model lives in constant memory).
struct SmallParams {
  float b[9][256];   // Dense_0..8 biases
  float b10[128];    // colour layer bias
  float w11[128 * 3];
  float b9, b11[3];
};
constexpr int64_t kSmallBytes = (int64_t(sizeof(SmallParams)) + 255) / 256 * 256;
constexpr int64_t kSmallOffset = kFwdPackedBytes + kBwChunks * int64_t(kChunkBytes256);
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
  uint32_t offset;  // byte offset inside the packed buffer
};
struct ChunkTable {
  ChunkInfo f[kTcChunks];
  ChunkInfo b[kBwChunks];
};
constexpr int kAllChunks = kTcChunks + kBwChunks;

// ---------------------------------------------------------------- tile geometry
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

// Each translation unit owns a private copy of the constant tables (no -rdc); it must call
// upload_tc_tables() once from its init function.
static __constant__ ChunkTable c_chunks;
static __constant__ NerfLayout c_nerf;

static ChunkTable build_chunk_table() {
  ChunkTable t{};
  int n = 0;
  uint32_t off = 0;
  auto addf = [&](int layer, int k0, int kvalid, int ncols, int ablock, int tlayer) {
    t.f[n++] = ChunkInfo{layer, k0, kvalid, ncols, 0, ablock, tlayer, 0, off};
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
    t.b[n++] = ChunkInfo{layer, k0, 64, 256, 0, ablock, tlayer, 1, off};
    off += kChunkBytes256;
  };
  for (int b = 0; b < 2; ++b) addb(10, b * 64, b, 0);  // dc (128 colour units) -> g8
  for (int l = 8; l >= 1; --l)
    for (int b = 0; b < 4; ++b) addb(l, b * 64, b, 9 - l);
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
