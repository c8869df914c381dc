// K7/K8: Instant-NGP multiresolution hash-grid encoding (gather + trilinear blend), its
// scatter-add backward, and the small InstantNGPModel MLP heads.
// Reference: learn_nerf/instant_ngp.py:33-54 (model), :92-118 (multires), :134-208 (one level),
// :211-224 (hash_table_lookup).
#include "embed.cuh"
#include "lnrf_common.cuh"
#include "lnrf_math.cuh"
#include "sgemm.cuh"

namespace lnrf {

constexpr int kMaxLevels = 16;
constexpr int kNgpHidden = 64, kNgpDensity = 16, kNgpDE = 24;

struct GridLevels {
  int64_t offset[kMaxLevels];  // float offset of the level's table
  int grid[kMaxLevels];
  uint32_t rows[kMaxLevels];   // table rows (hashed: table_size; dense: grid^3)
  int hashed[kMaxLevels];      // grid^3 > table_size  (instant_ngp.py:178)
  int L;
  float lo[3], inv_extent_unused[3], hi[3];
  int smooth;
};

// Corner indices and trilinear weights of one (point, level); corner order is the reference's
// x-major nesting `for xo for yo for zo` (instant_ngp.py:160-176).
struct Corners {
  uint32_t idx[8];
  float w[8];
};

__device__ __forceinline__ Corners level_corners(const GridLevels& g, int l, const float x[3]) {
  const int G = g.grid[l];
  uint32_t base[3];
  float cf[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float frac = (x[a] - g.lo[a]) / (g.hi[a] - g.lo[a]);                 // :138-139
    frac = fminf(fmaxf(frac, 0.0f), 1.0f);                                // clip :138-140
    // :141-146; the reference rounds the product before the add (no FMA contraction)
    const float fi = g.smooth ? __fadd_rn(0.5f, __fmul_rn(float(G - 2), frac)) : float(G - 1) * frac;
    const float fl = fminf(floorf(fi), float(G - 2));                     // :147-150
    float c = fi - fl;                                                    // :152
    if (g.smooth) c = (c * c) * (3.0f - 2.0f * c);                        // :154
    cf[a] = c;
    base[a] = uint32_t(fl);                                               // :156
  }
  Corners out;
  int k = 0;
#pragma unroll
  for (int xo = 0; xo < 2; ++xo)
#pragma unroll
    for (int yo = 0; yo < 2; ++yo)
#pragma unroll
      for (int zo = 0; zo < 2; ++zo, ++k) {
        const uint32_t cx = base[0] + xo, cy = base[1] + yo, cz = base[2] + zo;
        // weight = prod_axis (1 + (2 cf - 1) o - cf)   (:167-174)
        const float wx = (1.0f + (2.0f * cf[0] - 1.0f) * float(xo)) - cf[0];
        const float wy = (1.0f + (2.0f * cf[1] - 1.0f) * float(yo)) - cf[1];
        const float wz = (1.0f + (2.0f * cf[2] - 1.0f) * float(zo)) - cf[2];
        out.w[k] = (wx * wy) * wz;
        if (g.hashed[l]) {  // uint32 wrap-around is intended (:219-223)
          out.idx[k] = (cx ^ (19349663u * cy) ^ (83492791u * cz)) % g.rows[l];
        } else {            // dense grid: x + G (y + G z)   (:197-199)
          out.idx[k] = cx + uint32_t(G) * (cy + uint32_t(G) * cz);
        }
      }
  return out;
}

__device__ __forceinline__ void load_point(const float* __restrict__ x, const float* __restrict__ rays,
                                           const float* __restrict__ ts, int T, int64_t s, float p[3]) {
  if (x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = __ldg(x + s * 3 + a);
  } else {
    const int64_t r = s / T;
    const float t = __ldg(ts + s);
#pragma unroll
    for (int a = 0; a < 3; ++a)
      p[a] = __fadd_rn(__ldg(rays + r * 6 + a), __fmul_rn(__ldg(rays + r * 6 + 3 + a), t));
  }
}

// One thread per (point, level), level fastest: the [m, 2L] output row is written contiguously.
__global__ void __launch_bounds__(256)
hashgrid_fwd_kernel(const float* __restrict__ tables, GridLevels g, const float* __restrict__ x,
                    const float* __restrict__ rays, const float* __restrict__ ts, int T, int64_t m,
                    float* __restrict__ enc) {
  const int64_t total = m * g.L;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t s = i / g.L;
    const int l = int(i - s * g.L);
    float p[3];
    load_point(x, rays, ts, T, s, p);
    const Corners c = level_corners(g, l, p);
    const float2* tab = reinterpret_cast<const float2*>(tables + g.offset[l]);
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // sum over the 8 corners (:206-208)
      const float2 v = __ldg(tab + c.idx[k]);
      acc.x += c.w[k] * v.x;
      acc.y += c.w[k] * v.y;
    }
    reinterpret_cast<float2*>(enc)[i] = acc;
  }
}

// Transposed gather: d_table[idx_c] += w_c * d_enc (what autodiff of :203/:224 produces).
__global__ void __launch_bounds__(256)
hashgrid_bwd_kernel(GridLevels g, const float* __restrict__ x, const float* __restrict__ rays,
                    const float* __restrict__ ts, int T, int64_t m, const float* __restrict__ d_enc,
                    float* __restrict__ d_tables) {
  const int64_t total = m * g.L;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t s = i / g.L;
    const int l = int(i - s * g.L);
    const float2 d = __ldg(reinterpret_cast<const float2*>(d_enc) + i);
    if (d.x == 0.0f && d.y == 0.0f) continue;  // masked rays / dead units contribute nothing
    float p[3];
    load_point(x, rays, ts, T, s, p);
    const Corners c = level_corners(g, l, p);
    float2* tab = reinterpret_cast<float2*>(d_tables + g.offset[l]);
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(tab + c.idx[k], make_float2(c.w[k] * d.x, c.w[k] * d.y));
  }
}

static int make_levels(GridLevels& g, const int64_t* level_offsets, const int32_t* grid_sizes,
                       const int32_t* table_sizes, int L, const float* bmin, const float* bmax,
                       int smooth, const char* who) {
  LNRF_REQUIRE(L >= 1 && L <= kMaxLevels, LNRF_E_UNSUPPORTED, "%s: L=%d (max %d)", who, L, kMaxLevels);
  LNRF_REQUIRE(level_offsets && grid_sizes && table_sizes && bmin && bmax, LNRF_E_INVALID,
               "%s: null host pointer", who);
  g.L = L;
  g.smooth = smooth;
  for (int a = 0; a < 3; ++a) {
    g.lo[a] = bmin[a];
    g.hi[a] = bmax[a];
    g.inv_extent_unused[a] = 0.f;
  }
  for (int l = 0; l < L; ++l) {
    const int64_t G = grid_sizes[l];
    LNRF_REQUIRE(G >= 2 && G <= 4096 && table_sizes[l] >= 1 && level_offsets[l] % 2 == 0,
                 LNRF_E_UNSUPPORTED, "%s: level %d grid=%lld table=%d offset=%lld", who, l, (long long)G,
                 table_sizes[l], (long long)level_offsets[l]);
    g.offset[l] = level_offsets[l];
    g.grid[l] = int(G);
    g.hashed[l] = (G * G * G > int64_t(table_sizes[l])) ? 1 : 0;
    g.rows[l] = g.hashed[l] ? uint32_t(table_sizes[l]) : uint32_t(G * G * G);
  }
  return LNRF_OK;
}

// ================================================================ InstantNGPModel heads
// Dense_0: 2L->64 relu; Dense_1: 64->16 (col 0 -> exp -> density); [d_emb(24) | out(16)] ->
// Dense_2: 40->64 relu; Dense_3: 64->64 relu; Dense_4: 64->3 tanh   (instant_ngp.py:37,46-53)
struct NgpLayout {
  int in[5], out[5];
  int64_t w[5], b[5], total;
};
static NgpLayout ngp_layout(int L) {
  NgpLayout n{};
  const int ins[5] = {2 * L, kNgpHidden, kNgpDE + kNgpDensity, kNgpHidden, kNgpHidden};
  const int outs[5] = {kNgpHidden, kNgpDensity, kNgpHidden, kNgpHidden, 3};
  int64_t off = 0;
  for (int i = 0; i < 5; ++i) {
    n.in[i] = ins[i];
    n.out[i] = outs[i];
    n.w[i] = off;
    off = align_up(off + int64_t(ins[i]) * outs[i], 4);
    n.b[i] = off;
    off = align_up(off + outs[i], 4);
  }
  n.total = off;
  return n;
}

struct NgpWs {
  float *de, *h0, *o1, *h2, *h3, *dp4, *sdens, *gA, *gB, *go1, *e0;
  int64_t bytes;
};
static NgpWs carve_ngp(void* base, int64_t m) {
  NgpWs w{};
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t floats) {
    float* r = reinterpret_cast<float*>(p + off);
    off += align_up(floats * 4, 256);
    return r;
  };
  w.de = take(m * kNgpDE);
  w.h0 = take(m * kNgpHidden);
  w.o1 = take(m * kNgpDensity);
  w.h2 = take(m * kNgpHidden);
  w.h3 = take(m * kNgpHidden);
  w.dp4 = take(m * 4);
  w.sdens = take(m);
  w.gA = take(m * kNgpHidden);
  w.gB = take(m * kNgpHidden);
  w.go1 = take(m * kNgpDensity);
  w.e0 = take(kNgpDensity);
  w.bytes = off;
  return w;
}

// density = exp(out[:, 0])  (:49) and, for the backward, the unit vector e0 used by the rank-1
// epilogue that injects dL/d out[:,0] = d_dens * density.
__global__ void __launch_bounds__(256)
ngp_density_kernel(const float* __restrict__ o1, int64_t m, float* __restrict__ dens, float* __restrict__ e0) {
  const int64_t i0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i0 < kNgpDensity) e0[i0] = i0 == 0 ? 1.0f : 0.0f;
  for (int64_t i = i0; i < m; i += int64_t(gridDim.x) * blockDim.x) dens[i] = expf(__ldg(o1 + i * kNgpDensity));
}

// rgb = tanh(h3 @ W4 + b4): warp per sample, lane owns 2 of the 64 inputs.
__global__ void __launch_bounds__(256)
ngp_rgb_fwd_kernel(const float* __restrict__ h3, const float* __restrict__ w4, const float* __restrict__ b4,
                   int64_t m, float* __restrict__ rgb) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float w[2][3];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) w[k][j] = __ldg(w4 + (lane * 2 + k) * 3 + j);
  const float bb = lane < 3 ? __ldg(b4 + lane) : 0.0f;
  for (int64_t s = warp; s < m; s += nwarps) {
    const float2 a = __ldg(reinterpret_cast<const float2*>(h3 + s * kNgpHidden) + lane);
    float o[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) o[j] = warp_sum(a.x * w[0][j] + a.y * w[1][j]);
    if (lane < 3) rgb[s * 3 + lane] = tanhf((lane == 0 ? o[0] : (lane == 1 ? o[1] : o[2])) + bb);
  }
}

// Backward of the rgb head: dp = d_rgb (1 - rgb^2); dW4 += h3^T dp; db4 += sum dp;
// g3 = (dp @ W4^T) * (h3 > 0); also sdens = d_dens * density for the density path.
__global__ void __launch_bounds__(256)
ngp_rgb_bwd_kernel(const float* __restrict__ h3, const float* __restrict__ rgb, const float* __restrict__ d_rgb,
                   const float* __restrict__ dens, const float* __restrict__ d_dens,
                   const float* __restrict__ w4, int64_t m, float* __restrict__ g3,
                   float* __restrict__ sdens, float* __restrict__ dw4, float* __restrict__ db4) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float w[2][3], gw[2][3], gb[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      w[k][j] = __ldg(w4 + (lane * 2 + k) * 3 + j);
      gw[k][j] = 0.0f;
    }
  for (int64_t s = warp; s < m; s += nwarps) {
    float dp[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float y = __ldg(rgb + s * 3 + j);
      dp[j] = __ldg(d_rgb + s * 3 + j) * (1.0f - y * y);
      gb[j] += dp[j];
    }
    if (lane == 0) sdens[s] = __ldg(d_dens + s) * __ldg(dens + s);
    const float2 a = __ldg(reinterpret_cast<const float2*>(h3 + s * kNgpHidden) + lane);
    const float av[2] = {a.x, a.y};
    float o[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float t = dp[0] * w[k][0] + dp[1] * w[k][1] + dp[2] * w[k][2];
      o[k] = av[k] > 0.0f ? t : 0.0f;
#pragma unroll
      for (int j = 0; j < 3; ++j) gw[k][j] = fmaf(av[k], dp[j], gw[k][j]);
    }
    reinterpret_cast<float2*>(g3 + s * kNgpHidden)[lane] = make_float2(o[0], o[1]);
  }
  __shared__ float s_gw[8][kNgpHidden * 3];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) s_gw[wib][(lane * 2 + k) * 3 + j] = gw[k][j];
  __syncthreads();
  for (int i = threadIdx.x; i < kNgpHidden * 3; i += blockDim.x) {
    float t = 0.0f;
    for (int ww = 0; ww < 8; ++ww) t += s_gw[ww][i];
    atomicAdd(dw4 + i, t);
  }
  if (lane == 0) {
    atomicAdd(db4 + 0, gb[0]);
    atomicAdd(db4 + 1, gb[1]);
    atomicAdd(db4 + 2, gb[2]);
  }
}

}  // namespace lnrf

extern "C" {

int lnrf_hashgrid_fwd(const float* tables, const int64_t* level_offsets_host,
                      const int32_t* grid_sizes_host, const int32_t* table_sizes_host, int32_t L,
                      const float* bbox_min_host, const float* bbox_max_host, int32_t smooth,
                      const float* x, const float* rays, const float* ts, int64_t n, int32_t T,
                      float* enc, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && T >= 1, LNRF_E_INVALID, "lnrf_hashgrid_fwd: n=%lld T=%d", (long long)n, T);
  lnrf::GridLevels g;
  int rc = lnrf::make_levels(g, level_offsets_host, grid_sizes_host, table_sizes_host, L, bbox_min_host,
                             bbox_max_host, smooth, "lnrf_hashgrid_fwd");
  if (rc) return rc;
  const int64_t m = n * T;
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(tables && enc && ((x && !rays) || (!x && rays && ts)), LNRF_E_INVALID,
               "lnrf_hashgrid_fwd: null pointer / pass either x or (rays, ts)");
  lnrf::hashgrid_fwd_kernel<<<lnrf::ew_blocks(m * L, 256), 256, 0, lnrf::as_stream(stream)>>>(
      tables, g, x, rays, ts, T, m, enc);
  LNRF_LAUNCH_CHECK("hashgrid_fwd_kernel");
  return LNRF_OK;
}

int lnrf_hashgrid_bwd(const int64_t* level_offsets_host, const int32_t* grid_sizes_host,
                      const int32_t* table_sizes_host, int32_t L, const float* bbox_min_host,
                      const float* bbox_max_host, int32_t smooth, const float* x, const float* rays,
                      const float* ts, int64_t n, int32_t T, const float* d_enc, float* d_tables,
                      lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && T >= 1, LNRF_E_INVALID, "lnrf_hashgrid_bwd: n=%lld T=%d", (long long)n, T);
  lnrf::GridLevels g;
  int rc = lnrf::make_levels(g, level_offsets_host, grid_sizes_host, table_sizes_host, L, bbox_min_host,
                             bbox_max_host, smooth, "lnrf_hashgrid_bwd");
  if (rc) return rc;
  const int64_t m = n * T;
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(d_enc && d_tables && ((x && !rays) || (!x && rays && ts)), LNRF_E_INVALID,
               "lnrf_hashgrid_bwd: null pointer / pass either x or (rays, ts)");
  LNRF_REQUIRE((uintptr_t)d_tables % 8 == 0, LNRF_E_INVALID, "lnrf_hashgrid_bwd: d_tables not 8-byte aligned");
  lnrf::hashgrid_bwd_kernel<<<lnrf::ew_blocks(m * L, 256), 256, 0, lnrf::as_stream(stream)>>>(
      g, x, rays, ts, T, m, d_enc, d_tables);
  LNRF_LAUNCH_CHECK("hashgrid_bwd_kernel");
  return LNRF_OK;
}

int64_t lnrf_ngp_mlp_param_count(int32_t L) { return lnrf::ngp_layout(L).total; }

int lnrf_ngp_mlp_param_offsets(int32_t L, int64_t* out_host) {
  LNRF_REQUIRE(out_host && L >= 1 && L <= lnrf::kMaxLevels, LNRF_E_INVALID, "lnrf_ngp_mlp_param_offsets: bad args");
  const lnrf::NgpLayout n = lnrf::ngp_layout(L);
  for (int i = 0; i < 5; ++i) {
    out_host[2 * i] = n.w[i];
    out_host[2 * i + 1] = n.b[i];
  }
  return LNRF_OK;
}

int lnrf_ngp_mlp_workspace_bytes(int64_t m, int32_t L, int64_t* bytes_out_host) {
  (void)L;
  LNRF_REQUIRE(m >= 0 && bytes_out_host, LNRF_E_INVALID, "lnrf_ngp_mlp_workspace_bytes: bad args");
  *bytes_out_host = lnrf::carve_ngp(nullptr, m).bytes;
  return LNRF_OK;
}

int lnrf_ngp_mlp_fwd(const float* params, int32_t L, const float* enc, const float* d,
                     const float* rays, int64_t n, int32_t T, void* workspace, int64_t workspace_bytes,
                     float* dens, float* rgb, lnrf_stream_t stream) {
  using namespace lnrf;
  LNRF_REQUIRE(n >= 0 && T >= 1 && L >= 1 && L <= kMaxLevels, LNRF_E_INVALID,
               "lnrf_ngp_mlp_fwd: n=%lld T=%d L=%d", (long long)n, T, L);
  LNRF_REQUIRE((2 * L) % 4 == 0, LNRF_E_UNSUPPORTED, "lnrf_ngp_mlp_fwd: 2L=%d must be a multiple of 4", 2 * L);
  const int64_t m = n * T;
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(params && enc && workspace && dens && rgb && ((d && !rays) || (!d && rays)), LNRF_E_INVALID,
               "lnrf_ngp_mlp_fwd: null pointer / pass either d or rays");
  LNRF_REQUIRE(workspace_bytes >= carve_ngp(nullptr, m).bytes, LNRF_E_WORKSPACE,
               "lnrf_ngp_mlp_fwd: workspace %lld < %lld bytes", (long long)workspace_bytes,
               (long long)carve_ngp(nullptr, m).bytes);
  const NgpLayout nl = ngp_layout(L);
  const NgpWs w = carve_ngp(workspace, m);
  cudaStream_t st = as_stream(stream);
  const float* P = params;
  // d_emb: ray mode needs no ts (which = 1 embeds the direction)
  embed_kernel<4><<<ew_blocks(m * 3 * 4, 256), 256, 0, st>>>(d, rays, nullptr, T, 1, m, w.de);  // :37
  LNRF_LAUNCH_CHECK("embed_kernel<d>");
  int rc;
  if ((rc = gemm_nn<EPI_BIAS_RELU>(st, m, kNgpHidden, enc, 2 * L, 2 * L, nullptr, 0, 0, P + nl.w[0],
                                   kNgpHidden, w.h0, kNgpHidden, P + nl.b[0]))) return rc;   // :46-47
  if ((rc = gemm_nn<EPI_BIAS>(st, m, kNgpDensity, w.h0, kNgpHidden, kNgpHidden, nullptr, 0, 0, P + nl.w[1],
                              kNgpDensity, w.o1, kNgpDensity, P + nl.b[1]))) return rc;        // :48
  ngp_density_kernel<<<ew_blocks(m, 256), 256, 0, st>>>(w.o1, m, dens, w.e0);                  // :49
  LNRF_LAUNCH_CHECK("ngp_density_kernel");
  if ((rc = gemm_nn<EPI_BIAS_RELU>(st, m, kNgpHidden, w.de, kNgpDE, kNgpDE, w.o1, kNgpDensity, kNgpDensity,
                                   P + nl.w[2], kNgpHidden, w.h2, kNgpHidden, P + nl.b[2]))) return rc;  // :50-52
  if ((rc = gemm_nn<EPI_BIAS_RELU>(st, m, kNgpHidden, w.h2, kNgpHidden, kNgpHidden, nullptr, 0, 0,
                                   P + nl.w[3], kNgpHidden, w.h3, kNgpHidden, P + nl.b[3]))) return rc;
  ngp_rgb_fwd_kernel<<<ew_blocks(m, 8), 256, 0, st>>>(w.h3, P + nl.w[4], P + nl.b[4], m, rgb);  // :53
  LNRF_LAUNCH_CHECK("ngp_rgb_fwd_kernel");
  return LNRF_OK;
}

int lnrf_ngp_mlp_bwd(const float* params, int32_t L, const float* enc, int64_t m, void* workspace,
                     int64_t workspace_bytes, const float* dens, const float* rgb, const float* d_dens,
                     const float* d_rgb, float* d_params, float* d_enc, lnrf_stream_t stream) {
  using namespace lnrf;
  LNRF_REQUIRE(m >= 0 && L >= 1 && L <= kMaxLevels && (2 * L) % 4 == 0, LNRF_E_INVALID,
               "lnrf_ngp_mlp_bwd: m=%lld L=%d", (long long)m, L);
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(params && enc && workspace && dens && rgb && d_dens && d_rgb && d_params && d_enc,
               LNRF_E_INVALID, "lnrf_ngp_mlp_bwd: null pointer");
  LNRF_REQUIRE(workspace_bytes >= carve_ngp(nullptr, m).bytes, LNRF_E_WORKSPACE,
               "lnrf_ngp_mlp_bwd: workspace too small");
  const NgpLayout nl = ngp_layout(L);
  const NgpWs w = carve_ngp(workspace, m);
  cudaStream_t st = as_stream(stream);
  const float* P = params;
  float* G = d_params;
  const unsigned cb = ew_blocks(m, 2048);
  int rc;
  ngp_rgb_bwd_kernel<<<ew_blocks(m, 8 * 16), 256, 0, st>>>(w.h3, rgb, d_rgb, dens, d_dens, P + nl.w[4], m,
                                                            w.gA, w.sdens, G + nl.w[4], G + nl.b[4]);
  LNRF_LAUNCH_CHECK("ngp_rgb_bwd_kernel");
  // Dense_3
  if ((rc = gemm_tn_acc(st, kNgpHidden, kNgpHidden, w.h2, kNgpHidden, w.gA, kNgpHidden, m, G + nl.w[3], kNgpHidden))) return rc;
  colsum_kernel<><<<cb, 256, 0, st>>>(w.gA, m, kNgpHidden, G + nl.b[3]);
  LNRF_LAUNCH_CHECK("colsum_kernel");
  if ((rc = gemm_nt<EPI_MASK>(st, m, kNgpHidden, w.gA, kNgpHidden, kNgpHidden, P + nl.w[3], kNgpHidden, w.gB,
                              kNgpHidden, w.h2, kNgpHidden))) return rc;
  // Dense_2: input [d_emb | out]
  if ((rc = gemm_tn_acc(st, kNgpDE, kNgpHidden, w.de, kNgpDE, w.gB, kNgpHidden, m, G + nl.w[2], kNgpHidden))) return rc;
  if ((rc = gemm_tn_acc(st, kNgpDensity, kNgpHidden, w.o1, kNgpDensity, w.gB, kNgpHidden, m,
                        G + nl.w[2] + int64_t(kNgpDE) * kNgpHidden, kNgpHidden))) return rc;
  colsum_kernel<><<<cb, 256, 0, st>>>(w.gB, m, kNgpHidden, G + nl.b[2]);
  LNRF_LAUNCH_CHECK("colsum_kernel");
  // d out = g2 @ W2[24:40]^T + (d_dens * density) e0      (density = exp(out[:,0]))
  if ((rc = gemm_nt<EPI_RANK1>(st, m, kNgpDensity, w.gB, kNgpHidden, kNgpHidden,
                               P + nl.w[2] + int64_t(kNgpDE) * kNgpHidden, kNgpHidden, w.go1, kNgpDensity,
                               nullptr, 0, w.sdens, w.e0))) return rc;
  // Dense_1
  if ((rc = gemm_tn_acc(st, kNgpHidden, kNgpDensity, w.h0, kNgpHidden, w.go1, kNgpDensity, m, G + nl.w[1], kNgpDensity))) return rc;
  colsum_kernel<><<<cb, 256, 0, st>>>(w.go1, m, kNgpDensity, G + nl.b[1]);
  LNRF_LAUNCH_CHECK("colsum_kernel");
  if ((rc = gemm_nt<EPI_MASK>(st, m, kNgpHidden, w.go1, kNgpDensity, kNgpDensity, P + nl.w[1], kNgpDensity, w.gA,
                              kNgpHidden, w.h0, kNgpHidden))) return rc;
  // Dense_0 and the gradient of the encoding
  if ((rc = gemm_tn_acc(st, 2 * L, kNgpHidden, enc, 2 * L, w.gA, kNgpHidden, m, G + nl.w[0], kNgpHidden))) return rc;
  colsum_kernel<><<<cb, 256, 0, st>>>(w.gA, m, kNgpHidden, G + nl.b[0]);
  LNRF_LAUNCH_CHECK("colsum_kernel");
  if ((rc = gemm_nt<EPI_STORE>(st, m, 2 * L, w.gA, kNgpHidden, kNgpHidden, P + nl.w[0], kNgpHidden, d_enc, 2 * L))) return rc;
  return LNRF_OK;
}

}  // extern "C"
