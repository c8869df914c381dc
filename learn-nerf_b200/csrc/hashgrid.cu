// K7/K8: Instant-NGP multiresolution hash-grid encoding (gather + trilinear blend) and its
// scatter-add backward.  The InstantNGPModel MLP heads live in ngp_mlp.cu.
// Reference: learn_nerf/instant_ngp.py:33-54 (model), :92-118 (multires), :134-208 (one level),
// :211-224 (hash_table_lookup).
#include "embed.cuh"
#include "lnrf_common.cuh"
#include "lnrf_math.cuh"

namespace lnrf {

constexpr int kMaxLevels = 16;

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

// ---------------------------------------------------------------- one thread per POINT
// Consecutive threads are consecutive samples of a ray, so at the coarse levels (16^3 .. 128^3 cells)
// most lanes of a warp sit in the same or a neighbouring cell and their gathers / scatters fall on
// the same 32-byte L2 sectors (the (point, level)-per-thread kernels above spread a warp over all the
// levels of two points: every lane a different table).  Each thread walks all LT levels and moves
// its whole [2 LT] row with 16-byte accesses.
template <int LT>
__global__ void __launch_bounds__(256)
hashgrid_fwd_pt_kernel(const float* __restrict__ tables, const __grid_constant__ GridLevels g,
                       const float* __restrict__ x, const float* __restrict__ rays, const float* __restrict__ ts, int T,
                       int64_t m, float* __restrict__ enc) {
  for (int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < m; s += int64_t(gridDim.x) * blockDim.x) {
    float p[3];
    load_point(x, rays, ts, T, s, p);
    float2 out[LT];
#pragma unroll
    for (int l = 0; l < LT; ++l) {
      const Corners c = level_corners(g, l, p);
      const float2* tab = reinterpret_cast<const float2*>(tables + g.offset[l]);
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {  // sum over the 8 corners in the reference's order (:206-208)
        const float2 v = __ldg(tab + c.idx[k]);
        acc.x += c.w[k] * v.x;
        acc.y += c.w[k] * v.y;
      }
      out[l] = acc;
    }
    float4* dst = reinterpret_cast<float4*>(enc + s * 2 * LT);
#pragma unroll
    for (int q = 0; q < LT / 2; ++q) dst[q] = make_float4(out[2 * q].x, out[2 * q].y, out[2 * q + 1].x, out[2 * q + 1].y);
  }
}

// Scatter-add.  On the DENSE levels (G^3 <= table size: 16^3 .. 64^3 cells) consecutive samples of a ray
// mostly share a cell, i.e. the same eight table rows: the warp first sums the contributions of every run
// of consecutive lanes with the same cell (segmented shuffle reduction, 16 values) and only the first lane
// of a run issues the eight atomics -- otherwise thousands of same-address float2 atomics per table row
// serialise in the L2.  Hashed levels scatter directly (their rows are unrelated).
template <int LT>
__global__ void __launch_bounds__(256)
hashgrid_bwd_pt_kernel(const __grid_constant__ GridLevels g, const float* __restrict__ x, const float* __restrict__ rays,
                       const float* __restrict__ ts, int T, int64_t m, const float* __restrict__ d_enc,
                       float* __restrict__ d_tables) {
  const int lane = threadIdx.x & 31;
  const int64_t m_pad = (m + 31) / 32 * 32;  // whole warps stay in the loop (the reduction needs all 32 lanes)
  for (int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < m_pad; s += int64_t(gridDim.x) * blockDim.x) {
    const bool in_range = s < m;
    float2 d[LT];
    bool any = false;
    if (in_range) {
      const float4* src = reinterpret_cast<const float4*>(d_enc + s * 2 * LT);
#pragma unroll
      for (int q = 0; q < LT / 2; ++q) {
        const float4 v = __ldg(src + q);
        d[2 * q] = make_float2(v.x, v.y);
        d[2 * q + 1] = make_float2(v.z, v.w);
        any |= (v.x != 0.0f) | (v.y != 0.0f) | (v.z != 0.0f) | (v.w != 0.0f);
      }
    } else {
#pragma unroll
      for (int l = 0; l < LT; ++l) d[l] = make_float2(0.f, 0.f);
    }
    if (!__any_sync(0xffffffffu, any)) continue;  // masked rays / dead units contribute nothing
    float p[3] = {0.f, 0.f, 0.f};
    if (in_range) load_point(x, rays, ts, T, s, p);
#pragma unroll
    for (int l = 0; l < LT; ++l) {
      const bool live = in_range && (d[l].x != 0.0f || d[l].y != 0.0f);
      const Corners c = level_corners(g, l, p);
      float2* tab = reinterpret_cast<float2*>(d_tables + g.offset[l]);
      if (g.hashed[l]) {
        if (live) {
#pragma unroll
          for (int k = 0; k < 8; ++k) atomicAdd(tab + c.idx[k], make_float2(c.w[k] * d[l].x, c.w[k] * d[l].y));
        }
        continue;
      }
      // dense level: runs of consecutive lanes in the same cell (idx[0] is the cell's (0,0,0) corner)
      const uint32_t cell = live ? c.idx[0] : 0xffffffffu - uint32_t(lane);  // dead lanes: singleton runs
      const uint32_t prev = __shfl_up_sync(0xffffffffu, cell, 1);
      const bool head = lane == 0 || prev != cell;
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      const int run = __popc(heads & (0xffffffffu >> (31 - lane)));  // run id: monotone over the lanes
      float v[16];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[2 * k] = live ? c.w[k] * d[l].x : 0.0f;
        v[2 * k + 1] = live ? c.w[k] * d[l].y : 0.0f;
      }
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int run2 = __shfl_down_sync(0xffffffffu, run, off);
        const bool take = (lane + off < 32) && run2 == run;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float t = __shfl_down_sync(0xffffffffu, v[j], off);
          if (take) v[j] += t;
        }
      }
      if (head && live) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(tab + c.idx[k], make_float2(v[2 * k], v[2 * k + 1]));
      }
    }
  }
}

static int launch_hash_fwd(const float* tables, const GridLevels& g, const float* x, const float* rays, const float* ts,
                           int T, int64_t m, float* enc, cudaStream_t st) {
  const bool aligned = (uintptr_t)enc % 16 == 0;
  if (g.L == 16 && aligned) {
    hashgrid_fwd_pt_kernel<16><<<ew_blocks(m, 256), 256, 0, st>>>(tables, g, x, rays, ts, T, m, enc);
  } else if (g.L == 6 && aligned) {
    hashgrid_fwd_pt_kernel<6><<<ew_blocks(m, 256), 256, 0, st>>>(tables, g, x, rays, ts, T, m, enc);
  } else {
    hashgrid_fwd_kernel<<<ew_blocks(m * g.L, 256), 256, 0, st>>>(tables, g, x, rays, ts, T, m, enc);
  }
  LNRF_LAUNCH_CHECK("hashgrid_fwd_kernel");
  return LNRF_OK;
}

static int launch_hash_bwd(const GridLevels& g, const float* x, const float* rays, const float* ts, int T, int64_t m,
                           const float* d_enc, float* d_tables, cudaStream_t st) {
  const bool aligned = (uintptr_t)d_enc % 16 == 0;
  if (g.L == 16 && aligned) {
    hashgrid_bwd_pt_kernel<16><<<ew_blocks(m, 256), 256, 0, st>>>(g, x, rays, ts, T, m, d_enc, d_tables);
  } else if (g.L == 6 && aligned) {
    hashgrid_bwd_pt_kernel<6><<<ew_blocks(m, 256), 256, 0, st>>>(g, x, rays, ts, T, m, d_enc, d_tables);
  } else {
    hashgrid_bwd_kernel<<<ew_blocks(m * g.L, 256), 256, 0, st>>>(g, x, rays, ts, T, m, d_enc, d_tables);
  }
  LNRF_LAUNCH_CHECK("hashgrid_bwd_kernel");
  return LNRF_OK;
}

// ---------------------------------------------------------------- input Jacobian (InstantNGPRefNERFModel)
// RefNERFBase differentiates the spatial block w.r.t. x (ref_nerf.py:38-43); with a hash-grid
// spatial block (instant_ngp.py:69-82) that needs d enc / d x.  Per axis a and corner c:
//   d w_c / d x_a = sign_a(c) * s_a * prod_{b != a} w_b(c),
//   s_a = [0 <= frac_raw <= 1] / (hi_a - lo_a) * (G - 2 or G - 1) * (smooth ? 6 cf (1 - cf) : 1)
// (clip -> affine -> fi - floor(fi) -> smoothstep -> cf or 1 - cf), floor / min have zero gradient.
struct CornersD {
  uint32_t idx[8];
  float dw[8][3];  // d w_c / d x_a
  float w[8];      // the trilinear weights themselves (same expression as level_corners)
};
__device__ __forceinline__ CornersD level_corners_dx(const GridLevels& g, int l, const float x[3]) {
  const int G = g.grid[l];
  uint32_t base[3];
  float cf[3], sa[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float raw = (x[a] - g.lo[a]) / (g.hi[a] - g.lo[a]);
    const float inside = (raw >= 0.0f && raw <= 1.0f) ? 1.0f : 0.0f;
    const float frac = fminf(fmaxf(raw, 0.0f), 1.0f);
    const float fi = g.smooth ? __fadd_rn(0.5f, __fmul_rn(float(G - 2), frac)) : float(G - 1) * frac;
    const float fl = fminf(floorf(fi), float(G - 2));
    float c = fi - fl;
    float dc = 1.0f;
    if (g.smooth) {
      dc = 6.0f * c * (1.0f - c);
      c = (c * c) * (3.0f - 2.0f * c);
    }
    cf[a] = c;
    sa[a] = inside / (g.hi[a] - g.lo[a]) * float(g.smooth ? G - 2 : G - 1) * dc;
    base[a] = uint32_t(fl);
  }
  CornersD out;
  int k = 0;
#pragma unroll
  for (int xo = 0; xo < 2; ++xo)
#pragma unroll
    for (int yo = 0; yo < 2; ++yo)
#pragma unroll
      for (int zo = 0; zo < 2; ++zo, ++k) {
        const uint32_t cx = base[0] + xo, cy = base[1] + yo, cz = base[2] + zo;
        const float wx = xo ? cf[0] : 1.0f - cf[0];
        const float wy = yo ? cf[1] : 1.0f - cf[1];
        const float wz = zo ? cf[2] : 1.0f - cf[2];
        out.dw[k][0] = (xo ? sa[0] : -sa[0]) * (wy * wz);
        out.dw[k][1] = (yo ? sa[1] : -sa[1]) * (wx * wz);
        out.dw[k][2] = (zo ? sa[2] : -sa[2]) * (wx * wy);
        out.w[k] = (((1.0f + (2.0f * cf[0] - 1.0f) * float(xo)) - cf[0]) * ((1.0f + (2.0f * cf[1] - 1.0f) * float(yo)) - cf[1])) *
                   ((1.0f + (2.0f * cf[2] - 1.0f) * float(zo)) - cf[2]);
        if (g.hashed[l]) out.idx[k] = (cx ^ (19349663u * cy) ^ (83492791u * cz)) % g.rows[l];
        else out.idx[k] = cx + uint32_t(G) * (cy + uint32_t(G) * cz);
      }
  return out;
}

// out[s, a] = sum_{l,f} vec[s, 2l+f] * d enc[s, 2l+f] / d x_a  (= J^T vec), [m,4] (3 used).
// One thread per point, levels in order: deterministic, no atomics.
__global__ void __launch_bounds__(256)
hashgrid_jtv_kernel(const float* __restrict__ tables, GridLevels g, const float* __restrict__ x,
                    const float* __restrict__ rays, const float* __restrict__ ts, int T, int64_t m,
                    const float* __restrict__ vec, float* __restrict__ out) {
  for (int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < m; s += int64_t(gridDim.x) * blockDim.x) {
    float p[3];
    load_point(x, rays, ts, T, s, p);
    float acc[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < g.L; ++l) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(vec + s * 2 * g.L) + l);
      const CornersD c = level_corners_dx(g, l, p);
      const float2* tab = reinterpret_cast<const float2*>(tables + g.offset[l]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 t = __ldg(tab + c.idx[k]);
        const float tv = t.x * v.x + t.y * v.y;
        acc[0] = fmaf(c.dw[k][0], tv, acc[0]);
        acc[1] = fmaf(c.dw[k][1], tv, acc[1]);
        acc[2] = fmaf(c.dw[k][2], tv, acc[2]);
      }
    }
    reinterpret_cast<float4*>(out)[s] = make_float4(acc[0], acc[1], acc[2], 0.0f);
  }
}

// Backward of out = J^T vec along u = dL/dout [m,4]:  tvec[s, 2l+f] = (J u)[s, 2l+f] = dL/dvec, and
// d_table[idx_c][f] += (u . d w_c / d x) * vec[s, 2l+f].  One thread per (point, level).
__global__ void __launch_bounds__(256)
hashgrid_jtv_bwd_kernel(const float* __restrict__ tables, GridLevels g, const float* __restrict__ x,
                        const float* __restrict__ rays, const float* __restrict__ ts, int T, int64_t m,
                        const float* __restrict__ vec, const float* __restrict__ u, float* __restrict__ tvec,
                        float* __restrict__ d_tables) {
  const int64_t total = m * g.L;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t s = i / g.L;
    const int l = int(i - s * g.L);
    const float4 uu = __ldg(reinterpret_cast<const float4*>(u) + s);
    const float2 v = __ldg(reinterpret_cast<const float2*>(vec) + i);
    float p[3];
    load_point(x, rays, ts, T, s, p);
    const CornersD c = level_corners_dx(g, l, p);
    const float2* tab = reinterpret_cast<const float2*>(tables + g.offset[l]);
    float2* dtab = reinterpret_cast<float2*>(d_tables + g.offset[l]);
    float2 acc = make_float2(0.f, 0.f);
    const bool live = (uu.x != 0.0f || uu.y != 0.0f || uu.z != 0.0f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float du = c.dw[k][0] * uu.x + c.dw[k][1] * uu.y + c.dw[k][2] * uu.z;
      const float2 t = __ldg(tab + c.idx[k]);
      acc.x = fmaf(du, t.x, acc.x);
      acc.y = fmaf(du, t.y, acc.y);
      if (live && du != 0.0f) atomicAdd(dtab + c.idx[k], make_float2(du * v.x, du * v.y));
    }
    reinterpret_cast<float2*>(tvec)[i] = acc;
  }
}

// The same, one thread per POINT (see hashgrid_bwd_pt_kernel: consecutive lanes are consecutive samples of a
// ray, the gathers share L2 sectors, and on the dense levels a run of lanes in one cell issues ONE set of eight
// atomics after a segmented shuffle reduction).  With d_enc != nullptr the first-order scatter
// d_table[idx_c] += w_c * d_enc rides on the same atomics (the Instant-NGP Ref-NeRF backward needs both: one
// pass over the tables instead of two).
template <int LT>
__global__ void __launch_bounds__(256)
hashgrid_jtv_bwd_pt_kernel(const float* __restrict__ tables, const __grid_constant__ GridLevels g,
                           const float* __restrict__ x, const float* __restrict__ rays, const float* __restrict__ ts,
                           int T, int64_t m, const float* __restrict__ vec, const float* __restrict__ u,
                           const float* __restrict__ d_enc, float* __restrict__ tvec, float* __restrict__ d_tables) {
  const int lane = threadIdx.x & 31;
  const int64_t m_pad = (m + 31) / 32 * 32;  // whole warps stay in the loop (the reduction needs all 32 lanes)
  for (int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < m_pad; s += int64_t(gridDim.x) * blockDim.x) {
    const bool in_range = s < m;
    float4 uu = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 v[LT], de[LT], out[LT];
    float p[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int l = 0; l < LT; ++l) v[l] = de[l] = make_float2(0.f, 0.f);
    if (in_range) {
      uu = __ldg(reinterpret_cast<const float4*>(u) + s);
      const float4* src = reinterpret_cast<const float4*>(vec + s * 2 * LT);
#pragma unroll
      for (int q = 0; q < LT / 2; ++q) {
        const float4 t = __ldg(src + q);
        v[2 * q] = make_float2(t.x, t.y);
        v[2 * q + 1] = make_float2(t.z, t.w);
      }
      if (d_enc != nullptr) {
        const float4* srcd = reinterpret_cast<const float4*>(d_enc + s * 2 * LT);
#pragma unroll
        for (int q = 0; q < LT / 2; ++q) {
          const float4 t = __ldg(srcd + q);
          de[2 * q] = make_float2(t.x, t.y);
          de[2 * q + 1] = make_float2(t.z, t.w);
        }
      }
      load_point(x, rays, ts, T, s, p);
    }
    const bool ulive = in_range && (uu.x != 0.0f || uu.y != 0.0f || uu.z != 0.0f);
#pragma unroll
    for (int l = 0; l < LT; ++l) {
      const CornersD c = level_corners_dx(g, l, p);
      const float2* tab = reinterpret_cast<const float2*>(tables + g.offset[l]);
      float2* dtab = reinterpret_cast<float2*>(d_tables + g.offset[l]);
      float du[8];
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        du[k] = c.dw[k][0] * uu.x + c.dw[k][1] * uu.y + c.dw[k][2] * uu.z;
        const float2 t = __ldg(tab + c.idx[k]);
        acc.x = fmaf(du[k], t.x, acc.x);
        acc.y = fmaf(du[k], t.y, acc.y);
      }
      out[l] = acc;
      const bool live = ulive || de[l].x != 0.0f || de[l].y != 0.0f;
      if (g.hashed[l]) {
        if (live) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float ax = fmaf(c.w[k], de[l].x, du[k] * v[l].x), ay = fmaf(c.w[k], de[l].y, du[k] * v[l].y);
            if (ax != 0.0f || ay != 0.0f) atomicAdd(dtab + c.idx[k], make_float2(ax, ay));
          }
        }
        continue;
      }
      if (!__any_sync(0xffffffffu, live)) continue;
      const uint32_t cell = live ? c.idx[0] : 0xffffffffu - uint32_t(lane);  // dead lanes: singleton runs
      const uint32_t prev = __shfl_up_sync(0xffffffffu, cell, 1);
      const bool head = lane == 0 || prev != cell;
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      const int run = __popc(heads & (0xffffffffu >> (31 - lane)));
      float w[16];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        w[2 * k] = live ? fmaf(c.w[k], de[l].x, du[k] * v[l].x) : 0.0f;
        w[2 * k + 1] = live ? fmaf(c.w[k], de[l].y, du[k] * v[l].y) : 0.0f;
      }
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int run2 = __shfl_down_sync(0xffffffffu, run, off);
        const bool take = (lane + off < 32) && run2 == run;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float t = __shfl_down_sync(0xffffffffu, w[j], off);
          if (take) w[j] += t;
        }
      }
      if (head && live) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(dtab + c.idx[k], make_float2(w[2 * k], w[2 * k + 1]));
      }
    }
    if (in_range) {
      float4* dst = reinterpret_cast<float4*>(tvec + s * 2 * LT);
#pragma unroll
      for (int q = 0; q < LT / 2; ++q) dst[q] = make_float4(out[2 * q].x, out[2 * q].y, out[2 * q + 1].x, out[2 * q + 1].y);
    }
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

// internal entry points used by the InstantNGPRefNERFModel path (refnerf.cu)
int hashgrid_launch(int which, const float* tables, const int64_t* level_offsets, const int32_t* grid_sizes,
                    const int32_t* table_sizes, int L, const float* bmin, const float* bmax, int smooth,
                    const float* x, const float* rays, const float* ts, int T, int64_t m, const float* in0,
                    const float* in1, float* out0, float* out1, cudaStream_t st, const float* in2) {
  GridLevels g;
  int rc = make_levels(g, level_offsets, grid_sizes, table_sizes, L, bmin, bmax, smooth, "hashgrid_launch");
  if (rc) return rc;
  if (m == 0) return LNRF_OK;
  switch (which) {
    case 0:  // enc = encode(x)
      return launch_hash_fwd(tables, g, x, rays, ts, T, m, out0, st);
    case 1:  // d_tables += scatter(d_enc = in0)
      return launch_hash_bwd(g, x, rays, ts, T, m, in0, out0, st);
    case 2:  // out0[m,4] = J^T in0
      hashgrid_jtv_kernel<<<ew_blocks(m, 256), 256, 0, st>>>(tables, g, x, rays, ts, T, m, in0, out0);
      LNRF_LAUNCH_CHECK("hashgrid_jtv_kernel");
      break;
    case 3: {  // out0 = J in1(u), d_tables(out1) += second-order scatter with vec = in0 (+ first-order scatter of
               // d_enc = in2 when given)
      const bool al = (uintptr_t)in0 % 16 == 0 && (uintptr_t)out0 % 16 == 0 && (uintptr_t)in2 % 16 == 0;
      if (L == 16 && al) {
        hashgrid_jtv_bwd_pt_kernel<16><<<ew_blocks(m, 256), 256, 0, st>>>(tables, g, x, rays, ts, T, m, in0, in1, in2, out0, out1);
        LNRF_LAUNCH_CHECK("hashgrid_jtv_bwd_pt_kernel");
        break;
      }
      if (L == 6 && al) {
        hashgrid_jtv_bwd_pt_kernel<6><<<ew_blocks(m, 256), 256, 0, st>>>(tables, g, x, rays, ts, T, m, in0, in1, in2, out0, out1);
        LNRF_LAUNCH_CHECK("hashgrid_jtv_bwd_pt_kernel");
        break;
      }
      if (in2 != nullptr) {
        const int rc2 = launch_hash_bwd(g, x, rays, ts, T, m, in2, out1, st);
        if (rc2) return rc2;
      }
      hashgrid_jtv_bwd_kernel<<<ew_blocks(m * L, 256), 256, 0, st>>>(tables, g, x, rays, ts, T, m, in0, in1, out0,
                                                                       out1);
      LNRF_LAUNCH_CHECK("hashgrid_jtv_bwd_kernel");
      break;
    }
    default:
      return LNRF_E_INVALID;
  }
  return LNRF_OK;
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
  return lnrf::launch_hash_fwd(tables, g, x, rays, ts, T, m, enc, lnrf::as_stream(stream));
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
  return lnrf::launch_hash_bwd(g, x, rays, ts, T, m, d_enc, d_tables, lnrf::as_stream(stream));
}

}  // extern "C"
