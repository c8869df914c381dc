// fp32 SIMT GEMM building block (FFMA, 128x128x8 tiles, 8x8 per thread) used by
// the 1e-5-accurate MLP paths.  C[M,N] = epi(sum_seg A_seg[M,K_seg] * B[K,N]).
//
//  ATRANS=false: A_seg[m*lda + k]        ATRANS=true : A[k*lda + m]  (one segment, split-K)
//  BTRANS=false: B[k*ldb + n]            BTRANS=true : B[n*ldb + k]
// Requirements: every K_seg, lda, ldb multiple of 4 and 16-byte aligned bases;
// M (ATRANS) / N (!BTRANS) multiples of 4.
#pragma once
#include "lnrf_common.cuh"

namespace lnrf {

enum Epi {
  EPI_BIAS_RELU = 0,  // C = relu(acc + bias[n])
  EPI_BIAS = 1,       // C = acc + bias[n]
  EPI_MASK = 2,       // C = acc * (aux[m,n] > 0)
  EPI_RANK1 = 3,      // C = acc + r1s[m] * r1w[n]
  EPI_ATOMIC = 4,     // atomicAdd(C, acc)   (split-K partial sums)
  EPI_STORE = 5       // C = acc
};

struct GemmArgs {
  const float* A0; int lda0; int K0;   // first K segment (for ATRANS: the only one; K0 = total K)
  const float* A1; int lda1; int K1;   // optional second K segment (K1 = 0 if unused)
  const float* B; int ldb;
  float* C; int ldc;
  int M, N;
  const float* bias;                   // [N]
  const float* aux; int ldaux;         // EPI_MASK
  const float* r1s; const float* r1w;  // EPI_RANK1
  int k_per_split;                     // ATRANS split-K chunk (multiple of 8)
};

constexpr int GBM = 128, GBN = 128, GBK = 8, GPAD = 4;

template <bool ATRANS, bool BTRANS, int EPI>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(GemmArgs g) {  // <= 128 registers: two blocks per SM
  __shared__ __align__(16) float As[2][GBK][GBM + GPAD];
  __shared__ __align__(16) float Bs[2][GBK][GBN + GPAD];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int64_t m0 = int64_t(blockIdx.x) * GBM;
  const int n0 = blockIdx.y * GBN;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  // K iteration space: for !ATRANS, segments 0 and 1 concatenated; for ATRANS a split-K range
  int kbeg = 0, kend = g.K0 + g.K1;
  if (ATRANS) {
    kbeg = blockIdx.z * g.k_per_split;
    kend = min(kbeg + g.k_per_split, g.K0);
  }
  const int nk = (kend - kbeg + GBK - 1) / GBK;

  float4 ra, rb;
  auto load_tiles = [&](int kt) {
    const int k0 = kbeg + kt * GBK;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!ATRANS) {
      const int row = t >> 1, kq = (t & 1) * 4;
      const int k = k0 + kq;
      const int64_t m = m0 + row;
      if (m < g.M && k < kend) {
        if (k < g.K0) ra = __ldg(reinterpret_cast<const float4*>(g.A0 + m * g.lda0 + k));
        else ra = __ldg(reinterpret_cast<const float4*>(g.A1 + m * g.lda1 + (k - g.K0)));
      }
    } else {
      const int kk = t >> 5, mq = (t & 31) * 4;
      const int64_t k = int64_t(k0) + kk;
      const int64_t m = m0 + mq;
      if (k < kend && m < g.M) ra = __ldg(reinterpret_cast<const float4*>(g.A0 + k * g.lda0 + m));
    }
    if (!BTRANS) {
      const int kk = t >> 5, nq = (t & 31) * 4;
      const int64_t k = int64_t(k0) + kk;
      const int n = n0 + nq;
      if (k < kend && n < g.N) rb = __ldg(reinterpret_cast<const float4*>(g.B + k * g.ldb + n));
    } else {
      const int row = t >> 1, kq = (t & 1) * 4;
      const int k = k0 + kq;
      const int n = n0 + row;
      if (n < g.N && k < kend) rb = __ldg(reinterpret_cast<const float4*>(g.B + int64_t(n) * g.ldb + k));
    }
  };
  auto store_tiles = [&](int buf) {
    if (!ATRANS) {
      const int row = t >> 1, kq = (t & 1) * 4;
      As[buf][kq + 0][row] = ra.x; As[buf][kq + 1][row] = ra.y;
      As[buf][kq + 2][row] = ra.z; As[buf][kq + 3][row] = ra.w;
    } else {
      const int kk = t >> 5, mq = (t & 31) * 4;
      *reinterpret_cast<float4*>(&As[buf][kk][mq]) = ra;
    }
    if (!BTRANS) {
      const int kk = t >> 5, nq = (t & 31) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][kk][nq]) = rb;
    } else {
      const int row = t >> 1, kq = (t & 1) * 4;
      Bs[buf][kq + 0][row] = rb.x; Bs[buf][kq + 1][row] = rb.y;
      Bs[buf][kq + 2][row] = rb.z; Bs[buf][kq + 3][row] = rb.w;
    }
  };

  if (nk > 0) {
    load_tiles(0);
    store_tiles(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(kt + 1);
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
    float r1 = 0.0f;
    if (EPI == EPI_RANK1) r1 = __ldg(g.r1s + m);
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      if (n >= g.N) continue;  // N multiple of 4: the float4 is all-in or all-out
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      float* cptr = g.C + m * g.ldc + n;
      if (EPI == EPI_ATOMIC) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(cptr + j, v[j]);
        continue;
      }
      if (EPI == EPI_BIAS_RELU || EPI == EPI_BIAS) {
        float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n));
        v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        if (EPI == EPI_BIAS_RELU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
      } else if (EPI == EPI_MASK) {
        float4 a = __ldg(reinterpret_cast<const float4*>(g.aux + m * g.ldaux + n));
        v[0] = a.x > 0.f ? v[0] : 0.f; v[1] = a.y > 0.f ? v[1] : 0.f;
        v[2] = a.z > 0.f ? v[2] : 0.f; v[3] = a.w > 0.f ? v[3] : 0.f;
      } else if (EPI == EPI_RANK1) {
        float4 w = __ldg(reinterpret_cast<const float4*>(g.r1w + n));
        v[0] += r1 * w.x; v[1] += r1 * w.y; v[2] += r1 * w.z; v[3] += r1 * w.w;
      }
      *reinterpret_cast<float4*>(cptr) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// db[n] += sum_m G[m,n] (bias gradients); N a multiple of 4 with N/4 dividing 256.  128-bit
// coalesced row loads, one block per row range, a shared-memory fold and one atomic per column.
// Template only so that the definition can live in this header.
template <int UNUSED = 0>
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ G, int64_t m, int N, float* __restrict__ db) {
  const int n4 = N >> 2;                    // float4 per row
  const int c4 = threadIdx.x % n4;
  const int rsub = threadIdx.x / n4, rstep = blockDim.x / n4;
  const int64_t rows_per_block = ceil_div(m, gridDim.x);
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, m);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* Gv = reinterpret_cast<const float4*>(G);
#pragma unroll 4
  for (int64_t r = r0 + rsub; r < r1; r += rstep) {
    const float4 v = __ldg(Gv + r * n4 + c4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  __shared__ float4 s[256];
  s[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < n4) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < rstep; ++k) {
      const float4 v = s[k * n4 + threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    atomicAdd(db + threadIdx.x * 4 + 0, t.x);
    atomicAdd(db + threadIdx.x * 4 + 1, t.y);
    atomicAdd(db + threadIdx.x * 4 + 2, t.z);
    atomicAdd(db + threadIdx.x * 4 + 3, t.w);
  }
}

template <bool ATRANS, bool BTRANS, int EPI>
static int launch_sgemm(const GemmArgs& g, int splits, cudaStream_t stream) {
  dim3 grid((unsigned)ceil_div(g.M, GBM), (unsigned)ceil_div(g.N, GBN), (unsigned)splits);
  sgemm_kernel<ATRANS, BTRANS, EPI><<<grid, 256, 0, stream>>>(g);
  LNRF_LAUNCH_CHECK("sgemm_kernel");
  return LNRF_OK;
}

// C[M,N] = epi(A0[M,K0] (| A1[M,K1]) @ B[K0+K1,N])
template <int EPI>
static int gemm_nn(cudaStream_t st, int64_t M, int N, const float* A0, int lda0, int K0,
                   const float* A1, int lda1, int K1, const float* B, int ldb, float* C, int ldc,
                   const float* bias, const float* aux = nullptr, int ldaux = 0,
                   const float* r1s = nullptr, const float* r1w = nullptr) {
  GemmArgs g{A0, lda0, K0, A1, lda1, K1, B, ldb, C, ldc, (int)M, N, bias, aux, ldaux, r1s, r1w, 0};
  return launch_sgemm<false, false, EPI>(g, 1, st);
}
// C[M,N] = epi(A[M,K] @ Bt[N,K]^T)      (dX = dZ @ W^T with W stored [N(in), K(out)])
template <int EPI>
static int gemm_nt(cudaStream_t st, int64_t M, int N, const float* A, int lda, int K, const float* Bt,
                   int ldb, float* C, int ldc, const float* aux = nullptr, int ldaux = 0,
                   const float* r1s = nullptr, const float* r1w = nullptr) {
  GemmArgs g{A, lda, K, nullptr, 0, 0, Bt, ldb, C, ldc, (int)M, N, nullptr, aux, ldaux, r1s, r1w, 0};
  return launch_sgemm<false, true, EPI>(g, 1, st);
}
// C[M,N] += At[K,M]^T @ B[K,N] with split-K atomics   (dW = act^T @ dZ, K = samples)
static int gemm_tn_acc(cudaStream_t st, int M, int N, const float* At, int lda, const float* B,
                       int ldb, int64_t K, float* C, int ldc) {
  int tiles = int(ceil_div(M, GBM) * ceil_div(N, GBN));
  int target = sm_count() * 4;
  int64_t splits = target / tiles;
  if (splits < 1) splits = 1;
  int64_t kps = align_up(ceil_div(K, splits), GBK);
  if (kps < 256) kps = 256;
  splits = ceil_div(K, kps);
  GemmArgs g{At, lda, (int)K, nullptr, 0, 0, B, ldb, C, ldc, M, N, nullptr, nullptr, 0, nullptr, nullptr, (int)kps};
  return launch_sgemm<true, false, EPI_ATOMIC>(g, (int)splits, st);
}

// ---------------------------------------------------------------- small dW kernel
// C[M,N] += At[K,M]^T @ B[K,N] and (optionally) db[N] += column sums of B, for M, N <= 64
// (the Instant-NGP head layers).  One 64x64 tile per block, 4x4 accumulators per thread, the K
// range (= samples) split over the grid and reduced with atomics.  M, N, lda, ldb multiples of 4.
constexpr int SBK = 16;
// 128 threads per 64x64 tile, 8 (rows) x 4 (cols) accumulators per thread: three 128-bit shared
// loads feed 32 FFMAs (the 4x4 variant needed four: it ran at 80 % of the shared-memory wavefront
// peak), five blocks per SM (96 registers, no spills): Instant-NGP train step 32.0 -> 31.3 ms.
template <int UNUSED = 0>  // template only so that the definition can live in this header
__global__ void __launch_bounds__(128, 5)
dw_small_kernel(const float* __restrict__ At, int lda, const float* __restrict__ B, int ldb, int M, int N,
                int64_t K, int64_t k_per_block, float* __restrict__ C, int ldc, float* __restrict__ db) {
  __shared__ __align__(16) float As[2][SBK][64 + 4];
  __shared__ __align__(16) float Bs[2][SBK][64 + 4];
  __shared__ float s_db[64];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;        // output micro-tile: rows ty*8.., cols tx*4..
  const int lk = t >> 4, lq = (t & 15) * 4;  // loader: rows lk and lk + 8 of the K tile, columns lq..lq+3
  const int64_t kbeg = int64_t(blockIdx.x) * k_per_block;
  const int64_t kend = min(kbeg + k_per_block, K);
  if (t < 64) s_db[t] = 0.0f;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  float4 ra[2], rb[2], bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load = [&](int64_t k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t k = k0 + lk + 8 * h;
      ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < kend) {
        if (lq < M) ra[h] = __ldg(reinterpret_cast<const float4*>(At + k * lda + lq));
        if (lq < N) rb[h] = __ldg(reinterpret_cast<const float4*>(B + k * ldb + lq));
      }
      bsum.x += rb[h].x; bsum.y += rb[h].y; bsum.z += rb[h].z; bsum.w += rb[h].w;
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(&As[buf][lk + 8 * h][lq]) = ra[h];
      *reinterpret_cast<float4*>(&Bs[buf][lk + 8 * h][lq]) = rb[h];
    }
  };
  const int64_t nk = (kend - kbeg + SBK - 1) / SBK;
  if (nk > 0) {
    load(kbeg);
    store(0);
  }
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int buf = int(kt & 1);
    if (kt + 1 < nk) load(kbeg + (kt + 1) * SBK);
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = tx * 4 + j;
      if (n < N) atomicAdd(C + int64_t(m) * ldc + n, acc[i][j]);
    }
  }
  if (db != nullptr) {
    if (lq < N) {
      atomicAdd(&s_db[lq + 0], bsum.x);
      atomicAdd(&s_db[lq + 1], bsum.y);
      atomicAdd(&s_db[lq + 2], bsum.z);
      atomicAdd(&s_db[lq + 3], bsum.w);
    }
    __syncthreads();
    if (t < N) atomicAdd(db + t, s_db[t]);
  }
}

static int gemm_tn_small(cudaStream_t st, int M, int N, const float* At, int lda, const float* B, int ldb,
                         int64_t K, float* C, int ldc, float* db) {
  LNRF_REQUIRE(M <= 64 && N <= 64 && M % 4 == 0 && N % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0, LNRF_E_UNSUPPORTED,
               "gemm_tn_small: M=%d N=%d lda=%d ldb=%d", M, N, lda, ldb);
  int64_t blocks = int64_t(sm_count()) * 5;
  int64_t kpb = align_up(ceil_div(K, blocks), SBK);
  if (kpb < 512) kpb = 512;
  blocks = ceil_div(K, kpb);
  dw_small_kernel<><<<(unsigned)blocks, 128, 0, st>>>(At, lda, B, ldb, M, N, K, kpb, C, ldc, db);
  LNRF_LAUNCH_CHECK("dw_small_kernel");
  return LNRF_OK;
}

}  // namespace lnrf
