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
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
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

// db[n] += sum_m G[m,n]; N must divide 256 (bias gradients).  Template only so that the
// definition can live in this header.
template <int UNUSED = 0>
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ G, int64_t m, int N, float* __restrict__ db) {
  const int col = threadIdx.x % N;
  const int rsub = threadIdx.x / N, rstep = blockDim.x / N;
  const int64_t rows_per_block = ceil_div(m, gridDim.x);
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, m);
  float acc = 0.0f;
  for (int64_t r = r0 + rsub; r < r1; r += rstep) acc += __ldg(G + r * N + col);
  __shared__ float s[256];
  s[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < N) {
    float t = 0.0f;
    for (int k = 0; k < rstep; ++k) t += s[k * N + threadIdx.x];
    atomicAdd(db + threadIdx.x, t);
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

}  // namespace lnrf
