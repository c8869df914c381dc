// fp32-accurate GEMMs on the tcgen05 tensor cores ("split fp16": every fp32 operand is the sum of two
// fp16 numbers, three MMAs per product, fp32 accumulation in TMEM) -- the building block of the 1e-5
// paths (fp32 NeRF, Ref-NeRF, Instant-NGP Ref-NeRF).  Same calling convention as the FFMA GEMMs of
// sgemm.cuh, which they replace for every shape with N <= 256 and K <= 320; see gemm_tc.cu.
#pragma once
#include "lnrf_common.cuh"

namespace lnrf {

// epilogues of tcg_rows (numbering follows sgemm.cuh's Epi)
constexpr int TCG_BIAS_RELU = 0;  // C = relu(acc + bias[n])
constexpr int TCG_BIAS = 1;       // C = acc + bias[n]
constexpr int TCG_MASK = 2;       // C = acc * (aux[m,n] > 0)
constexpr int TCG_RANK1 = 3;      // C = acc + r1s[m] * r1w[n]
constexpr int TCG_STORE = 5;      // C = acc
constexpr int TCG_MASKBITS = 6;   // C = acc * bit, bits written by an earlier TCG_BIAS_RELU call with mask_out (same M, N)

// Operand range: an operand whose magnitude is far from 1 (gradients) must come with the device
// address of max|operand| (`*_amax`, a float written by the producer of that operand: every tcg_rows
// call can publish max|C| through `c_amax`, an atomic max the caller zeroes once per step); the kernels
// derive a power-of-two scale from it so that the fp16 pair keeps 22 significant bits.  nullptr = the
// operand is O(1) (activations, encodings).  Weights are scaled inside the kernels.  The two K segments of
// tcg_rows share one scale, taken from the larger of a_amax / a1_amax.
//
// C[M,N] = epi((A0[M,K0] | A1[M,K1]) @ W),  W = B[K0+K1, N] (btrans = false) or Bt[N, K]^T (btrans = true)
int tcg_rows(cudaStream_t st, int epi, bool btrans, int64_t M, int N, const float* A0, int lda0, int K0,
             const float* A1, int lda1, int K1, const float* B, int ldb, float* C, int ldc, const float* bias,
             const float* aux, int ldaux, const float* r1s, const float* r1w, const float* a_amax, const float* a1_amax,
             float* c_amax, const uint32_t* mask_in = nullptr, uint32_t* mask_out = nullptr);
// 32-bit words a [M, N] bit mask occupies (layout private to gemm_tc.cu)
inline int64_t tcg_mask_words(int64_t M, int N) { return ceil_div(M, 32) * ceil_div(N, 32) * 32; }

// C[M,N] += At[K,M]^T @ B[K,N] (K = samples, split over the grid, atomic accumulation) and, if db is not
// null, db[N] += column sums of B.
int tcg_tn_acc(cudaStream_t st, int M, int N, const float* At, int lda, const float* B, int ldb, int64_t K, float* C,
               int ldc, float* db, const float* a_amax, const float* b_amax);

// max|x| over n floats into *amax (atomic max; the caller zeroes the slot)
int tcg_amax(cudaStream_t st, const float* x, int64_t n, float* amax);

bool tcg_supported(int N, int K0, int K1);
int init_gemm_tc();

}  // namespace lnrf
