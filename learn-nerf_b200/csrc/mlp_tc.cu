// bf16 tcgen05 path of the NeRF MLP, host side: weight packing (B-operand images), the debug GEMMs
// that pin the UMMA descriptor encodings, per-device init and the forward dispatch.  The kernels
// live in mlp_tc_cta2_fwd.cu (forward), mlp_tc_cta2_bwd.cu (dX chain) and mlp_tc_bwd.cu (dW).
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace lnrf {

using namespace ptx;

// One thread per (chunk, n, 16-byte group of 8 k): writes the bf16 B-operand image in the
// SW128 K-major layout the MMA expects (see tc_common.cuh for the two chunk forms).
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ P, uint8_t* __restrict__ packed) {
  const int ci = blockIdx.y;
  if (ci == kAllChunks) {  // gather the small fp32 parameters (SmallParams image)
    float* out = reinterpret_cast<float*>(packed + kSmallOffset);
    const int total = int(sizeof(SmallParams) / 4);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      float v;
      if (i < 9 * 256) v = __ldg(P + c_nerf.b[i >> 8] + (i & 255));
      else if (i < 9 * 256 + 128) v = __ldg(P + c_nerf.b[10] + (i - 9 * 256));
      else if (i < 9 * 256 + 128 + 384) v = __ldg(P + c_nerf.w[11] + (i - 9 * 256 - 128));
      else if (i == 9 * 256 + 128 + 384) v = __ldg(P + c_nerf.b[9]);
      else v = __ldg(P + c_nerf.b[11] + (i - 9 * 256 - 128 - 384 - 1));
      out[i] = v;
    }
    return;
  }
  const ChunkInfo c = ci < kTcChunks ? c_chunks.f[ci] : c_chunks.b[ci - kTcChunks];
  const int items = c.n * 8;
  const int out_dim = c_nerf.out[c.layer];
  const float* W = P + c_nerf.w[c.layer];
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
    const int n = it >> 3, kg = it & 7;
    const int ng = c.n0 + n;  // global row index of this B-operand row
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kg * 8 + j;
      float w = 0.0f;
      if (k < c.kvalid) {
        if (c.transposed) {
          w = __ldg(W + int64_t(ng) * out_dim + (c.k0 + k));
        } else if (ng < out_dim) {
          w = __ldg(W + int64_t(c.k0 + k) * out_dim + ng);
        } else if (c.layer == 10 && ng == kHC && c.ablock < 4) {  // density column (Dense_9)
          w = __ldg(P + c_nerf.w[9] + (c.k0 + k));
        }
      }
      v[j] = w;
    }
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]);
    q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]);
    q.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(packed + c.offset + sw128_offset(n, kg * 8)) = q;
  }
}

// ---------------------------------------------------------------- debug GEMMs
// D[128,N] = A[128,K] * B[N,K]^T, one CTA, same operand layouts / descriptors /
// TMEM read-back as the fused kernel.  K multiple of 64 (<= 256), N multiple of 16.
__global__ void __launch_bounds__(128)
debug_umma_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, int N, int K,
                       float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kb = K / 64;
  uint8_t* sA = smem;               // kb blocks of 128 x 128 B
  uint8_t* sB = smem + kb * 16384;  // kb blocks of N x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K; i += 128) {
    int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + (k / 64) * 16384 + sw128_offset(r, k % 64)) = __float2bfloat16(A[i]);
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + (k / 64) * (N * 128) + sw128_offset(r, k % 64)) = __float2bfloat16(B[i]);
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 256);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    for (int b = 0; b < kb; ++b) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t ad = umma_desc_sw128_kmajor(smem_u32(sA + b * 16384) + k * 32);
        uint64_t bd = umma_desc_sw128_kmajor(smem_u32(sB + b * (N * 128)) + k * 32);
        umma_bf16(tmem, ad, bd, idesc, (b | k) ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 32 && c0 + j < N; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// D[M,N] = At[K,M]^T * Bt[K,N] with BOTH operands MN-major (the dW = act^T @ grad shape:
// K = 128 samples are the rows of the same [128 x 64] SW128 block images the forward
// writes).  M in {128, 256} (two M halves -> two TMEM column ranges), N multiple of 64.
__global__ void __launch_bounds__(128)
debug_umma_gemm_tn_kernel(const float* __restrict__ At, const float* __restrict__ Bt, int M, int N,
                          float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                     // M/64 blocks of [128 samples x 64 features]
  uint8_t* sB = smem + (M / 64) * 16384;  // N/64 blocks
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * M; i += 128) {
    int r = i / M, c = i % M;
    *reinterpret_cast<__nv_bfloat16*>(sA + (c / 64) * 16384 + sw128_offset(r, c % 64)) = __float2bfloat16(At[i]);
  }
  for (int i = tid; i < 128 * N; i += 128) {
    int r = i / N, c = i % N;
    *reinterpret_cast<__nv_bfloat16*>(sB + (c / 64) * 16384 + sw128_offset(r, c % 64)) = __float2bfloat16(Bt[i]);
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16_mn(128, N);
    for (int h = 0; h < M / 128; ++h) {
      for (int k = 0; k < 8; ++k) {  // 8 x 16 samples
        uint64_t ad = umma_desc_sw128_mnmajor(smem_u32(sA + h * 2 * 16384) + k * 2048, 16384);
        uint64_t bd = umma_desc_sw128_mnmajor(smem_u32(sB) + k * 2048, 16384);
        umma_bf16(tmem + h * N, ad, bd, idesc, k ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int h = 0; h < M / 128; ++h) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + h * N + c0, v);
      tmem_wait_ld();
      for (int j = 0; j < 32; ++j) D[(h * 128 + tid) * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static bool g_tc_ready = false;

int init_mlp_tc_bwd();  // mlp_tc_bwd.cu
int init_mlp_tc_cta2_fwd();  // mlp_tc_cta2_fwd.cu
int init_mlp_tc_cta2_bwd();  // mlp_tc_cta2_bwd.cu
int nerf_fwd_cta2(const void* packed, const float* x, const float* d, const float* rays, const float* ts, int64_t m,
                  int T, bool save, const TcStash& stash, float* dens, float* rgb, cudaStream_t st);
template <typename K>
static int set_smem(K kernel, int bytes) {
  LNRF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  LNRF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
  return LNRF_OK;
}

int init_mlp_tc() {
  int rc = upload_tc_tables();
  if (rc) return rc;
  if ((rc = set_smem(debug_umma_gemm_kernel, 200 * 1024))) return rc;
  if ((rc = set_smem(debug_umma_gemm_tn_kernel, 200 * 1024))) return rc;
  if ((rc = init_mlp_tc_bwd())) return rc;
  if ((rc = init_mlp_tc_cta2_fwd())) return rc;
  if ((rc = init_mlp_tc_cta2_bwd())) return rc;
  g_tc_ready = true;
  return LNRF_OK;
}

bool tc_ready() { return g_tc_ready; }

int64_t tc_workspace_bytes(int64_t m, bool save) {
  return save ? carve_stash(nullptr, m).bytes : 0;  // render keeps every activation on chip
}

int nerf_fwd_tc(const float* P, const void* packed, const float* x, const float* d, const float* rays,
                const float* ts, int64_t m, int T, bool save, void* ws, int64_t ws_bytes, float* dens,
                float* rgb, cudaStream_t st) {
  (void)P;
  LNRF_REQUIRE(g_tc_ready, LNRF_E_INVALID, "lnrf_nerf_mlp_fwd(bf16): call lnrf_init first");
  TcStash stash{};
  if (save) {
    LNRF_REQUIRE(ws && (uintptr_t)ws % 1024 == 0 && ws_bytes >= tc_workspace_bytes(m, true),
                 LNRF_E_WORKSPACE, "lnrf_nerf_mlp_fwd(bf16): workspace %lld < %lld bytes or not "
                 "1024-byte aligned", (long long)ws_bytes, (long long)tc_workspace_bytes(m, true));
    stash = carve_stash(ws, m);
  }
  return nerf_fwd_cta2(packed, x, d, rays, ts, m, T, save, stash, dens, rgb, st);
}

int nerf_pack_weights(const float* P, void* packed, cudaStream_t st) {
  LNRF_REQUIRE(g_tc_ready, LNRF_E_INVALID, "lnrf_nerf_pack_weights: call lnrf_init first");
  dim3 grid(8, kAllChunks + 1);
  pack_weights_kernel<<<grid, 256, 0, st>>>(P, reinterpret_cast<uint8_t*>(packed));
  LNRF_LAUNCH_CHECK("pack_weights_kernel");
  return LNRF_OK;
}

int64_t nerf_packed_bytes() { return kPackedBytes; }

}  // namespace lnrf

extern "C" {

int lnrf_debug_umma_gemm(const float* a, const float* b, int32_t N, int32_t K, float* d_out,
                         lnrf_stream_t stream) {
  LNRF_REQUIRE(lnrf::g_tc_ready, LNRF_E_INVALID, "lnrf_debug_umma_gemm: call lnrf_init first");
  LNRF_REQUIRE(a && b && d_out, LNRF_E_INVALID, "lnrf_debug_umma_gemm: null pointer");
  LNRF_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 64 && K <= 256 && K % 64 == 0,
               LNRF_E_UNSUPPORTED, "lnrf_debug_umma_gemm: N=%d K=%d", N, K);
  size_t smem = size_t(K / 64) * (16384 + N * 128) + 1024;
  lnrf::debug_umma_gemm_kernel<<<1, 128, smem, lnrf::as_stream(stream)>>>(a, b, N, K, d_out);
  LNRF_LAUNCH_CHECK("debug_umma_gemm_kernel");
  return LNRF_OK;
}

int lnrf_debug_umma_gemm_tn(const float* at, const float* bt, int32_t M, int32_t N, float* d_out,
                            lnrf_stream_t stream) {
  LNRF_REQUIRE(lnrf::g_tc_ready, LNRF_E_INVALID, "lnrf_debug_umma_gemm_tn: call lnrf_init first");
  LNRF_REQUIRE(at && bt && d_out, LNRF_E_INVALID, "lnrf_debug_umma_gemm_tn: null pointer");
  LNRF_REQUIRE((M == 128 || M == 256) && N >= 64 && N <= 256 && N % 64 == 0 && (M / 128) * N <= 512,
               LNRF_E_UNSUPPORTED, "lnrf_debug_umma_gemm_tn: M=%d N=%d", M, N);
  size_t smem = size_t(M / 64 + N / 64) * 16384 + 1024;
  lnrf::debug_umma_gemm_tn_kernel<<<1, 128, smem, lnrf::as_stream(stream)>>>(at, bt, M, N, d_out);
  LNRF_LAUNCH_CHECK("debug_umma_gemm_tn_kernel");
  return LNRF_OK;
}

}  // extern "C"
