// Micro-benchmark: what does a tcgen05.commit cost?  One thread issues groups of G MMAs (M=128, N, K=16, bf16, SS)
// each followed by a commit on an mbarrier that a SECOND warp consumes (like a ring-slot release), and we measure
// clk per MMA for G = 1, 2, 4, 8, 16, no commits at all, and the commit -> mbarrier-phase latency with an idle pipe.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../learn-nerf_b200/csrc/sm100_ptx.cuh"
using namespace lnrf::ptx;
__global__ void __launch_bounds__(128, 1) k(int N, int n_mma, int G, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar[8];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar[i]), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t da = umma_desc_sw128_kmajor(smem_u32(smem)), db = umma_desc_sw128_kmajor(smem_u32(smem) + 65536);
    long long t0 = clock64();
    if (mode == 0) {          // groups of G MMAs + commit on bar[g & 3]; nobody waits (fire and forget)
      if (G == 0) {
        for (int i = 0; i < n_mma; ++i) umma_bf16(tmem, da, db, idesc, 1u);
      } else {
        uint32_t b = 0;
        for (int i = 0; i < n_mma; i += G) {
          for (int j = 0; j < G; ++j) umma_bf16(tmem, da, db, idesc, 1u);
          umma_commit(smem_u32(&bar[b]));
          b = (b + 1) & 3;
        }
      }
      umma_commit(smem_u32(&bar[4]));
      // drain: bar[4] completes exactly once
      mbar_wait(smem_u32(&bar[4]), 0);
    } else if (mode == 1) {   // idle pipe: commit -> wait round trips (latency of one commit)
      for (int i = 0; i < n_mma; ++i) {
        umma_commit(smem_u32(&bar[5]));
        mbar_wait(smem_u32(&bar[5]), i & 1);
      }
    } else {                  // G MMAs + commit, then WAIT for that commit before the next group (serialised)
      int ph = 0;
      for (int i = 0; i < n_mma; i += G) {
        for (int j = 0; j < G; ++j) umma_bf16(tmem, da, db, idesc, 1u);
        umma_commit(smem_u32(&bar[6]));
        mbar_wait(smem_u32(&bar[6]), ph);
        ph ^= 1;
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
  long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n_mma = 4096;
  for (int N : {128, 256}) {
    for (int G : {0, 16, 8, 4, 2, 1}) {
      k<<<148, 128, 200 * 1024>>>(N, n_mma, G, 0, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("N=%3d fire-and-forget commit every %2d MMAs err=%d: %.1f clk per MMA\n", N, G, (int)e, double(h) / n_mma);
    }
    for (int G : {16, 4, 1}) {
      k<<<148, 128, 200 * 1024>>>(N, n_mma, G, 2, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("N=%3d serialised: %2d MMAs + commit + wait err=%d: %.1f clk per group (%.1f per MMA)\n", N, G, (int)e,
             double(h) / (n_mma / G), double(h) / n_mma);
    }
  }
  k<<<148, 128, 200 * 1024>>>(128, n_mma, 1, 1, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("idle pipe: commit -> mbarrier wait round trip err=%d: %.1f clk\n", (int)e, double(h) / n_mma);
  return 0;
}
