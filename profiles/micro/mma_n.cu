// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16, SS mode) as a function of N.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../learn-nerf_b200/csrc/sm100_ptx.cuh"
using namespace lnrf::ptx;
__global__ void __launch_bounds__(128, 1) k(int N, int n_mma, int distinct, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      // distinct: rotate over 4 A blocks / 4 B chunks and 4 K offsets like the real kernel
      const uint32_t a = smem_u32(smem) + (distinct ? ((i >> 2) & 3) * 16384 + (i & 3) * 32 : 0);
      const uint32_t b = smem_u32(smem) + 65536 + (distinct ? ((i >> 2) & 3) * 16384 + (i & 3) * 32 : 0);
      umma_bf16(tmem, umma_desc_sw128_kmajor(a), umma_desc_sw128_kmajor(b), idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
  long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n_mma = 4096;
  for (int distinct = 0; distinct < 2; ++distinct)
    for (int N : {16, 64, 128, 256}) {
      k<<<148, 128, 200 * 1024>>>(N, n_mma, distinct, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("N=%3d distinct=%d err=%d: %.1f clk per MMA (M128 K16)\n", N, distinct, (int)e, double(h) / n_mma);
    }
  return 0;
}
