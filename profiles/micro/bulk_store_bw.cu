// Micro-benchmark: HBM write bandwidth reachable with cp.async.bulk shared->global stores issued by one
// thread per CTA (the stash path of the fused MLP kernels) versus plain coalesced st.global.v4.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../learn-nerf_b200/csrc/sm100_ptx.cuh"
using namespace lnrf::ptx;

// each CTA writes `iters` chunks of `chunk` bytes; at most `depth` bulk groups outstanding
__global__ void __launch_bounds__(128, 1) k_bulk(uint8_t* dst, int chunk, int iters, int depth) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint8_t* p = dst + size_t(blockIdx.x) * size_t(iters) * chunk;
    for (int it = 0; it < iters; ++it) {
      bulk_s2g(p + size_t(it) * chunk, smem_u32(smem) + (it % 4) * 32768, chunk);
      bulk_commit();
      if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else if (depth == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
    }
    bulk_wait0();
  }
}
__global__ void __launch_bounds__(256) k_st(uint4* dst, size_t n16) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += size_t(gridDim.x) * blockDim.x) dst[i] = v;
}
int main() {
  const size_t total = size_t(4) << 30;
  uint8_t* dst; cudaMalloc(&dst, total);
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_st<<<148 * 8, 256>>>(reinterpret_cast<uint4*>(dst), total / 16); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  printf("st.global.v4 fill: %.0f GB/s\n", total / ms / 1e6);
  for (int ctas : {148})
    for (int chunk : {8192, 32768})
      for (int depth : {1, 2, 4, 8}) {
        const int iters = int(total / ctas / chunk);
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0); k_bulk<<<ctas, 128, 132 * 1024>>>(dst, chunk, iters, depth); cudaEventRecord(e1);
          cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("bulk S2G ctas=%d chunk=%d depth=%d: %.0f GB/s (err %d)\n", ctas, chunk, depth,
               double(ctas) * iters * chunk / ms / 1e6, (int)cudaGetLastError());
      }
  return 0;
}
