// Micro-benchmark: TMEM read bandwidth (tcgen05.ld) per SM, and with a concurrent MMA stream.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../learn-nerf_b200/csrc/sm100_ptx.cuh"
using namespace lnrf::ptx;

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),
      "=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]) : "r"(taddr) : "memory");
}

// mode 0: 4 warps read 256 cols x iters with x32 loads; mode 1: 8 warps (2 per lane quadrant), each half of the columns
// mma: 1 => warp 8 thread 0 issues M128 N256 K16 MMAs continuously into columns 256..511 (garbage smem operands)
__global__ void __launch_bounds__(320, 1) k(int warps, int iters, int mma, long long* out, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 9) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  uint32_t acc = 0;
  long long t0 = clock64();
  if (warp < warps) {
    const uint32_t lane_base = tmem + (uint32_t((warp & 3) * 32) << 16);
    const int c_begin = (warps == 8) ? (warp >> 2) * 128 : 0;
    const int c_end = (warps == 8) ? c_begin + 128 : 256;
    for (int it = 0; it < iters; ++it) {
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_base + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j];
      }
    }
  } else if (warp == 8 && mma && (tid & 31) == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 256);
    const int n_mma = iters * 16;  // same nominal duration as reading 256 cols at 128 clk per 16 cols... 
    for (int i = 0; i < n_mma; ++i) {
      umma_bf16(tmem + 256, umma_desc_sw128_kmajor(smem_u32(smem)), umma_desc_sw128_kmajor(smem_u32(smem) + 16384), idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
  }
  long long t1 = clock64();
  if ((tid & 31) == 0) out[blockIdx.x * 10 + warp] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 10 * 8); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 200;
  for (int mma = 0; mma < 2; ++mma)
    for (int warps : {0, 4, 8}) {
      if (warps == 0 && !mma) continue;
      cudaMemset(out, 0, 148 * 10 * 8);
      k<<<148, 320, 64 * 1024>>>(warps, iters, mma, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[10];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double bytes = 128.0 * 256 * 4 * iters;  // whole 128x256 fp32 tile per iter
      printf("mma=%d warps=%d err=%d: warp0 %lld clk (%.1f B/clk TMEM read), mma warp %lld clk (%.1f clk per MMA)\n", mma, warps,
             (int)e, h[0], warps ? bytes / h[0] : 0.0, h[8], mma ? double(h[8]) / (iters * 16) : 0.0);
    }
  return 0;
}
