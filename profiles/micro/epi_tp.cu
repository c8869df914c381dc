// Micro-benchmark: issue throughput of the epilogue's instruction mix (per SM, 8 warps = 2 per SMSP).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
  uint32_t r; asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo) : "memory"); return r;
}
__constant__ float c_b[4096];
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, long long* out, uint32_t* sink, const float* g) {
  __shared__ __align__(16) uint32_t sm[128 * 32 * 2];
  const int tid = threadIdx.x;
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = g[tid + j * 256];
  uint32_t acc = 0;
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm) + (tid & 127) * 128;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // 16 F2FP per 32 values
#pragma unroll
      for (int j = 0; j < 32; j += 2) { uint32_t p = pack_relu(f[j], f[j + 1]); acc ^= p; f[j] = __uint_as_float(p | 0x3f000000u); }
    } else if (MODE == 1) {  // 32 FADD (const operand) + 16 F2FP
#pragma unroll
      for (int j = 0; j < 32; j += 2) { uint32_t p = pack_relu(f[j] + c_b[(it & 63) * 32 + j], f[j + 1] + c_b[(it & 63) * 32 + j + 1]); acc ^= p; f[j] = __uint_as_float(p | 0x3f000000u); }
    } else if (MODE == 2) {  // + 4 STS.128 in the swizzled row pattern
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) pk[j / 2] = pack_relu(f[j] + c_b[(it & 63) * 32 + j], f[j + 1] + c_b[(it & 63) * 32 + j + 1]);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr + ((((it * 4 + q) ^ (tid & 7)) & 7) << 4)), "r"(pk[q*4]), "r"(pk[q*4+1]), "r"(pk[q*4+2]), "r"(pk[q*4+3]) : "memory");
    } else if (MODE == 3) {  // only the 4 STS.128
#pragma unroll
      for (int q = 0; q < 4; ++q)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr + ((((it * 4 + q) ^ (tid & 7)) & 7) << 4)), "r"(acc), "r"(acc), "r"(acc), "r"(acc) : "memory");
    } else if (MODE == 4) {  // 32 FADD with constant operands only
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += c_b[(it & 63) * 32 + j];
      asm volatile("" ::: "memory");
    }
  }
  long long t1 = clock64();
  if (MODE == 4) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= __float_as_uint(f[j]);
  }
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u) sink[0] = acc + sm[tid];
}
int main() {
  long long* out; uint32_t* sink; float* g;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4); cudaMalloc(&g, 256 * 33 * 4); cudaMemset(g, 0, 256 * 33 * 4);
  const int iters = 4096;
  long long h;
  const char* names[] = {"16 F2FP.RELU", "32 FADD(c[]) + 16 F2FP", "32 FADD + 16 F2FP + 4 STS.128", "4 STS.128", "32 FADD(c[])"};
#define RUN(M) cudaMemset(out, 0, 8); k<M><<<148, 256>>>(iters, out, sink, g); { cudaError_t e = cudaDeviceSynchronize(); if (e) printf("err %s\n", cudaGetErrorString(e)); } cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); \
  printf("mode %d (%s): %.1f clk per 32-column group per warp-pair-SMSP (8 warps/SM)\n", M, names[M], double(h) / iters);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
  return 0;
}
