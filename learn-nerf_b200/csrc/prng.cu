// jax.random.uniform(key, shape) for fp32 on the device: Threefry-2x32, 20
// rounds, counters = iota(size) split into two halves as jax/_src/prng.py threefry_2x32 does
// [recalled; JAX is not installable here], then ((bits >> 9) | 0x3F800000) as float - 1.
// Bit-exact with oracle/prng_np.py; 4 B written per sample.
#include "lnrf_common.cuh"

namespace lnrf {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    if ((i & 1) == 0) {
      x0 += x1; x1 = rotl32(x1, 13); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 15); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 26); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 6); x1 ^= x0;
    } else {
      x0 += x1; x1 = rotl32(x1, 17); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 29); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 16); x1 ^= x0;
      x0 += x1; x1 = rotl32(x1, 24); x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + uint32_t(i + 1);
  }
}

__device__ __forceinline__ float bits_to_uniform(uint32_t bits) {
  return fmaxf(0.0f, __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f);
}

// pair j = (count j, count half + j); outputs land at j and half + j
__global__ void __launch_bounds__(256)
threefry_uniform_kernel(uint32_t k0, uint32_t k1, int64_t n, float* __restrict__ out) {
  const int64_t half = (n + 1) / 2;
  for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < half; j += int64_t(gridDim.x) * blockDim.x) {
    uint32_t x0 = uint32_t(j);
    uint32_t x1 = (half + j < n) ? uint32_t(half + j) : 0u;  // odd sizes are padded with a zero counter
    threefry2x32(k0, k1, x0, x1);
    out[j] = bits_to_uniform(x0);
    if (half + j < n) out[half + j] = bits_to_uniform(x1);
  }
}

// same stream with the key read from device memory (CUDA-graph replays: the key changes every step)
__global__ void __launch_bounds__(256)
threefry_uniform_dk_kernel(const uint32_t* __restrict__ key, int64_t n, float* __restrict__ out) {
  const uint32_t k0 = __ldg(key), k1 = __ldg(key + 1);
  const int64_t half = (n + 1) / 2;
  for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < half; j += int64_t(gridDim.x) * blockDim.x) {
    uint32_t x0 = uint32_t(j);
    uint32_t x1 = (half + j < n) ? uint32_t(half + j) : 0u;
    threefry2x32(k0, k1, x0, x1);
    out[j] = bits_to_uniform(x0);
    if (half + j < n) out[half + j] = bits_to_uniform(x1);
  }
}

}  // namespace lnrf

extern "C" {

int lnrf_threefry_uniform_dk(const uint32_t* key_dev, int64_t n, float* out, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && n < (int64_t(1) << 32), LNRF_E_INVALID, "lnrf_threefry_uniform_dk: n=%lld", (long long)n);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(out && key_dev, LNRF_E_INVALID, "lnrf_threefry_uniform_dk: null pointer");
  int64_t blocks = lnrf::ceil_div((n + 1) / 2, 256);
  const int64_t cap = int64_t(lnrf::sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  lnrf::threefry_uniform_dk_kernel<<<(unsigned)blocks, 256, 0, lnrf::as_stream(stream)>>>(key_dev, n, out);
  LNRF_LAUNCH_CHECK("threefry_uniform_dk_kernel");
  return LNRF_OK;
}

int lnrf_threefry_uniform(uint32_t key0, uint32_t key1, int64_t n, float* out, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && n < (int64_t(1) << 32), LNRF_E_INVALID, "lnrf_threefry_uniform: n=%lld", (long long)n);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(out, LNRF_E_INVALID, "lnrf_threefry_uniform: null pointer");
  int64_t blocks = lnrf::ceil_div((n + 1) / 2, 256);
  const int64_t cap = int64_t(lnrf::sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  lnrf::threefry_uniform_kernel<<<(unsigned)blocks, 256, 0, lnrf::as_stream(stream)>>>(key0, key1, n, out);
  LNRF_LAUNCH_CHECK("threefry_uniform_kernel");
  return LNRF_OK;
}

}  // extern "C"
