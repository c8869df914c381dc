// K10: tree_norm x2 + optax.adam + apply_updates over one flat fp32 buffer
// (learn_nerf/train.py:59,92-106).  HBM-bound: 16 B read + 12 B written per parameter.
#include "lnrf_common.cuh"
#include "lnrf_math.cuh"

namespace lnrf {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ params, const float* __restrict__ grads, float* __restrict__ m,
            float* __restrict__ v, int64_t count, float lr, float b1, float b2, float eps,
            float inv_bc1, float inv_bc2, float grad_scale, float* __restrict__ norms_out,
            const float* __restrict__ bc_dev = nullptr) {
  if (bc_dev) {  // CUDA-graph replays: the step-dependent bias corrections live in device memory
    inv_bc1 = __ldg(bc_dev);
    inv_bc2 = __ldg(bc_dev + 1);
  }
  float gsq = 0.0f, psq = 0.0f;
  const int64_t nvec = count >> 2;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  auto upd = [&](float& p, float g, float& mm, float& vv) {
    g *= grad_scale;
    gsq = fmaf(g, g, gsq);
    psq = fmaf(p, p, psq);
    mm = b1 * mm + (1.0f - b1) * g;
    vv = b2 * vv + (1.0f - b2) * g * g;
    p -= lr * ((mm * inv_bc1) / (sqrtf(vv * inv_bc2) + eps));
  };
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 p = reinterpret_cast<float4*>(params)[i];
    float4 g = __ldg(reinterpret_cast<const float4*>(grads) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    upd(p.x, g.x, mm.x, vv.x);
    upd(p.y, g.y, mm.y, vv.y);
    upd(p.z, g.z, mm.z, vv.z);
    upd(p.w, g.w, mm.w, vv.w);
    reinterpret_cast<float4*>(params)[i] = p;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = (nvec << 2) + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    float p = params[i], mm = m[i], vv = v[i];
    upd(p, grads[i], mm, vv);
    params[i] = p; m[i] = mm; v[i] = vv;
  }
  if (norms_out) {
    gsq = warp_sum(gsq);
    psq = warp_sum(psq);
    __shared__ float s[2][8];
    if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = gsq; s[1][threadIdx.x >> 5] = psq; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < 8; ++w) { a += s[0][w]; b += s[1][w]; }
      atomicAdd(norms_out + 0, a);
      atomicAdd(norms_out + 1, b);
    }
  }
}

// ---------------------------------------------------------------- fused all-reduce + Adam
// Data-parallel training exchanges one flat gradient buffer per step (train.py is single-device;
// SURVEY 8e).  Instead of ncclAllReduce followed by the optimiser, every rank reads all ranks'
// gradient buffers directly over NVLink (symmetric-memory peer pointers), sums them in rank order
// (bitwise identical on every rank, so the replicas cannot drift) and applies Adam in the same
// pass: the summed gradient never touches local HBM.  The `extra` floats behind the parameters
// (per-rank loss sums) are reduced the same way and written to `extra_out`.
constexpr int kMaxPeers = 16;
struct PeerPtrs {
  const float* g[kMaxPeers];
};

__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 r;  // system-scope relaxed load: never served from a stale L1 line
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ float ld_peer(const float* p) {
  float r;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}

__global__ void __launch_bounds__(256)
adam_peers_kernel(float* __restrict__ params, const __grid_constant__ PeerPtrs peers, int world,
                  float* __restrict__ m, float* __restrict__ v, int64_t count, int extra, float lr, float b1,
                  float b2, float eps, float inv_bc1, float inv_bc2, float grad_scale,
                  float* __restrict__ norms_out, float* __restrict__ extra_out,
                  const float* __restrict__ bc_dev) {
  if (bc_dev) {  // CUDA-graph replays: the step-dependent bias corrections live in device memory
    inv_bc1 = __ldg(bc_dev);
    inv_bc2 = __ldg(bc_dev + 1);
  }
  float gsq = 0.0f, psq = 0.0f;
  const int64_t nvec = count >> 2;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  auto upd = [&](float& p, float g, float& mm, float& vv) {
    g *= grad_scale;
    gsq = fmaf(g, g, gsq);
    psq = fmaf(p, p, psq);
    mm = b1 * mm + (1.0f - b1) * g;
    vv = b2 * vv + (1.0f - b2) * g * g;
    p -= lr * ((mm * inv_bc1) / (sqrtf(vv * inv_bc2) + eps));
  };
  for (int64_t i = tid; i < nvec; i += stride) {
    float4 g = ld_peer_v4(peers.g[0] + 4 * i);
    for (int r = 1; r < world; ++r) {
      const float4 t = ld_peer_v4(peers.g[r] + 4 * i);
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    float4 p = reinterpret_cast<float4*>(params)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    upd(p.x, g.x, mm.x, vv.x);
    upd(p.y, g.y, mm.y, vv.y);
    upd(p.z, g.z, mm.z, vv.z);
    upd(p.w, g.w, mm.w, vv.w);
    reinterpret_cast<float4*>(params)[i] = p;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = (nvec << 2) + tid; i < count + extra; i += stride) {
    float g = ld_peer(peers.g[0] + i);
    for (int r = 1; r < world; ++r) g += ld_peer(peers.g[r] + i);
    if (i < count) {
      float p = params[i], mm = m[i], vv = v[i];
      upd(p, g, mm, vv);
      params[i] = p; m[i] = mm; v[i] = vv;
    } else {
      extra_out[i - count] = g;
    }
  }
  if (norms_out) {
    gsq = warp_sum(gsq);
    psq = warp_sum(psq);
    __shared__ float s[2][8];
    if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = gsq; s[1][threadIdx.x >> 5] = psq; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < 8; ++w) { a += s[0][w]; b += s[1][w]; }
      atomicAdd(norms_out + 0, a);
      atomicAdd(norms_out + 1, b);
    }
  }
}

}  // namespace lnrf

extern "C" int lnrf_adam_step_peers(float* params, const uint64_t* peer_grads, int32_t world, float* m, float* v,
                                    int64_t count, int32_t extra, float lr, float b1, float b2, float eps,
                                    int32_t step, float grad_scale, float* norms_out, float* extra_out,
                                    const float* inv_bias_corr_dev, lnrf_stream_t stream) {
  LNRF_REQUIRE(count > 0 && step >= 1 && extra >= 0, LNRF_E_INVALID, "lnrf_adam_step_peers: count=%lld step=%d",
               (long long)count, step);
  LNRF_REQUIRE(world >= 1 && world <= lnrf::kMaxPeers, LNRF_E_UNSUPPORTED, "lnrf_adam_step_peers: world=%d (max %d)",
               world, lnrf::kMaxPeers);
  LNRF_REQUIRE(params && peer_grads && m && v && (extra == 0 || extra_out), LNRF_E_INVALID,
               "lnrf_adam_step_peers: null pointer");
  lnrf::PeerPtrs pp{};
  uintptr_t al = (uintptr_t)params | (uintptr_t)m | (uintptr_t)v;
  for (int r = 0; r < world; ++r) {  // host array of device addresses (one gradient buffer per rank)
    LNRF_REQUIRE(peer_grads[r] != 0, LNRF_E_INVALID, "lnrf_adam_step_peers: null peer pointer %d", r);
    pp.g[r] = reinterpret_cast<const float*>(peer_grads[r]);
    al |= (uintptr_t)peer_grads[r];
  }
  LNRF_REQUIRE(al % 16 == 0, LNRF_E_INVALID, "lnrf_adam_step_peers: buffers must be 16-byte aligned");
  double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
  int64_t blocks = lnrf::ceil_div(lnrf::ceil_div(count, 4), 256);
  int64_t cap = int64_t(lnrf::sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  lnrf::adam_peers_kernel<<<(unsigned)blocks, 256, 0, lnrf::as_stream(stream)>>>(
      params, pp, world, m, v, count, extra, lr, b1, b2, eps, (float)(1.0 / bc1), (float)(1.0 / bc2), grad_scale,
      norms_out, extra_out, inv_bias_corr_dev);
  LNRF_LAUNCH_CHECK("adam_peers_kernel");
  return LNRF_OK;
}

extern "C" int lnrf_adam_step_dk(float* params, const float* grads, float* m, float* v, int64_t count,
                                 float lr, float b1, float b2, float eps, const float* inv_bias_corr_dev,
                                 float grad_scale, float* norms_out, lnrf_stream_t stream) {
  LNRF_REQUIRE(count >= 0, LNRF_E_INVALID, "lnrf_adam_step_dk: count=%lld", (long long)count);
  if (count == 0) return LNRF_OK;
  LNRF_REQUIRE(params && grads && m && v && inv_bias_corr_dev, LNRF_E_INVALID, "lnrf_adam_step_dk: null pointer");
  LNRF_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v) % 16 == 0,
               LNRF_E_INVALID, "lnrf_adam_step_dk: buffers must be 16-byte aligned");
  int64_t blocks = lnrf::ceil_div(lnrf::ceil_div(count, 4), 256);
  int64_t cap = int64_t(lnrf::sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  lnrf::adam_kernel<<<(unsigned)blocks, 256, 0, lnrf::as_stream(stream)>>>(
      params, grads, m, v, count, lr, b1, b2, eps, 0.0f, 0.0f, grad_scale, norms_out, inv_bias_corr_dev);
  LNRF_LAUNCH_CHECK("adam_kernel");
  return LNRF_OK;
}

extern "C" int lnrf_adam_step(float* params, const float* grads, float* m, float* v, int64_t count,
                              float lr, float b1, float b2, float eps, int32_t step, float grad_scale,
                              float* norms_out, lnrf_stream_t stream) {
  LNRF_REQUIRE(count >= 0 && step >= 1, LNRF_E_INVALID, "lnrf_adam_step: count=%lld step=%d",
               (long long)count, step);
  if (count == 0) return LNRF_OK;
  LNRF_REQUIRE(params && grads && m && v, LNRF_E_INVALID, "lnrf_adam_step: null pointer");
  LNRF_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v) % 16 == 0,
               LNRF_E_INVALID, "lnrf_adam_step: buffers must be 16-byte aligned");
  // bias corrections in double on the host, as optax does with python scalars
  double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
  int64_t blocks = lnrf::ceil_div(lnrf::ceil_div(count, 4), 256);
  int64_t cap = int64_t(lnrf::sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  lnrf::adam_kernel<<<(unsigned)blocks, 256, 0, lnrf::as_stream(stream)>>>(
      params, grads, m, v, count, lr, b1, b2, eps, (float)(1.0 / bc1), (float)(1.0 / bc2), grad_scale,
      norms_out);
  LNRF_LAUNCH_CHECK("adam_kernel");
  return LNRF_OK;
}
