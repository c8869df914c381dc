// extern "C" entry points of the NeRF MLP (K2/K6): argument checks + dispatch to
// the fp32 SIMT path (mlp_fp32.cu) or the bf16 tcgen05 path (mlp_tc.cu).
#include "lnrf_common.cuh"
#include "nerf_layout.cuh"

namespace lnrf {
int64_t fp32_workspace_bytes(int64_t m, bool save);
int nerf_fwd_fp32(const float* P, const float* x, const float* d, const float* rays, const float* ts,
                  int64_t m, int T, bool save, void* ws, int64_t ws_bytes, float* dens, float* rgb,
                  cudaStream_t st);
int nerf_bwd_fp32(const float* P, int64_t m, void* ws, int64_t ws_bytes, const float* dens,
                  const float* rgb, const float* d_dens, const float* d_rgb, float* G, cudaStream_t st);
int64_t nerf_packed_bytes();
int nerf_pack_weights(const float* P, void* packed, cudaStream_t st);
int64_t tc_workspace_bytes(int64_t m, bool save);
int nerf_fwd_tc(const float* P, const void* packed, const float* x, const float* d, const float* rays,
                const float* ts, int64_t m, int T, bool save, void* ws, int64_t ws_bytes, float* dens,
                float* rgb, cudaStream_t st);
int nerf_bwd_tc(const float* P, const void* packed, int64_t m, void* ws, int64_t ws_bytes,
                const float* dens, const float* rgb, const float* d_dens, const float* d_rgb, float* G,
                cudaStream_t st);
}  // namespace lnrf

extern "C" {

int64_t lnrf_nerf_param_count(void) {
  int64_t n = 0;
  for (int i = 0; i < lnrf::kNerfLayers; ++i)
    n += int64_t(lnrf::kNerf.in[i]) * lnrf::kNerf.out[i] + lnrf::kNerf.out[i];
  return n;
}

int64_t lnrf_nerf_param_floats(void) { return lnrf::kNerf.total; }

int lnrf_nerf_param_offsets(int64_t* out_host) {
  LNRF_REQUIRE(out_host, LNRF_E_INVALID, "lnrf_nerf_param_offsets: null pointer");
  for (int i = 0; i < lnrf::kNerfLayers; ++i) {
    out_host[2 * i] = lnrf::kNerf.w[i];
    out_host[2 * i + 1] = lnrf::kNerf.b[i];
  }
  return LNRF_OK;
}

int64_t lnrf_nerf_packed_bytes(void) { return lnrf::nerf_packed_bytes(); }

int lnrf_nerf_pack_weights(const float* params, void* packed, lnrf_stream_t stream) {
  LNRF_REQUIRE(params && packed, LNRF_E_INVALID, "lnrf_nerf_pack_weights: null pointer");
  LNRF_REQUIRE((uintptr_t)packed % 1024 == 0, LNRF_E_INVALID,
               "lnrf_nerf_pack_weights: packed must be 1024-byte aligned");
  return lnrf::nerf_pack_weights(params, packed, lnrf::as_stream(stream));
}

int lnrf_nerf_mlp_workspace_bytes(int64_t m, int32_t precision, int32_t save_for_backward,
                                  int64_t* bytes_out_host) {
  LNRF_REQUIRE(m >= 0 && bytes_out_host, LNRF_E_INVALID, "lnrf_nerf_mlp_workspace_bytes: bad args");
  if (precision == LNRF_PREC_FP32) {
    *bytes_out_host = lnrf::fp32_workspace_bytes(m, save_for_backward != 0);
  } else if (precision == LNRF_PREC_BF16) {
    *bytes_out_host = lnrf::tc_workspace_bytes(m, save_for_backward != 0);
  } else {
    LNRF_REQUIRE(false, LNRF_E_INVALID, "lnrf_nerf_mlp_workspace_bytes: precision=%d", precision);
  }
  return LNRF_OK;
}

int lnrf_nerf_mlp_fwd(const float* params, const void* packed, const float* x, const float* d,
                      const float* rays, const float* ts, int64_t n, int32_t T, int32_t precision,
                      int32_t save_for_backward, void* workspace, int64_t workspace_bytes, float* dens,
                      float* rgb, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && T >= 1, LNRF_E_INVALID, "lnrf_nerf_mlp_fwd: n=%lld T=%d", (long long)n, T);
  const int64_t m = n * T;
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(params && dens && rgb, LNRF_E_INVALID, "lnrf_nerf_mlp_fwd: null pointer");
  LNRF_REQUIRE((x && d && !rays && !ts) || (!x && !d && rays && ts), LNRF_E_INVALID,
               "lnrf_nerf_mlp_fwd: pass either (x,d) or (rays,ts)");
  LNRF_REQUIRE(m < (int64_t(1) << 31), LNRF_E_UNSUPPORTED,
               "lnrf_nerf_mlp_fwd: %lld samples per call; chunk the batch", (long long)m);
  if (precision == LNRF_PREC_FP32) {
    LNRF_REQUIRE(workspace, LNRF_E_WORKSPACE, "lnrf_nerf_mlp_fwd: null workspace");
    return lnrf::nerf_fwd_fp32(params, x, d, rays, ts, m, T, save_for_backward != 0, workspace,
                               workspace_bytes, dens, rgb, lnrf::as_stream(stream));
  }
  LNRF_REQUIRE(precision == LNRF_PREC_BF16, LNRF_E_INVALID, "lnrf_nerf_mlp_fwd: precision=%d", precision);
  LNRF_REQUIRE(packed && (uintptr_t)packed % 1024 == 0, LNRF_E_INVALID,
               "lnrf_nerf_mlp_fwd(bf16): packed weights missing or not 1024-byte aligned");
  return lnrf::nerf_fwd_tc(params, packed, x, d, rays, ts, m, T, save_for_backward != 0, workspace,
                           workspace_bytes, dens, rgb, lnrf::as_stream(stream));
}

int lnrf_nerf_mlp_bwd(const float* params, const void* packed, int64_t m, int32_t precision,
                      void* workspace, int64_t workspace_bytes, const float* dens, const float* rgb,
                      const float* d_dens, const float* d_rgb, float* d_params, lnrf_stream_t stream) {
  LNRF_REQUIRE(m >= 0, LNRF_E_INVALID, "lnrf_nerf_mlp_bwd: m=%lld", (long long)m);
  if (m == 0) return LNRF_OK;
  LNRF_REQUIRE(params && workspace && dens && rgb && d_dens && d_rgb && d_params, LNRF_E_INVALID,
               "lnrf_nerf_mlp_bwd: null pointer");
  if (precision == LNRF_PREC_FP32)
    return lnrf::nerf_bwd_fp32(params, m, workspace, workspace_bytes, dens, rgb, d_dens, d_rgb,
                               d_params, lnrf::as_stream(stream));
  LNRF_REQUIRE(precision == LNRF_PREC_BF16, LNRF_E_INVALID, "lnrf_nerf_mlp_bwd: precision=%d", precision);
  LNRF_REQUIRE(packed, LNRF_E_INVALID, "lnrf_nerf_mlp_bwd(bf16): packed weights missing");
  return lnrf::nerf_bwd_tc(params, packed, m, workspace, workspace_bytes, dens, rgb, d_dens, d_rgb,
                           d_params, lnrf::as_stream(stream));
}

}  // extern "C"
