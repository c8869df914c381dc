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

// ---------------------------------------------------------------- one call per NeRFRenderer.render_rays
// render.py:39-91 for two NeRFModels: t_range + stratified coarse sampling (K1), coarse MLP (K2), compositing (K3),
// inverse-CDF fine sampling (K4), fine MLP, compositing -- the same six launches the Python mirror issues, in
// one C call for a binding that wants a single custom call per seam.  The workspace holds the per-ray / per-sample
// intermediates and the MLP workspace of the larger level.
namespace {
struct RenderWs {
  float *t_min, *t_max, *ts_c, *dens_c, *rgb_c, *ts_f, *dens_f, *rgb_f;
  uint8_t* mask;
  void* mlp;
  int64_t mlp_bytes, bytes;
};
int carve_render(void* base, int64_t n, int Tc, int Tf, int precision, RenderWs* w) {
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    char* r = p + off;
    off += (bytes + 1023) / 1024 * 1024;
    return r;
  };
  const int T2 = Tc + Tf;
  w->t_min = reinterpret_cast<float*>(take(n * 4));
  w->t_max = reinterpret_cast<float*>(take(n * 4));
  w->mask = reinterpret_cast<uint8_t*>(take(n));
  w->ts_c = reinterpret_cast<float*>(take(n * Tc * 4));
  w->dens_c = reinterpret_cast<float*>(take(n * Tc * 4));
  w->rgb_c = reinterpret_cast<float*>(take(n * Tc * 12));
  w->ts_f = reinterpret_cast<float*>(take(n * T2 * 4));
  w->dens_f = reinterpret_cast<float*>(take(n * T2 * 4));
  w->rgb_f = reinterpret_cast<float*>(take(n * T2 * 12));
  int64_t bc = 0, bf = 0;
  int rc = lnrf_nerf_mlp_workspace_bytes(n * Tc, precision, 0, &bc);
  if (rc) return rc;
  if ((rc = lnrf_nerf_mlp_workspace_bytes(n * T2, precision, 0, &bf))) return rc;
  w->mlp_bytes = bc > bf ? bc : bf;
  w->mlp = take(w->mlp_bytes);
  w->bytes = off;
  return LNRF_OK;
}
}  // namespace

int lnrf_nerf_render_workspace_bytes(int64_t n, int32_t Tc, int32_t Tf, int32_t precision, int64_t* bytes_out_host) {
  LNRF_REQUIRE(n >= 0 && Tc >= 1 && Tf >= 1 && bytes_out_host, LNRF_E_INVALID, "lnrf_nerf_render_workspace_bytes: bad args");
  RenderWs w{};
  const int rc = carve_render(nullptr, n, Tc, Tf, precision, &w);
  if (rc) return rc;
  *bytes_out_host = w.bytes;
  return LNRF_OK;
}

int lnrf_nerf_render_rays(const float* rays, const float* bbox_min_host, const float* bbox_max_host, float min_t_range,
                          const float* u_coarse, const float* u_fine, const float* coarse_params,
                          const void* coarse_packed, const float* fine_params, const void* fine_packed,
                          int32_t precision, const float* background, int64_t n, int32_t Tc, int32_t Tf,
                          void* workspace, int64_t workspace_bytes, float* coarse_outputs, float* fine_outputs,
                          float* fine_alphas, float* fine_coords, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && Tc >= 1 && Tf >= 1, LNRF_E_INVALID, "lnrf_nerf_render_rays: n=%lld Tc=%d Tf=%d", (long long)n, Tc, Tf);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(rays && bbox_min_host && bbox_max_host && u_coarse && u_fine && coarse_params && fine_params && background &&
                   workspace && coarse_outputs && fine_outputs,
               LNRF_E_INVALID, "lnrf_nerf_render_rays: null pointer");
  LNRF_REQUIRE((uintptr_t)workspace % 1024 == 0, LNRF_E_WORKSPACE, "lnrf_nerf_render_rays: workspace not 1024-byte aligned");
  RenderWs w{};
  int rc = carve_render(workspace, n, Tc, Tf, precision, &w);
  if (rc) return rc;
  LNRF_REQUIRE(workspace_bytes >= w.bytes, LNRF_E_WORKSPACE, "lnrf_nerf_render_rays: workspace %lld < %lld bytes",
               (long long)workspace_bytes, (long long)w.bytes);
  const int T2 = Tc + Tf;
  // coarse level (render.py:53-66)
  if ((rc = lnrf_sample_coarse(rays, n, bbox_min_host, bbox_max_host, min_t_range, 1e-8f, u_coarse, Tc, w.t_min, w.t_max,
                               w.mask, w.ts_c, stream))) return rc;
  if ((rc = lnrf_nerf_mlp_fwd(coarse_params, coarse_packed, nullptr, nullptr, rays, w.ts_c, n, Tc, precision, 0, w.mlp,
                              w.mlp_bytes, w.dens_c, w.rgb_c, stream))) return rc;
  if ((rc = lnrf_composite_fwd(rays, w.ts_c, w.t_min, w.t_max, w.mask, w.dens_c, w.rgb_c, background, n, Tc,
                               coarse_outputs, nullptr, nullptr, stream))) return rc;
  // fine level (:68-84): inverse-CDF samples from the coarse densities (stop-gradient), merged with the coarse ones
  if ((rc = lnrf_sample_fine(w.ts_c, w.dens_c, w.t_min, w.t_max, u_fine, n, Tc, Tf, 1e-8f, w.ts_f, nullptr, nullptr,
                             stream))) return rc;
  if ((rc = lnrf_nerf_mlp_fwd(fine_params, fine_packed, nullptr, nullptr, rays, w.ts_f, n, T2, precision, 0, w.mlp,
                              w.mlp_bytes, w.dens_f, w.rgb_f, stream))) return rc;
  return lnrf_composite_fwd(rays, w.ts_f, w.t_min, w.t_max, w.mask, w.dens_f, w.rgb_f, background, n, T2, fine_outputs,
                            fine_alphas, fine_coords, stream);
}

// ---------------------------------------------------------------- one call per TrainLoop.step_fn
// train.py:78-112 (step_fn) around train.py:114-151 (losses) for two NeRFModels on one device: both levels rendered
// with the activation stash, MSE losses, compositing and MLP backward of both levels, tree norms + Adam -- the 19
// launches of the Python mirror's step in one C call.  Parameters / gradients / Adam moments are flat buffers
// [coarse model | fine model | background(3) + 1 pad].
namespace {
struct TrainWs {
  float *rays, *t_min, *t_max, *ts_c, *dens_c, *rgb_c, *ts_f, *dens_f, *rgb_f, *out_c, *out_f, *d_out, *d_dens, *d_rgb;
  uint8_t* mask;
  void *mlp_c, *mlp_f;
  int64_t mlp_c_bytes, mlp_f_bytes, bytes;
};
int carve_train(void* base, int64_t n, int Tc, int Tf, int precision, TrainWs* w) {
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    char* r = p + off;
    off += (bytes + 1023) / 1024 * 1024;
    return r;
  };
  const int T2 = Tc + Tf;
  w->rays = reinterpret_cast<float*>(take(n * 24));
  w->t_min = reinterpret_cast<float*>(take(n * 4));
  w->t_max = reinterpret_cast<float*>(take(n * 4));
  w->mask = reinterpret_cast<uint8_t*>(take(n));
  w->ts_c = reinterpret_cast<float*>(take(n * Tc * 4));
  w->dens_c = reinterpret_cast<float*>(take(n * Tc * 4));
  w->rgb_c = reinterpret_cast<float*>(take(n * Tc * 12));
  w->ts_f = reinterpret_cast<float*>(take(n * T2 * 4));
  w->dens_f = reinterpret_cast<float*>(take(n * T2 * 4));
  w->rgb_f = reinterpret_cast<float*>(take(n * T2 * 12));
  w->out_c = reinterpret_cast<float*>(take(n * 12));
  w->out_f = reinterpret_cast<float*>(take(n * 12));
  w->d_out = reinterpret_cast<float*>(take(n * 12));
  w->d_dens = reinterpret_cast<float*>(take(n * T2 * 4));
  w->d_rgb = reinterpret_cast<float*>(take(n * T2 * 12));
  int rc = lnrf_nerf_mlp_workspace_bytes(n * Tc, precision, 1, &w->mlp_c_bytes);
  if (rc) return rc;
  if ((rc = lnrf_nerf_mlp_workspace_bytes(n * T2, precision, 1, &w->mlp_f_bytes))) return rc;
  w->mlp_c = take(w->mlp_c_bytes);
  w->mlp_f = take(w->mlp_f_bytes);
  w->bytes = off;
  return LNRF_OK;
}
}  // namespace

int lnrf_nerf_train_workspace_bytes(int64_t n, int32_t Tc, int32_t Tf, int32_t precision, int64_t* bytes_out_host) {
  LNRF_REQUIRE(n >= 0 && Tc >= 1 && Tf >= 1 && bytes_out_host, LNRF_E_INVALID, "lnrf_nerf_train_workspace_bytes: bad args");
  TrainWs w{};
  const int rc = carve_train(nullptr, n, Tc, Tf, precision, &w);
  if (rc) return rc;
  *bytes_out_host = w.bytes;
  return LNRF_OK;
}

int lnrf_nerf_train_step(const float* batch, const float* bbox_min_host, const float* bbox_max_host, float min_t_range,
                         const float* u_coarse, const float* u_fine, float* params, float* adam_m, float* adam_v,
                         float* grads, void* coarse_packed, void* fine_packed, int32_t precision, int64_t n, int32_t Tc,
                         int32_t Tf, float lr, float b1, float b2, float eps, int32_t step, void* workspace,
                         int64_t workspace_bytes, float* scalars_out, lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 1 && Tc >= 1 && Tf >= 1 && step >= 1, LNRF_E_INVALID, "lnrf_nerf_train_step: n=%lld Tc=%d Tf=%d step=%d",
               (long long)n, Tc, Tf, step);
  LNRF_REQUIRE(batch && bbox_min_host && bbox_max_host && u_coarse && u_fine && params && adam_m && adam_v && grads &&
                   workspace && scalars_out,
               LNRF_E_INVALID, "lnrf_nerf_train_step: null pointer");
  LNRF_REQUIRE(precision == LNRF_PREC_FP32 || (coarse_packed && fine_packed), LNRF_E_INVALID,
               "lnrf_nerf_train_step(bf16): packed-weight buffers missing");
  LNRF_REQUIRE((uintptr_t)workspace % 1024 == 0, LNRF_E_WORKSPACE, "lnrf_nerf_train_step: workspace not 1024-byte aligned");
  TrainWs w{};
  int rc = carve_train(workspace, n, Tc, Tf, precision, &w);
  if (rc) return rc;
  LNRF_REQUIRE(workspace_bytes >= w.bytes, LNRF_E_WORKSPACE, "lnrf_nerf_train_step: workspace %lld < %lld bytes",
               (long long)workspace_bytes, (long long)w.bytes);
  cudaStream_t st = lnrf::as_stream(stream);
  const int64_t np = lnrf::kNerf.total, count = 2 * np + 4;
  float* p_c = params;
  float* p_f = params + np;
  float* bg = params + 2 * np;
  const int T2 = Tc + Tf;
  LNRF_CUDA(cudaMemsetAsync(grads, 0, count * sizeof(float), st));
  LNRF_CUDA(cudaMemsetAsync(scalars_out, 0, 4 * sizeof(float), st));
  // rays = batch[:, :2] (train.py:134): rows of 6 floats out of rows of 9
  LNRF_CUDA(cudaMemcpy2DAsync(w.rays, 24, batch, 36, 24, size_t(n), cudaMemcpyDeviceToDevice, st));
  if (precision == LNRF_PREC_BF16) {
    if ((rc = lnrf_nerf_pack_weights(p_c, coarse_packed, stream))) return rc;
    if ((rc = lnrf_nerf_pack_weights(p_f, fine_packed, stream))) return rc;
  }
  // ---- forward with the stash (render.py:39-91)
  if ((rc = lnrf_sample_coarse(w.rays, n, bbox_min_host, bbox_max_host, min_t_range, 1e-8f, u_coarse, Tc, w.t_min, w.t_max,
                               w.mask, w.ts_c, stream))) return rc;
  if ((rc = lnrf_nerf_mlp_fwd(p_c, coarse_packed, nullptr, nullptr, w.rays, w.ts_c, n, Tc, precision, 1, w.mlp_c,
                              w.mlp_c_bytes, w.dens_c, w.rgb_c, stream))) return rc;
  if ((rc = lnrf_composite_fwd(w.rays, w.ts_c, w.t_min, w.t_max, w.mask, w.dens_c, w.rgb_c, bg, n, Tc, w.out_c, nullptr,
                               nullptr, stream))) return rc;
  if ((rc = lnrf_sample_fine(w.ts_c, w.dens_c, w.t_min, w.t_max, u_fine, n, Tc, Tf, 1e-8f, w.ts_f, nullptr, nullptr,
                             stream))) return rc;
  if ((rc = lnrf_nerf_mlp_fwd(p_f, fine_packed, nullptr, nullptr, w.rays, w.ts_f, n, T2, precision, 1, w.mlp_f,
                              w.mlp_f_bytes, w.dens_f, w.rgb_f, stream))) return rc;
  if ((rc = lnrf_composite_fwd(w.rays, w.ts_f, w.t_min, w.t_max, w.mask, w.dens_f, w.rgb_f, bg, n, T2, w.out_f, nullptr,
                               nullptr, stream))) return rc;
  // ---- losses and backward, coarse level then fine level (train.py:140-151)
  const float inv_count = 1.0f / (3.0f * float(n));
  const float* targets = batch + 6;  // batch[:, 2], row stride 9
  struct Level { const float *ts, *dens, *rgb, *out; int T; float* P; const void* packed; void* ws; int64_t ws_bytes; int64_t goff; };
  const Level lv[2] = {{w.ts_c, w.dens_c, w.rgb_c, w.out_c, Tc, p_c, coarse_packed, w.mlp_c, w.mlp_c_bytes, 0},
                       {w.ts_f, w.dens_f, w.rgb_f, w.out_f, T2, p_f, fine_packed, w.mlp_f, w.mlp_f_bytes, np}};
  for (int li = 0; li < 2; ++li) {
    const Level& L = lv[li];
    if ((rc = lnrf_mse_loss(L.out, targets, 9, n, inv_count, scalars_out + li, w.d_out, stream))) return rc;
    if ((rc = lnrf_composite_bwd(L.ts, w.t_min, w.t_max, w.mask, L.dens, L.rgb, bg, w.d_out, n, L.T, w.d_dens, w.d_rgb,
                                 grads + 2 * np, stream))) return rc;
    if ((rc = lnrf_nerf_mlp_bwd(L.P, L.packed, n * L.T, precision, L.ws, L.ws_bytes, L.dens, L.rgb, w.d_dens, w.d_rgb,
                                grads + L.goff, stream))) return rc;
  }
  // ---- tree norms + optax.adam + apply_gradients (train.py:59, 92-106)
  return lnrf_adam_step(params, grads, adam_m, adam_v, count, lr, b1, b2, eps, step, 1.0f, scalars_out + 2, stream);
}

}  // extern "C"
