/* lnrf.h — C ABI of the B200-native learn-nerf hot path (liblnrf.so).
 *
 * One entry point per implicit op group of the reference's hot path
 * (SURVEY.md §2.1 K1..K10).  The reference (unixpickle/learn-nerf) has no
 * plugin/FFI interface of its own: the path is ordinary JAX code behind three
 * Python call seams (render.py:320 model.apply, render.py:39 render_rays,
 * train.py:78 step_fn).  Each function below names the reference lines whose
 * arithmetic it replaces; learn-nerf_b200/learn_nerf/ binds them with ctypes and
 * re-exposes the reference's Python signatures, and INTEGRATION.md shows the
 * XLA-FFI handler a JAX maintainer would register instead.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the parameter is marked "host";
 *  - all arrays are dense, row-major, fp32 unless stated; `mask` is uint8 0/1;
 *  - the library never allocates, frees or synchronises: the caller owns every
 *    buffer (incl. workspaces sized by the *_workspace_bytes queries) and every
 *    call only enqueues work on `stream` (a cudaStream_t passed as void*);
 *  - return value: 0 ok; <0 invalid argument (see LNRF_E_*); >0 a cudaError_t.
 *    lnrf_last_error() returns a thread-local message for the last failure;
 *  - no CPU fallback exists anywhere: without a CUDA device every compute
 *    entry fails with a cudaError_t.
 */
#ifndef LNRF_H_
#define LNRF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LNRF_OK 0
#define LNRF_E_INVALID (-1)     /* null pointer, negative size, bad enum */
#define LNRF_E_UNSUPPORTED (-2) /* shape outside what the kernels implement */
#define LNRF_E_WORKSPACE (-3)   /* workspace too small / misaligned */

#define LNRF_PREC_FP32 0 /* fp32 SIMT FFMA path: 1e-5 abs vs the oracle */
#define LNRF_PREC_BF16 1 /* bf16 tcgen05/TMEM path: 2e-2 abs vs the oracle */

typedef void* lnrf_stream_t; /* cudaStream_t */

const char* lnrf_last_error(void);
int lnrf_version(void);
/* One-time per-process setup for `device` (function attributes, SM count). */
int lnrf_init(int device);
/* Number of kernels this process has launched through the library so far. */
int64_t lnrf_launch_count(void);

/* ---------------------------------------------------------------- K1 sampling
 * ray_t_range (render.py:346-389, vmapped at :93-111) + stratified_sampling
 * (render.py:121-143).  Bit-exact with oracle.render_np given the same `u`.
 * rays [n,2,3]; bbox_min/bbox_max host[3]; u [n,T] in [0,1);
 * out: t_min[n], t_max[n], mask[n] (uint8), ts[n,T].                         */
int lnrf_sample_coarse(const float* rays, int64_t n, const float* bbox_min_host,
                       const float* bbox_max_host, float min_t_range, float epsilon,
                       const float* u, int32_t T, float* t_min, float* t_max,
                       uint8_t* mask, float* ts, lnrf_stream_t stream);

/* stratified_sampling alone (render.py:121-143) on given bounds t_min/t_max[n]:
 * ts[n,T] = (k*bin + t_min) + u*bin, bin = (t_max - t_min)/T.  Bit-exact.        */
int lnrf_stratified(const float* t_min, const float* t_max, const float* u, int64_t n, int32_t T,
                    float* ts, lnrf_stream_t stream);

/* ---------------------------------------------------------------- K4 fine sampling
 * RaySamples.fine_sampling (render.py:211-257) incl. termination_probs
 * (:270-287), the inverse CDF via jnp.interp and the final sort.  Bit-exact with
 * the oracle given the same (ts, densities, u).  ts_c/dens_c [n,Tc]; u [n,Tf];
 * out ts_out [n,Tc+Tf] ascending.  Optional (nullable) debug outputs:
 * idx_out [n,Tf] int32 = interp bin index i in [1,Tc]; new_ts_out [n,Tf] =
 * the unsorted inverse-CDF samples.  Tc <= 256, Tc+Tf <= 1024.               */
int lnrf_sample_fine(const float* ts_c, const float* dens_c, const float* t_min,
                     const float* t_max, const float* u, int64_t n, int32_t Tc, int32_t Tf,
                     float eps, float* ts_out, int32_t* idx_out, float* new_ts_out,
                     lnrf_stream_t stream);

/* RaySamples.starts / ends / deltas (render.py:259-268): each [n,T]; any of the three outputs may
 * be NULL.  Bit-exact with the oracle.                                         */
int lnrf_ray_intervals(const float* ts, const float* t_min, const float* t_max, int64_t n, int32_t T,
                       float* starts, float* ends, float* deltas, lnrf_stream_t stream);
/* RaySamples.termination_probs (render.py:270-287): probs[n,T+1], the last column is the
 * probability of escaping to the background.  Strictly sequential fp32 cumsum and the shared
 * deterministic exp, as inside lnrf_sample_fine: bit-exact with the oracle.  T <= 1024.        */
int lnrf_termination_probs(const float* ts, const float* t_min, const float* t_max, const float* dens,
                           int64_t n, int32_t T, float* probs, lnrf_stream_t stream);

/* ---------------------------------------------------------------- K3 compositing
 * termination_probs + RaySamples.render_rays / render_alpha and the coords
 * render (render.py:155-190, 270-287, 329-331).  dens [n,T]; rgb [n,T,3];
 * background device[3]; out outputs[n,3], alphas[n] ([n,1]), coords[n,3]
 * (alphas/coords nullable).                                                  */
int lnrf_composite_fwd(const float* rays, const float* ts, const float* t_min,
                       const float* t_max, const uint8_t* mask, const float* dens,
                       const float* rgb, const float* background, int64_t n, int32_t T,
                       float* outputs, float* alphas, float* coords, lnrf_stream_t stream);

/* K5: reverse-scan gradient of K3 w.r.t. dens, rgb and background given
 * d_outputs[n,3] (what jax.grad derives at train.py:90).
 * d_dens[n,T], d_rgb[n,T,3] are overwritten; d_background[3] is ACCUMULATED
 * (caller zeroes it).                                                         */
int lnrf_composite_bwd(const float* ts, const float* t_min, const float* t_max,
                       const uint8_t* mask, const float* dens, const float* rgb,
                       const float* background, const float* d_outputs, int64_t n, int32_t T,
                       float* d_dens, float* d_rgb, float* d_background, lnrf_stream_t stream);

/* MSE loss and its gradient (train.py:140-142): loss_sum[0] += sum((out-tgt)^2)
 * over this call's n*3 values (caller divides by the global count);
 * d_outputs = 2*(out-tgt)*inv_count.  targets is a strided view: element
 * (i,c) at targets[i*target_stride + c] (batch[:,2] of an [n,3,3] batch has
 * stride 9).                                                                  */
int lnrf_mse_loss(const float* outputs, const float* targets, int64_t target_stride, int64_t n,
                  float inv_count, float* loss_sum, float* d_outputs, lnrf_stream_t stream);

/* ---------------------------------------------------------------- K2/K6 NeRF MLP
 * NeRFModel.__call__ with the default architecture (model.py:35-62):
 * sinusoidal_emb(x,10)/(d,4), 5+4 Dense(256) with the skip concat, softplus
 * density head, 128-wide colour layer, tanh rgb.  Other sizes: LNRF_E_UNSUPPORTED.
 *
 * params: flat fp32, Dense_0.kernel[60,256], Dense_0.bias[256], Dense_1.kernel,
 * ... Dense_11.bias[3] (kernel row-major [in,out] as in Flax), every tensor
 * starting on a 4-float boundary (zero padding): see lnrf_nerf_param_offsets.
 *
 * Inputs are either explicit points (x[m,3], d[m,3]; rays==NULL) as in
 * model.apply (render.py:320-324), or ray mode (x==NULL): rays[n,2,3] and
 * ts[n,T] with m = n*T, point = o + d*t (render.py:145-153, :319).
 * Outputs dens[m], rgb[m,3].  With save_for_backward != 0 the activations
 * needed by lnrf_nerf_mlp_bwd are kept in `workspace`.                        */
int64_t lnrf_nerf_param_count(void);  /* logical parameters: 593,924 */
int64_t lnrf_nerf_param_floats(void); /* floats in the flat buffer incl. 16-byte padding */
/* host out[24]: float offsets of kernel_i (out[2i]) and bias_i (out[2i+1]). */
int lnrf_nerf_param_offsets(int64_t* out_host);
/* bf16 path: bf16 UMMA operand images of `params` (lnrf_nerf_packed_bytes() bytes,
 * 1024-byte aligned).  Rebuild with lnrf_nerf_pack_weights after every update.  */
int64_t lnrf_nerf_packed_bytes(void);
int lnrf_nerf_pack_weights(const float* params, void* packed, lnrf_stream_t stream);
int lnrf_nerf_mlp_workspace_bytes(int64_t m, int32_t precision, int32_t save_for_backward,
                                  int64_t* bytes_out_host);
/* `packed` is required for LNRF_PREC_BF16 and ignored (may be NULL) for fp32. */
int lnrf_nerf_mlp_fwd(const float* params, const void* packed, const float* x, const float* d,
                      const float* rays, const float* ts, int64_t n, int32_t T, int32_t precision,
                      int32_t save_for_backward, void* workspace, int64_t workspace_bytes,
                      float* dens, float* rgb, lnrf_stream_t stream);
/* NeRFRenderer.render_rays (render.py:39-91) for two NeRFModels as ONE call: t_range + stratified coarse
 * sampling, coarse model, compositing, inverse-CDF fine sampling (eps 1e-8, render.py:217) of Tf more points,
 * fine model over the Tc + Tf sorted points, compositing.  rays[n,2,3]; u_coarse[n,Tc] / u_fine[n,Tf] are the
 * uniforms of jax.random.uniform(coarse_key) / (fine_key) (render.py:55,142; lnrf_threefry_uniform produces
 * them from a key); *_packed as for lnrf_nerf_mlp_fwd (bf16 only).  Outputs: coarse_outputs[n,3],
 * fine_outputs[n,3], and optionally the fine level's alphas[n] and coords[n,3] (nullable).  Same kernels and
 * bits as the six separate calls.  workspace: lnrf_nerf_render_workspace_bytes, 1024-byte aligned.        */
int lnrf_nerf_render_workspace_bytes(int64_t n, int32_t Tc, int32_t Tf, int32_t precision, int64_t* bytes_out_host);
int lnrf_nerf_render_rays(const float* rays, const float* bbox_min_host, const float* bbox_max_host, float min_t_range,
                          const float* u_coarse, const float* u_fine, const float* coarse_params,
                          const void* coarse_packed, const float* fine_params, const void* fine_packed,
                          int32_t precision, const float* background, int64_t n, int32_t Tc, int32_t Tf,
                          void* workspace, int64_t workspace_bytes, float* coarse_outputs, float* fine_outputs,
                          float* fine_alphas, float* fine_coords, lnrf_stream_t stream);
/* TrainLoop.step_fn (train.py:78-112 around the losses of train.py:114-151) for two NeRFModels on one device as
 * ONE call: both levels rendered with the activation stash, the two MSE losses, compositing and MLP backward of both
 * levels, tree norms + optax.adam + apply_gradients.  batch[n,3,3] = (origin, direction, target colour);
 * params / adam_m / adam_v / grads are flat buffers [coarse model | fine model | background(3) + 1 pad] of
 * 2 * lnrf_nerf_param_floats() + 4 floats (grads is overwritten); coarse_packed / fine_packed: scratch of
 * lnrf_nerf_packed_bytes() each, 1024-byte aligned, re-packed inside the call (bf16 only, NULL for fp32);
 * step is 1-based.  scalars_out (device[4]) receives {sum of squared errors coarse, fine (divide by 3 n),
 * |grad|^2, |params before the update|^2}.  No density penalty / aux losses / gradient exchange: those stay with
 * the per-op entries.  workspace: lnrf_nerf_train_workspace_bytes, 1024-byte aligned.                       */
int lnrf_nerf_train_workspace_bytes(int64_t n, int32_t Tc, int32_t Tf, int32_t precision, int64_t* bytes_out_host);
int lnrf_nerf_train_step(const float* batch, const float* bbox_min_host, const float* bbox_max_host, float min_t_range,
                         const float* u_coarse, const float* u_fine, float* params, float* adam_m, float* adam_v,
                         float* grads, void* coarse_packed, void* fine_packed, int32_t precision, int64_t n, int32_t Tc,
                         int32_t Tf, float lr, float b1, float b2, float eps, int32_t step, void* workspace,
                         int64_t workspace_bytes, float* scalars_out, lnrf_stream_t stream);
/* Gradient of the above w.r.t. params given d_dens[m], d_rgb[m,3]; uses the
 * workspace written by the matching forward call.  d_params
 * (lnrf_nerf_param_floats() floats) is ACCUMULATED.  Inputs x/d/rays/ts carry
 * no gradient.                                                 */
int lnrf_nerf_mlp_bwd(const float* params, const void* packed, int64_t m, int32_t precision,
                      void* workspace, int64_t workspace_bytes, const float* dens, const float* rgb,
                      const float* d_dens, const float* d_rgb, float* d_params,
                      lnrf_stream_t stream);

/* ---------------------------------------------------------------- K10 optimiser
 * tree_norm x2 + optax.adam + apply_updates (train.py:59,92-106) over one flat
 * buffer.  g' = grads*grad_scale (1/world after an all-reduce);
 * norms_out[0] += sum(g'^2), norms_out[1] += sum(params_before^2) (caller
 * zeroes, takes sqrt); m,v,params updated in place; step is 1-based.          */
int lnrf_adam_step(float* params, const float* grads, float* m, float* v, int64_t count,
                   float lr, float b1, float b2, float eps, int32_t step, float grad_scale,
                   float* norms_out, lnrf_stream_t stream);
/* Variants for CUDA-graph replays (captured once, replayed every step): the two inputs that change
 * from step to step are read from device memory.  inv_bias_corr_dev = {1/(1-b1^t), 1/(1-b2^t)}
 * (what lnrf_adam_step derives from `step` on the host); key_dev = the two Threefry key words. */
int lnrf_adam_step_dk(float* params, const float* grads, float* m, float* v, int64_t count,
                      float lr, float b1, float b2, float eps, const float* inv_bias_corr_dev,
                      float grad_scale, float* norms_out, lnrf_stream_t stream);
int lnrf_threefry_uniform_dk(const uint32_t* key_dev, int64_t n, float* out, lnrf_stream_t stream);
/* Data-parallel variant (the reference is single-device, train.py:85-106): fused
 * all-reduce + Adam.  peer_grads is a HOST array of `world` device addresses, one flat gradient
 * buffer per rank (NVLink peer / symmetric-memory mappings, own rank included), each holding
 * count + extra floats.  Gradients are summed in rank order and scaled by grad_scale inside the
 * optimiser pass; the `extra` trailing floats (per-rank loss sums) are summed into extra_out.
 * The caller brackets the call with cross-rank barriers.                                      */
int lnrf_adam_step_peers(float* params, const uint64_t* peer_grads, int32_t world, float* m, float* v,
                         int64_t count, int32_t extra, float lr, float b1, float b2, float eps,
                         int32_t step, float grad_scale, float* norms_out, float* extra_out,
                         const float* inv_bias_corr_dev /* nullable: overrides `step`, see lnrf_adam_step_dk */,
                         lnrf_stream_t stream);

/* ---------------------------------------------------------------- K7/K8 hash grid
 * MultiresHashTableEncoding / HashTableEncoding / hash_table_lookup
 * (instant_ngp.py:92-224), feature_dim F=2.  tables: all levels concatenated,
 * level l starts at float offset level_offsets_host[l]; level l has
 * grid_sizes_host[l]; it is hashed iff grid^3 > table_sizes_host[l].
 * x[m,3] -> enc[m,2L].  L <= 16.                                              */
int lnrf_hashgrid_fwd(const float* tables, const int64_t* level_offsets_host,
                      const int32_t* grid_sizes_host, const int32_t* table_sizes_host, int32_t L,
                      const float* bbox_min_host, const float* bbox_max_host, int32_t smooth,
                      const float* x, const float* rays, const float* ts, int64_t n, int32_t T,
                      float* enc, lnrf_stream_t stream);
/* scatter-add of d_enc[m,2L] into d_tables (ACCUMULATED; same layout as tables). */
int lnrf_hashgrid_bwd(const int64_t* level_offsets_host, const int32_t* grid_sizes_host,
                      const int32_t* table_sizes_host, int32_t L, const float* bbox_min_host,
                      const float* bbox_max_host, int32_t smooth, const float* x,
                      const float* rays, const float* ts, int64_t n, int32_t T,
                      const float* d_enc, float* d_tables, lnrf_stream_t stream);

/* InstantNGPModel heads (instant_ngp.py:37,46-53) on a precomputed encoding:
 * enc[m,2L] (+ d[m,3] or ray mode) -> dens[m], rgb[m,3].  params: Dense_0..4
 * flat (kernel then bias each).  fp32-accurate (1e-5).  With save_for_backward
 * the layers run as split-fp16 tcgen05 GEMMs and the workspace keeps the layer
 * inputs, ReLU bit masks and operand ranges for lnrf_ngp_mlp_bwd; without it
 * the forward is one fused kernel and workspace may be NULL.                  */
int64_t lnrf_ngp_mlp_param_count(int32_t L); /* floats incl. 16-byte padding of each tensor */
/* host out[10]: float offsets of kernel_i (out[2i]) and bias_i (out[2i+1]), i = 0..4. */
int lnrf_ngp_mlp_param_offsets(int32_t L, int64_t* out_host);
int lnrf_ngp_mlp_workspace_bytes(int64_t m, int32_t L, int64_t* bytes_out_host);
int lnrf_ngp_mlp_fwd(const float* params, int32_t L, const float* enc, const float* d,
                     const float* rays, int64_t n, int32_t T, int32_t save_for_backward, void* workspace,
                     int64_t workspace_bytes, float* dens, float* rgb, lnrf_stream_t stream);
int lnrf_ngp_mlp_bwd(const float* params, int32_t L, const float* enc, int64_t m, void* workspace,
                     int64_t workspace_bytes, const float* dens, const float* rgb,
                     const float* d_dens, const float* d_rgb, float* d_params, float* d_enc,
                     lnrf_stream_t stream);

/* bf16 tensor-core variant of the same heads (tcgen05 / TMEM; 2e-2 abs on density / rgb like the bf16 NeRF
 * path, gradients rel-L2 5e-2).  `packed`: bf16 operand images + biases built from the flat head
 * parameters by lnrf_ngp_pack_weights (lnrf_ngp_packed_bytes() bytes, 1024-byte aligned; rebuild after
 * every update).  With save_for_backward the forward leaves the five layer inputs of every 128-sample
 * tile in `workspace` (lnrf_ngp_mlp_tc_workspace_bytes, 1024-byte aligned: 640 B/sample); the backward
 * reads them, ACCUMULATES the head gradients into d_params (layout of lnrf_ngp_mlp_param_offsets) and
 * overwrites d_enc[m,2L].                                                                          */
int64_t lnrf_ngp_packed_bytes(void);
int lnrf_ngp_pack_weights(const float* params, int32_t L, void* packed, lnrf_stream_t stream);
int lnrf_ngp_mlp_tc_workspace_bytes(int64_t m, int64_t* bytes_out_host);
int lnrf_ngp_mlp_fwd_tc(const void* packed, int32_t L, const float* enc, const float* d, const float* rays,
                        int64_t n, int32_t T, int32_t save_for_backward, void* workspace,
                        int64_t workspace_bytes, float* dens, float* rgb, lnrf_stream_t stream);
int lnrf_ngp_mlp_bwd_tc(const void* packed, int32_t L, int64_t m, const void* workspace,
                        int64_t workspace_bytes, const float* dens, const float* rgb, const float* d_dens,
                        const float* d_rgb, float* d_params, float* d_enc, lnrf_stream_t stream);

/* ---------------------------------------------------------------- K9 Ref-NeRF
 * RefNERFModel(sh_degree=4) (ref_nerf.py:34-107): spatial MLP (as NeRF's trunk), real_normal
 * from the input gradient of -spatial_out[:,0] (:38-43), activations, reflection direction,
 * integrated directional encoding, directional block, sRGB colour and the two aux losses
 * normal_mse / neg_normal (:72-75).  fp32 only.  params: Dense_0..10 flat, kernel then bias,
 * 16-byte aligned; Dense_9's kernel is stored with 276 rows (273 used, 3 zero rows).
 * Inputs as for lnrf_nerf_mlp_fwd: (x[m,3], d[m,3]) or ray mode (rays[n,2,3], ts[n,T]).
 * Outputs dens[m], rgb[m,3], aux_normal_mse[m], aux_neg_normal[m].                    */
int64_t lnrf_refnerf_param_count(void);  /* logical parameters: 592,771 */
int64_t lnrf_refnerf_param_floats(void);
/* host out[22]: float offsets of kernel_i (out[2i]) and bias_i (out[2i+1]), i = 0..10. */
int lnrf_refnerf_param_offsets(int64_t* out_host);
int lnrf_refnerf_workspace_bytes(int64_t m, int32_t save_for_backward, int64_t* bytes_out_host);
int lnrf_refnerf_fwd(const float* params, const float* x, const float* d, const float* rays, const float* ts,
                     int64_t n, int32_t T, int32_t save_for_backward, void* workspace, int64_t workspace_bytes,
                     float* dens, float* rgb, float* aux_normal_mse, float* aux_neg_normal,
                     lnrf_stream_t stream);
/* Gradient w.r.t. params (ACCUMULATED into d_params) given the gradients of all four outputs;
 * includes the second-order term through real_normal.  Same inputs and workspace as the
 * forward call it follows.                                                             */
int lnrf_refnerf_bwd(const float* params, const float* x, const float* d, const float* rays, const float* ts,
                     int64_t n, int32_t T, void* workspace, int64_t workspace_bytes, const float* d_dens,
                     const float* d_rgb, const float* d_aux_normal_mse, const float* d_aux_neg_normal,
                     float* d_params, lnrf_stream_t stream);

/* ---------------------------------------------------------------- ray generation / image assembly
 * CameraView.bare_rays (dataset.py:52-78): rays of image rows [row0, row0+rows) of a width x height
 * view in raster order, rays[rows*width, 2, 3] = (origin, normalised direction).  tan_half_x_fov /
 * tan_half_y_fov are tan(fov/2) evaluated on the host in double precision (as math.tan at
 * dataset.py:61,66) and rounded to fp32.  Bit-exact with the oracle restatement.            */
int lnrf_bare_rays(const float* origin_host, const float* x_axis_host, const float* y_axis_host,
                   const float* z_host, float tan_half_x_fov, float tan_half_y_fov, int32_t width,
                   int32_t height, int32_t row0, int32_t rows, float* rays, lnrf_stream_t stream);
/* ((colors + 1) * 127.5).astype(uint8) (render_nerf.py:93-96); colors are clamped to [-1, 1].   */
int lnrf_rgb_to_u8(const float* colors, int64_t count, uint8_t* out, lnrf_stream_t stream);

/* Depth image of scripts/render_new_dataset.py:96-133 from the fine level's coords[n,3] / alphas[n]:
 * z = clip(where(alpha > 0.9, ((coords - origin) . direction) / (alpha + 1e-8), max_depth), 0, max_depth)
 * / max_depth -> z_out[n] (nullable) and depth_u32_out[n] = (z * 0xFFFF) truncated (nullable).     */
int lnrf_z_depth(const float* coords, const float* alphas, const float* camera_origin_host,
                 const float* camera_direction_host, float max_depth, int64_t n, float* z_out,
                 uint32_t* depth_u32_out, lnrf_stream_t stream);

/* jax.random.uniform(key, [n]) in fp32 (render.py:142): Threefry-2x32 over iota(n) as JAX's
 * non-partitionable threefry does, 23 mantissa bits; key = the two uint32 words of the key.  */
int lnrf_threefry_uniform(uint32_t key0, uint32_t key1, int64_t n, float* out, lnrf_stream_t stream);

/* ---------------------------------------------------------------- diagnostics
 * Single 128xNxK bf16 GEMM tile on tcgen05 (A[128,K], B[N,K] both K-major,
 * D fp32 [128,N]) used by tests to pin the UMMA descriptor encodings.         */
int lnrf_debug_umma_gemm(const float* a, const float* b, int32_t N, int32_t K, float* d_out,
                         lnrf_stream_t stream);
/* Same for the transposed (dW) shape: D[M,N] = At[128,M]^T * Bt[128,N], both operands
 * MN-major views of [128 x 64] SW128 block images; M in {128,256}, N in {64,128,192,256}. */
int lnrf_debug_umma_gemm_tn(const float* at, const float* bt, int32_t M, int32_t N, float* d_out,
                            lnrf_stream_t stream);
/* The fp32-accurate GEMM engine of the 1e-5 paths on its own (split-fp16 tcgen05, csrc/gemm_tc.cu): what
 * nn.Dense (model.py:51-60, ref_nerf.py:92-107) and its jax.grad transposes lower to.
 *   mode 0: C[M,N] = epi((A0[M,K0] | A1[M,K1]) @ B[K0+K1,N])     mode 1: the same with B given as Bt[N,K]
 *   mode 2: C[K0,N] += A0[M,K0]^T @ B[M,N] (M = samples), db[N] += column sums of B (nullable)
 *   mode 3: *c_amax = max(*c_amax, max |A0[0..M)|)
 * epi: 0 bias+ReLU (also writes the bit mask [C > 0] to mask_out if given: M/32 * N/32 * 32 words, rounded up),
 * 1 bias, 2 mask by aux[m,n] > 0, 3 + r1s[m] r1w[n], 5 plain store, 6 mask by the bits of mask_in.  a_amax / b_amax:
 * device addresses of max|operand| for operands far from O(1) (gradients), nullable; c_amax: receives
 * max |C| by atomic max (the caller zeroes it), nullable.  N <= 256, K0 + K1 <= 320 (64-column chunks),
 * every dimension / leading dimension a multiple of 4, 16-byte aligned bases.                           */
int lnrf_tcgemm(int32_t mode, int32_t epi, int64_t M, int32_t N, const float* A0, int32_t lda0, int32_t K0,
                const float* A1, int32_t lda1, int32_t K1, const float* B, int32_t ldb, float* C, int32_t ldc,
                const float* bias, const float* aux, int32_t ldaux, const float* r1s, const float* r1w, float* db,
                const float* a_amax, const float* b_amax, float* c_amax, const uint32_t* mask_in, uint32_t* mask_out,
                lnrf_stream_t stream);
/* ---- InstantNGPRefNERFModel (instant_ngp.py:57-89 on RefNERFBase, ref_nerf.py:34-77): Ref-NeRF
 * heads on a SMOOTH multiresolution hash grid.  Flat parameter layout: Dense_0 [2L,64], Dense_1
 * [64,16], Dense_2 [36 rows (33 used),64], Dense_3 [64,64], Dense_4 [64,3] at
 * lnrf_ngpref_param_offsets (w0,b0,..,w4,b4), then the tables at level_offsets (floats from the
 * start of `params`, as for lnrf_hashgrid_fwd).  Outputs / aux losses as lnrf_refnerf_fwd; the
 * backward includes the second-order term through the hash grid's input Jacobian.            */
int64_t lnrf_ngpref_mlp_param_floats(int32_t L);
int lnrf_ngpref_param_offsets(int32_t L, int64_t* out_host);
int lnrf_ngpref_workspace_bytes(int64_t m, int32_t L, int32_t save_for_backward, int64_t* bytes_out_host);
int lnrf_ngpref_fwd(const float* params, const int64_t* level_offsets_host, const int32_t* grid_sizes_host,
                    const int32_t* table_sizes_host, int32_t L, const float* bbox_min_host,
                    const float* bbox_max_host, const float* x, const float* d, const float* rays, const float* ts,
                    int64_t n, int32_t T, int32_t save_for_backward, void* workspace, int64_t workspace_bytes,
                    float* dens, float* rgb, float* aux_normal_mse, float* aux_neg_normal, lnrf_stream_t stream);
int lnrf_ngpref_bwd(const float* params, const int64_t* level_offsets_host, const int32_t* grid_sizes_host,
                    const int32_t* table_sizes_host, int32_t L, const float* bbox_min_host,
                    const float* bbox_max_host, const float* x, const float* d, const float* rays, const float* ts,
                    int64_t n, int32_t T, void* workspace, int64_t workspace_bytes, const float* d_dens,
                    const float* d_rgb, const float* d_aux_normal_mse, const float* d_aux_neg_normal,
                    float* d_params, lnrf_stream_t stream);
#ifdef __cplusplus
}
#endif
#endif /* LNRF_H_ */
