// Device-side ray generation and image assembly around the render path:
// CameraView.bare_rays (learn_nerf/dataset.py:52-78) and the float -> uint8 conversion of
// render_nerf.py:93-96.  Both are trivially HBM-bound: 24 B written per ray, 15 B per pixel.
#include "lnrf_common.cuh"

namespace lnrf {

struct Camera {
  float origin[3], x_axis[3], y_axis[3], z[3];
  float tan_x, tan_y;  // tan(fov / 2), rounded to fp32 as JAX does with the Python scalar
};

// jnp.linspace(-1, 1, num)[i] in fp32: start + i * ((stop - start) / (num - 1)), last point = stop
__device__ __forceinline__ float linspace_pm1(int i, int num) {
  if (num == 1) return -1.0f;
  if (i == num - 1) return 1.0f;
  const float step = __fdiv_rn(2.0f, float(num - 1));
  return __fadd_rn(-1.0f, __fmul_rn(float(i), step));
}

// rays[(row - row0) * width + col] = (origin, normalize(xs[col] + ys[row] + z)), raster order
__global__ void __launch_bounds__(256)
bare_rays_kernel(Camera cam, int width, int height, int row0, int rows, float* __restrict__ rays) {
  const int64_t total = int64_t(rows) * width;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int row = row0 + int(i / width), col = int(i % width);
    const float ly = __fmul_rn(cam.tan_y, linspace_pm1(row, height));  // dataset.py:60-64
    const float lx = __fmul_rn(cam.tan_x, linspace_pm1(col, width));   // :65-69
    float d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)  // (xs + ys) + z, :70
      d[a] = __fadd_rn(__fadd_rn(__fmul_rn(lx, cam.x_axis[a]), __fmul_rn(ly, cam.y_axis[a])), cam.z[a]);
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]));
    const float nrm = __fsqrt_rn(n2);                                  // :71
    float* o = rays + i * 6;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      o[a] = cam.origin[a];
      o[3 + a] = __fdiv_rn(d[a], nrm);
    }
  }
}

// ((c + 1) * 127.5).astype(uint8) with c clamped to [-1, 1] first (render_nerf.py:93-96)
__global__ void __launch_bounds__(256)
rgb_to_u8_kernel(const float* __restrict__ colors, int64_t count, uint8_t* __restrict__ out) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += int64_t(gridDim.x) * blockDim.x) {
    const float c = fminf(fmaxf(__ldg(colors + i), -1.0f), 1.0f);
    out[i] = uint8_t(__fmul_rn(__fadd_rn(c, 1.0f), 127.5f));  // truncation, as numpy's astype
  }
}

// z-depth image of render_new_dataset.py:96-133: the expected hit point (coords / alpha) projected on
// the camera direction, max_depth where the ray is mostly transparent (alpha <= 0.9), clipped to
// [0, max_depth] and normalised; depth_u32 = (z * 0xFFFF).astype(uint32).  Individually rounded ops in
// the reference's order ((c - o) @ dir as a left-to-right dot product).
struct DepthCam { float origin[3], dir[3]; };
__global__ void __launch_bounds__(256)
z_depth_kernel(const float* __restrict__ coords, const float* __restrict__ alphas, DepthCam cam, float max_depth,
               int64_t n, float* __restrict__ z_out, uint32_t* __restrict__ u32_out) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float a = __ldg(alphas + i);
    float dot = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float t = __fmul_rn(__fsub_rn(__ldg(coords + i * 3 + k), cam.origin[k]), cam.dir[k]);
      dot = (k == 0) ? t : __fadd_rn(dot, t);
    }
    float z = (a > 0.9f) ? __fdiv_rn(dot, __fadd_rn(a, 1e-8f)) : max_depth;
    z = __fdiv_rn(fminf(fmaxf(z, 0.0f), max_depth), max_depth);
    if (z_out) z_out[i] = z;
    if (u32_out) u32_out[i] = uint32_t(__fmul_rn(z, 65535.0f));
  }
}

static inline unsigned rg_blocks(int64_t items) {
  int64_t b = ceil_div(items, 256);
  const int64_t cap = int64_t(sm_count()) * 16;
  return unsigned(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace lnrf

extern "C" {

int lnrf_bare_rays(const float* origin_host, const float* x_axis_host, const float* y_axis_host,
                   const float* z_host, float x_fov, float y_fov, int32_t width, int32_t height, int32_t row0,
                   int32_t rows, float* rays, lnrf_stream_t stream) {
  LNRF_REQUIRE(origin_host && x_axis_host && y_axis_host && z_host, LNRF_E_INVALID, "lnrf_bare_rays: null host pointer");
  LNRF_REQUIRE(width >= 1 && height >= 1 && row0 >= 0 && rows >= 0 && row0 + rows <= height, LNRF_E_INVALID,
               "lnrf_bare_rays: width=%d height=%d row0=%d rows=%d", width, height, row0, rows);
  if (rows == 0) return LNRF_OK;
  LNRF_REQUIRE(rays, LNRF_E_INVALID, "lnrf_bare_rays: null output");
  lnrf::Camera cam;
  for (int a = 0; a < 3; ++a) {
    cam.origin[a] = origin_host[a];
    cam.x_axis[a] = x_axis_host[a];
    cam.y_axis[a] = y_axis_host[a];
    cam.z[a] = z_host[a];
  }
  cam.tan_x = x_fov;  // the caller passes tan(fov / 2) evaluated in double and rounded once
  cam.tan_y = y_fov;
  lnrf::bare_rays_kernel<<<lnrf::rg_blocks(int64_t(rows) * width), 256, 0, lnrf::as_stream(stream)>>>(
      cam, width, height, row0, rows, rays);
  LNRF_LAUNCH_CHECK("bare_rays_kernel");
  return LNRF_OK;
}

int lnrf_rgb_to_u8(const float* colors, int64_t count, uint8_t* out, lnrf_stream_t stream) {
  LNRF_REQUIRE(count >= 0, LNRF_E_INVALID, "lnrf_rgb_to_u8: count=%lld", (long long)count);
  if (count == 0) return LNRF_OK;
  LNRF_REQUIRE(colors && out, LNRF_E_INVALID, "lnrf_rgb_to_u8: null pointer");
  lnrf::rgb_to_u8_kernel<<<lnrf::rg_blocks(count), 256, 0, lnrf::as_stream(stream)>>>(colors, count, out);
  LNRF_LAUNCH_CHECK("rgb_to_u8_kernel");
  return LNRF_OK;
}

int lnrf_z_depth(const float* coords, const float* alphas, const float* camera_origin_host,
                 const float* camera_direction_host, float max_depth, int64_t n, float* z_out, uint32_t* depth_u32_out,
                 lnrf_stream_t stream) {
  LNRF_REQUIRE(n >= 0 && max_depth > 0.0f, LNRF_E_INVALID, "lnrf_z_depth: n=%lld max_depth=%g", (long long)n, max_depth);
  if (n == 0) return LNRF_OK;
  LNRF_REQUIRE(coords && alphas && camera_origin_host && camera_direction_host && (z_out || depth_u32_out),
               LNRF_E_INVALID, "lnrf_z_depth: null pointer");
  lnrf::DepthCam cam;
  for (int a = 0; a < 3; ++a) {
    cam.origin[a] = camera_origin_host[a];
    cam.dir[a] = camera_direction_host[a];
  }
  lnrf::z_depth_kernel<<<lnrf::rg_blocks(n), 256, 0, lnrf::as_stream(stream)>>>(coords, alphas, cam, max_depth, n, z_out,
                                                                               depth_u32_out);
  LNRF_LAUNCH_CHECK("z_depth_kernel");
  return LNRF_OK;
}

}  // extern "C"
