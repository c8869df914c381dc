// Error plumbing, init and small shared entry points of liblnrf.so.
#include <stdarg.h>
#include <string.h>

#include "lnrf_common.cuh"

namespace lnrf {

static thread_local char g_err[512] = "";
static int g_sm_count = 0;
static unsigned long long g_launches = 0;

void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return static_cast<int>(e);
}

int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_sm_count = n;
    else
      return 148;
  }
  return g_sm_count;
}

int init_mlp_tc();  // mlp_tc.cu: opt into large dynamic smem

}  // namespace lnrf

extern "C" {

const char* lnrf_last_error(void) { return lnrf::g_err; }

int lnrf_version(void) { return 1; }

int64_t lnrf_launch_count(void) { return (int64_t)lnrf::g_launches; }

int lnrf_init(int device) {
  LNRF_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  LNRF_CUDA(cudaGetDeviceProperties(&prop, device));
  LNRF_REQUIRE(prop.major == 10, LNRF_E_UNSUPPORTED,
               "lnrf_init: device %d is sm_%d%d; liblnrf is built for sm_100a only", device,
               prop.major, prop.minor);
  lnrf::g_sm_count = prop.multiProcessorCount;
  return lnrf::init_mlp_tc();
}

}  // extern "C"
