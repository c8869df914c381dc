// Error plumbing, init and small shared entry points of liblnrf.so.
#include <stdarg.h>
#include <string.h>

#include "lnrf_common.cuh"

namespace lnrf {

static thread_local char g_err[512] = "";
static int g_sm_count[64] = {0};  // per device ordinal, filled by lnrf_init / on first use
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

int sm_count() {  // SMs of the CURRENT device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = __atomic_load_n(&g_sm_count[dev], __ATOMIC_RELAXED);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    __atomic_store_n(&g_sm_count[dev], n, __ATOMIC_RELAXED);
  }
  return n;
}

int init_mlp_tc();   // mlp_tc.cu: constant tables + large dynamic smem opt-ins of the tcgen05 kernels
int init_ngp_mlp();  // ngp_mlp.cu
int init_ngp_tc();   // ngp_tc.cu
void init_mlp_fp32();  // mlp_fp32.cu: LNRF_FP32_FFMA
int init_gemm_tc();  // gemm_tc.cu: split-fp16 tcgen05 GEMMs of the fp32-accurate paths

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
  if (device >= 0 && device < 64) __atomic_store_n(&lnrf::g_sm_count[device], prop.multiProcessorCount, __ATOMIC_RELAXED);
  int rc = lnrf::init_mlp_tc();
  if (rc) return rc;
  if ((rc = lnrf::init_ngp_mlp())) return rc;
  if ((rc = lnrf::init_ngp_tc())) return rc;
  lnrf::init_mlp_fp32();
  return lnrf::init_gemm_tc();
}

}  // extern "C"
