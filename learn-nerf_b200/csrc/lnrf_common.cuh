// Shared host/device helpers for liblnrf.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lnrf.h"

namespace lnrf {

// thread-local error text behind lnrf_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();
void count_launch();  // bumps the counter behind lnrf_launch_count()

#define LNRF_REQUIRE(cond, code, ...)     \
  do {                                    \
    if (!(cond)) {                        \
      ::lnrf::set_error(__VA_ARGS__);     \
      return (code);                      \
    }                                     \
  } while (0)

#define LNRF_CUDA(call)                                        \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return ::lnrf::cuda_fail(e__, #call); \
  } while (0)

// after a kernel launch: report launch-configuration errors without syncing
#define LNRF_LAUNCH_CHECK(name)                                        \
  do {                                                                 \
    ::lnrf::count_launch();                                            \
    cudaError_t e__ = cudaGetLastError();                              \
    if (e__ != cudaSuccess) return ::lnrf::cuda_fail(e__, name);       \
  } while (0)

static inline cudaStream_t as_stream(lnrf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ static inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

}  // namespace lnrf
