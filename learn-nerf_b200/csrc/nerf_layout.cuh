// Parameter layout of the default NeRFModel (learn_nerf/model.py:35-62) inside one
// flat fp32 buffer, shared by the fp32 and bf16 paths.
//
// Order: Dense_0.kernel, Dense_0.bias, Dense_1.kernel, ..., Dense_11.bias.  Each
// kernel is row-major [in,out] exactly as Flax stores it.  Every tensor starts on a
// 4-float (16-byte) boundary so the kernels can use 128-bit loads; the padding
// floats are zero, receive zero gradient and are invisible to the norms.
#pragma once
#include <stdint.h>

namespace lnrf {

constexpr int kNerfLayers = 12;
constexpr int kXFreqs = 10, kDFreqs = 4;
constexpr int kXE = 6 * kXFreqs;  // 60
constexpr int kDE = 6 * kDFreqs;  // 24
constexpr int kH = 256, kHC = 128;

struct NerfLayout {
  int in[kNerfLayers];
  int out[kNerfLayers];
  int64_t w[kNerfLayers];  // float offset of kernel_i
  int64_t b[kNerfLayers];  // float offset of bias_i
  int64_t total;           // padded float count
};

constexpr NerfLayout make_nerf_layout() {
  NerfLayout L{};
  const int ins[kNerfLayers] = {kXE, kH, kH, kH, kH, kH + kXE, kH, kH, kH, kH, kH + kDE, kHC};
  const int outs[kNerfLayers] = {kH, kH, kH, kH, kH, kH, kH, kH, kH, 1, kHC, 3};
  int64_t off = 0;
  for (int i = 0; i < kNerfLayers; ++i) {
    L.in[i] = ins[i];
    L.out[i] = outs[i];
    L.w[i] = off;
    off += int64_t(ins[i]) * outs[i];
    off = (off + 3) / 4 * 4;
    L.b[i] = off;
    off += outs[i];
    off = (off + 3) / 4 * 4;
  }
  L.total = off;
  return L;
}

constexpr NerfLayout kNerf = make_nerf_layout();

}  // namespace lnrf
