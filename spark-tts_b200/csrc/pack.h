// Host-side weight re-layout for the tcgen05 implicit-GEMM kernels (pure C++, no CUDA calls).
#pragma once

#include <cstdint>
#include <vector>

#include "common.cuh"

namespace sparkcodec {

struct PackedGemm {
  int c_in = 0, n_total = 0, kt = 0;
  TapTable taps;
  std::vector<uint16_t> w_hi, w_lo;   // (n_total, kt*c_in) bf16 bits (OPFMT_BF16: hi, lo)
  // OPFMT_F16F8 (common.cuh): fp16(W) bits, and the packed e5m2 plane -- per row and group of 32 K values, 64 bytes
  // [ e5m2(fp16(W) * 2^-kLoShift) x 32 | e5m2((W - fp16(W)) * 2^kHiShift) x 32 ] -- in the same (n_total, kt*c_in) x 2 B
  // footprint.  Empty when kt*c_in is not a multiple of 32.
  std::vector<uint16_t> w_h16, w_p8;
  std::vector<float> w_f32;           // same layout, fp32 (kept for the packing tests)
  std::vector<float> bias;            // (n_total)
};

// weight_norm(dim=0) fold: w[o, ...] = v[o, ...] * g[o] / ||v[o, ...]||_2
// (reference: torch.nn.utils.weight_norm as used by sparktts/modules/blocks/layers.py:24-29).
void fold_weight_norm(const float* v, const float* g, int64_t dim0, int64_t inner, std::vector<float>& w);

// Conv1d weight (C_out, C_in, k), dilation d, "same" padding (k-1)/2*d:
//   out[l, co] = sum_j sum_ci w[co, ci, j] * x[l + (j - (k-1)/2) * d, ci]
// optional per-output-row scale (ConvNeXt gamma folded into pwconv2).
void pack_conv1d(const float* w, int c_out, int c_in, int k, int dilation, const float* bias,
                 const float* row_scale, PackedGemm& out);

// ConvTranspose1d weight (C_in, C_out, k), stride s, padding (k-s)/2, as s polyphase branches:
//   out[s*q + r, co] = sum_m sum_ci w[ci, co, kk_m] * x[q + shift_m, ci],
//   kk_m = (r+p) % s + s*m,  shift_m = (r+p)/s - m      (reference: wave_generator.py:38-46)
// packed row n = r*C_out + co; output (batch, q, s*C_out) is bit-identical in memory to (batch, s*q+r, C_out).
void pack_conv_transpose1d(const float* w, int c_in, int c_out, int k, int stride, const float* bias,
                           PackedGemm& out);

}  // namespace sparkcodec
