// Host-side weight re-layout (see pack.h).  Compiled by nvcc as host code; unit-tested on CPU
// through sparkcodec_pack_conv (tests/test_pack_cpu.py).
#include "pack.h"

#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <climits>
#include <cmath>
#include <cstring>

namespace sparkcodec {

void fold_weight_norm(const float* v, const float* g, int64_t dim0, int64_t inner, std::vector<float>& w) {
  w.resize((size_t)dim0 * inner);
  for (int64_t o = 0; o < dim0; ++o) {
    double ss = 0.0;
    const float* row = v + o * inner;
    for (int64_t i = 0; i < inner; ++i) ss += (double)row[i] * (double)row[i];
    // torch: v * (g / norm) evaluated in fp32
    float scale = g[o] / (float)std::sqrt(ss);
    for (int64_t i = 0; i < inner; ++i) w[(size_t)o * inner + i] = row[i] * scale;
  }
}

static void split_planes(PackedGemm& p) {
  size_t n = p.w_f32.size();
  p.w_hi.resize(n);
  p.w_lo.resize(n);
  for (size_t i = 0; i < n; ++i) {
    float f = p.w_f32[i];
    uint16_t hi = f32_to_bf16_rn(f);
    p.w_hi[i] = hi;
    p.w_lo[i] = f32_to_bf16_rn(f - bf16_to_f32(hi));
  }
  // two-term fp32 mode (OPFMT_F16F8)
  const size_t K = (size_t)p.kt * p.c_in;
  p.w_h16.clear();
  p.w_p8.clear();
  if (K == 0 || K % 32 != 0) return;
  p.w_h16.resize(n);
  p.w_p8.resize(n);
  uint8_t* p8 = reinterpret_cast<uint8_t*>(p.w_p8.data());
  const float s_hi = std::ldexp(1.0f, -kLoShift), s_lo = std::ldexp(1.0f, kHiShift);
  for (size_t i = 0; i < n; ++i) {
    const float f = p.w_f32[i];
    float c = f;                                           // finite saturation, like the device-side conversion
    if (c > 65504.f) c = 65504.f;
    if (c < -65504.f) c = -65504.f;
    const __half h = __float2half_rn(c);
    const float hf = __half2float(h);
    p.w_h16[i] = __half_as_ushort(h);
    const size_t row = i / K, k = i % K;
    uint8_t* q = p8 + row * K * 2 + (k >> 5) * 64 + (k & 31);
    q[0] = (uint8_t)__nv_cvt_float_to_fp8(hf * s_hi, __NV_SATFINITE, __NV_E5M2);
    q[32] = (uint8_t)__nv_cvt_float_to_fp8((f - hf) * s_lo, __NV_SATFINITE, __NV_E5M2);
  }
}

void pack_conv1d(const float* w, int c_out, int c_in, int k, int dilation, const float* bias,
                 const float* row_scale, PackedGemm& out) {
  out.c_in = c_in;
  out.n_total = c_out;
  out.kt = k;
  out.taps = TapTable();
  out.taps.n_phase = 1;
  out.taps.cols_per_phase = c_out;
  out.taps.ntaps[0] = k;
  for (int j = 0; j < k; ++j) out.taps.shift[0][j] = (j - (k - 1) / 2) * dilation;
  out.w_f32.assign((size_t)c_out * k * c_in, 0.f);
  for (int co = 0; co < c_out; ++co) {
    float sc = row_scale ? row_scale[co] : 1.f;
    for (int ci = 0; ci < c_in; ++ci)
      for (int j = 0; j < k; ++j) {
        float v = w[((size_t)co * c_in + ci) * k + j];
        out.w_f32[(size_t)co * k * c_in + (size_t)j * c_in + ci] = row_scale ? v * sc : v;
      }
  }
  out.bias.resize(c_out);
  for (int co = 0; co < c_out; ++co) {
    float b = bias ? bias[co] : 0.f;
    out.bias[co] = row_scale ? b * row_scale[co] : b;
  }
  split_planes(out);
}

void pack_conv_transpose1d(const float* w, int c_in, int c_out, int k, int stride, const float* bias,
                           PackedGemm& out) {
  const int s = stride, p = (k - s) / 2;
  out.c_in = c_in;
  out.n_total = s * c_out;
  out.taps = TapTable();
  out.taps.n_phase = s;
  out.taps.cols_per_phase = c_out;
  int kt = 0;
  for (int r = 0; r < s; ++r) {
    int base = (r + p) % s, n = 0;
    for (int kk = base; kk < k; kk += s) ++n;
    out.taps.ntaps[r] = n;
    if (n > kt) kt = n;
  }
  out.kt = kt;
  out.w_f32.assign((size_t)out.n_total * kt * c_in, 0.f);
  for (int r = 0; r < s; ++r) {
    int base = (r + p) % s, fl = (r + p) / s;
    const int nt = out.taps.ntaps[r];
    for (int m = 0; m < nt; ++m) {
      // taps stored with ASCENDING row shift (slot m holds kernel tap nt-1-m) so that the halo mainloop
      // can address them as increasing row offsets into one A tile
      int kk = base + s * (nt - 1 - m);
      out.taps.shift[r][m] = fl - (nt - 1 - m);
      for (int co = 0; co < c_out; ++co)
        for (int ci = 0; ci < c_in; ++ci)
          out.w_f32[((size_t)(r * c_out + co) * kt + m) * c_in + ci] = w[((size_t)ci * c_out + co) * k + kk];
    }
  }
  out.bias.resize(out.n_total);
  for (int r = 0; r < s; ++r)
    for (int co = 0; co < c_out; ++co) out.bias[r * c_out + co] = bias ? bias[co] : 0.f;
  split_planes(out);
}

}  // namespace sparkcodec
