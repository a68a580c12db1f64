// Parameters and the fused epilogue shared by the tcgen05 kernel and the CUDA-core verification kernel.
#pragma once

#include "common.cuh"
#include "op_split.cuh"

namespace sparkcodec {

struct ConvGemmParams {
  int batch, L, n_total, c_in;
  TapTable taps;
  int m_tiles_per_utt;   // ceil(L / 128)
  int num_m_tiles;       // batch * m_tiles_per_utt
  int num_n_tiles;       // n_total / BLOCK_N
  int halo_rows;         // halo mainloop: rows of the A box (128 + tap span, multiple of 8)
  int halo_bo_mode;      // halo mainloop: descriptor base-offset convention (see tc_gemm.cu)
  int halo_stages;       // halo mainloop: halo tiles in flight (2 or 3); the weight ring gets the rest of smem
  int dbg_skip_store;    // timing experiments only (SPARKCODEC_DEBUG_SKIP_STORE): skip the epilogue maths + stores
  int ring_bytes;        // tcgen05 kernel: shared memory of the operand ring(s) (what the output staging leaves)
  int xs_bytes;          // per epilogue team: fp32 output staging tile (0 when there is no fp32 output or it goes through the residual slab)
  int ps_bytes;          // per epilogue team: operand-plane staging tiles (0 / 8 KB / 16 KB)
  // epilogue
  const float* bias;
  const float* rowbias;
  const float* residual;
  const float* alpha;
  const float* inv_alpha;
  int act;
  float* out_f32;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  int out_fmt;           // OPFMT_* of out_hi / out_lo
};

int fill_params(const GemmWeights& w, int batch, int L, const Epilogue& ep, int precision, ConvGemmParams* p);
int choose_block_n(int cols_per_phase, int* block_n);
int launch_block_n(int packed_bn, int n_total, int cols_per_phase, int num_m_tiles, int num_sms);
int choose_bk(int c_in, int block_n, int precision, int ring_bytes);

// snake(x) = x + sin(alpha x)^2 / (alpha + 1e-9)   (reference sparktts/modules/blocks/layers.py:32-39)
// sin: two-constant Cody-Waite reduction to [-pi, pi], then the SFU sine (abs err ~2^-21 there).
__device__ __forceinline__ float snake_f(float x, float a, float inv) {
  float t = a * x;
#ifdef SPARKCODEC_EXACT_SIN
  float s = sinf(t);
#else
  // k = round(t / 2pi) via the 1.5 * 2^23 magic constant: two full-rate FADDs instead of a quarter-rate FRND
  // (valid for |t / 2pi| < 2^22, far beyond any activation here)
  float k = __fadd_rn(__fmaf_rn(t, 0.15915494309189535f, 12582912.0f), -12582912.0f);
  float r = fmaf(k, -6.2831854820251465f, t);
  r = fmaf(k, 1.7484555e-7f, r);
  float s = __sinf(r);
#endif
  return fmaf(inv * s, s, x);
}

// nn.GELU() default = 0.5 v (1 + erf(v / sqrt 2)) (vocos.py:57).  erf by Abramowitz-Stegun 7.1.26 (|err| <= 6e-7 in
// fp32 arithmetic): two single-instruction SFU ops (rcp.approx / ex2.approx, flush-to-zero: no denormal fix-up
// code) + 9 FMA-pipe ops, no branches, instead of erff().  The pw1 epilogue applies this to 2048 hidden
// channels per frame and is issue bound on it; the result is then split into bf16 hi/lo (2^-17), so the
// approximation error is far below what the next layer sees.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bf16 mode and the two-term fp32 mode: the SFU sine's own range handling is enough.  Its absolute error grows with
// |alpha x| (~|alpha x| 2^-23) but stays far below what the operand rounding right after it costs (2^-9 in bf16 mode,
// ~2^-15 for the e5m2 cross terms of the two-term mode): measured on the B200, the two-term waveform SNR is 73.0 dB with
// either sine (profiles/r2_two_term_ab.txt).  Only the three-term bf16 split (~2^-17 per product) keeps the reduction.
__device__ __forceinline__ float snake_fast(float x, float a, float inv) {
  const float s = __sinf(a * x);
  return fmaf(inv * s, s, x);
}
template <bool EXACT>
__device__ __forceinline__ float snake_sel(float x, float a, float inv) {
  return EXACT ? snake_f(x, a, inv) : snake_fast(x, a, inv);
}

__device__ __forceinline__ float gelu_erf(float v) {
#ifdef SPARKCODEC_EXACT_ERF
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
#else
  const float av = fabsf(v);
  const float t = rcp_approx(fmaf(av, 0.3275911f * 0.70710678118654752f, 1.0f));   // 1 / (1 + p |v| / sqrt 2)
  const float s = av * 0.84932180028801904f;                                       // sqrt(log2(e) / 2) |v|
  const float e = ex2_approx(-s * s);                                              // exp(-v^2 / 2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  const float h = 0.5f * v;
  return fmaf(h, copysignf(erf_abs, v), h);
#endif
}

// Finishes 4 consecutive output columns [n, n+4) of one output row: bias (+rowbias)(+residual),
// optional fp32 store, activation, optional bf16 hi/lo operand store.  bias/alpha/inv are the per-column
// parameters of those 4 columns (loaded once by the caller when it keeps the same columns across rows).
__device__ __forceinline__ void epilogue_store4(const ConvGemmParams& p, float4 acc, int b, size_t row_off, int n,
                                                const float4& bias, const float4& alpha, const float4& inv,
                                                bool add_residual = true) {
  acc.x += bias.x; acc.y += bias.y; acc.z += bias.z; acc.w += bias.w;
  if (p.rowbias) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(p.rowbias + (size_t)b * p.n_total + n));
    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
  }
  if (add_residual && p.residual) {
    const float4 r = *reinterpret_cast<const float4*>(p.residual + row_off + n);
    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
  }
  if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + row_off + n) = acc;
  if (p.out_hi) {
    if (p.act == ACT_SNAKE) {
      if (p.out_fmt == OPFMT_BF16 && p.out_lo != nullptr) {   // three-term fp32 mode: range-reduced sine
        acc.x = snake_f(acc.x, alpha.x, inv.x); acc.y = snake_f(acc.y, alpha.y, inv.y);
        acc.z = snake_f(acc.z, alpha.z, inv.z); acc.w = snake_f(acc.w, alpha.w, inv.w);
      } else {
        acc.x = snake_fast(acc.x, alpha.x, inv.x); acc.y = snake_fast(acc.y, alpha.y, inv.y);
        acc.z = snake_fast(acc.z, alpha.z, inv.z); acc.w = snake_fast(acc.w, alpha.w, inv.w);
      }
    } else if (p.act == ACT_GELU) {
      acc.x = gelu_erf(acc.x); acc.y = gelu_erf(acc.y); acc.z = gelu_erf(acc.z); acc.w = gelu_erf(acc.w);
    }
    store_planes4(p.out_hi, p.out_lo, p.out_fmt, row_off, n, acc.x, acc.y, acc.z, acc.w);
  }
}

__device__ __forceinline__ void load_col_params4(const ConvGemmParams& p, int n, float4& bias, float4& alpha,
                                                 float4& inv) {
  bias = __ldg(reinterpret_cast<const float4*>(p.bias + n));
  alpha = make_float4(0.f, 0.f, 0.f, 0.f);
  inv = alpha;
  if (p.out_hi && p.act == ACT_SNAKE) {
    alpha = __ldg(reinterpret_cast<const float4*>(p.alpha + n));
    inv = __ldg(reinterpret_cast<const float4*>(p.inv_alpha + n));
  }
}

// 8-column convenience wrapper (CUDA-core verification kernel).
__device__ __forceinline__ void epilogue_store8(const ConvGemmParams& p, float (&acc)[8], int b,
                                                size_t row_off /* (b*L + l) * n_total */, int n) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float4 bias, alpha, inv;
    load_col_params4(p, n + 4 * h, bias, alpha, inv);
    epilogue_store4(p, make_float4(acc[4 * h], acc[4 * h + 1], acc[4 * h + 2], acc[4 * h + 3]), b, row_off,
                    n + 4 * h, bias, alpha, inv);
  }
}

}  // namespace sparkcodec
