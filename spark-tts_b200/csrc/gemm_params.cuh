// Parameters and the fused epilogue shared by the tcgen05 kernel and the CUDA-core verification kernel.
#pragma once

#include "common.cuh"

namespace sparkcodec {

struct ConvGemmParams {
  int batch, L, n_total, c_in;
  TapTable taps;
  int m_tiles_per_utt;   // ceil(L / 128)
  int num_m_tiles;       // batch * m_tiles_per_utt
  int num_n_tiles;       // n_total / BLOCK_N
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
};

int fill_params(const GemmWeights& w, int batch, int L, const Epilogue& ep, int precision, ConvGemmParams* p);
int choose_tile(int c_in, int cols_per_phase, int* block_n, int* bk);

// snake(x) = x + sin(alpha x)^2 / (alpha + 1e-9)   (reference sparktts/modules/blocks/layers.py:32-39)
// sin: two-constant Cody-Waite reduction to [-pi, pi], then the SFU sine (abs err ~2^-21 there).
__device__ __forceinline__ float snake_f(float x, float a, float inv) {
  float t = a * x;
#ifdef SPARKCODEC_EXACT_SIN
  float s = sinf(t);
#else
  float k = rintf(t * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, t);
  r = fmaf(k, 1.7484555e-7f, r);
  float s = __sinf(r);
#endif
  return fmaf(inv * s, s, x);
}

__device__ __forceinline__ float gelu_erf(float v) {   // nn.GELU() default (vocos.py:57)
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
}

// Finishes 8 consecutive output columns [n, n+8) of one output row: bias (+rowbias)(+residual),
// optional fp32 store, activation, optional bf16 hi/lo operand store.  acc[] holds the accumulators.
__device__ __forceinline__ void epilogue_store8(const ConvGemmParams& p, float (&acc)[8], int b,
                                                size_t row_off /* (b*L + l) * n_total */, int n) {
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
  acc[0] += b0.x; acc[1] += b0.y; acc[2] += b0.z; acc[3] += b0.w;
  acc[4] += b1.x; acc[5] += b1.y; acc[6] += b1.z; acc[7] += b1.w;
  if (p.rowbias) {
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(p.rowbias + (size_t)b * p.n_total + n));
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(p.rowbias + (size_t)b * p.n_total + n + 4));
    acc[0] += r0.x; acc[1] += r0.y; acc[2] += r0.z; acc[3] += r0.w;
    acc[4] += r1.x; acc[5] += r1.y; acc[6] += r1.z; acc[7] += r1.w;
  }
  if (p.residual) {
    const float4 r0 = *reinterpret_cast<const float4*>(p.residual + row_off + n);
    const float4 r1 = *reinterpret_cast<const float4*>(p.residual + row_off + n + 4);
    acc[0] += r0.x; acc[1] += r0.y; acc[2] += r0.z; acc[3] += r0.w;
    acc[4] += r1.x; acc[5] += r1.y; acc[6] += r1.z; acc[7] += r1.w;
  }
  if (p.out_f32) {
    *reinterpret_cast<float4*>(p.out_f32 + row_off + n) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(p.out_f32 + row_off + n + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  if (p.out_hi) {
    if (p.act == ACT_SNAKE) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.alpha + n));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.alpha + n + 4));
      const float4 i0 = __ldg(reinterpret_cast<const float4*>(p.inv_alpha + n));
      const float4 i1 = __ldg(reinterpret_cast<const float4*>(p.inv_alpha + n + 4));
      acc[0] = snake_f(acc[0], a0.x, i0.x); acc[1] = snake_f(acc[1], a0.y, i0.y);
      acc[2] = snake_f(acc[2], a0.z, i0.z); acc[3] = snake_f(acc[3], a0.w, i0.w);
      acc[4] = snake_f(acc[4], a1.x, i1.x); acc[5] = snake_f(acc[5], a1.y, i1.y);
      acc[6] = snake_f(acc[6], a1.z, i1.z); acc[7] = snake_f(acc[7], a1.w, i1.w);
    } else if (p.act == ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = gelu_erf(acc[i]);
    }
    __nv_bfloat162 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
    *reinterpret_cast<uint4*>(p.out_hi + row_off + n) = *reinterpret_cast<uint4*>(h);
    if (p.out_lo) {
      __nv_bfloat162 l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 hf = __bfloat1622float2(h[i]);
        l[i] = __floats2bfloat162_rn(acc[2 * i] - hf.x, acc[2 * i + 1] - hf.y);
      }
      *reinterpret_cast<uint4*>(p.out_lo + row_off + n) = *reinterpret_cast<uint4*>(l);
    }
  }
}

}  // namespace sparkcodec
