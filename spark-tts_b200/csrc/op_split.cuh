// Device-side conversion of fp32 values into the two operand-plane formats (common.cuh: OPFMT_BF16 / OPFMT_F16F8).
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"

namespace sparkcodec {

// ---- packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 do two IEEE fp32 operations per issue slot) -----------------
// Same FMA-pipe throughput as the scalar forms (tools/micro/ffma2_bench: 72 vs 70 TFLOP/s), but half the issue slots:
// they pay where a kernel is bound by instruction issue (the waveform head, 75 % issue-active under ncu).  Each half of
// a packed op rounds exactly like the scalar op.  In the conv / fused-unit epilogues (Snake, GELU, the operand split)
// they were measured at no gain on one box (profiles/r2_packed_math_ab.txt): those epilogues wait on latency chains,
// not on issue slots, so they keep the scalar code.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ f32x2 pk2(float v) { return pk2(v, v); }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// a += b, four lanes (two packed adds)
__device__ __forceinline__ void add4(float4& a, const float4& b) {
  upk2(add2(pk2(a.x, a.y), pk2(b.x, b.y)), a.x, a.y);
  upk2(add2(pk2(a.z, a.w), pk2(b.z, b.w)), a.z, a.w);
}

// {b : upper half, a : lower half} as fp16, round to nearest, finite saturation (an fp16 inf would poison the MMAs)
__device__ __forceinline__ uint32_t cvt_f16x2(float a, float b) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
// {b : upper byte, a : lower byte} as e5m2
__device__ __forceinline__ uint32_t cvt_e5m2x2(float a, float b) {
  uint16_t d;
  asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ uint32_t cvt_e5m2x2_h2(uint32_t h2) {
  uint16_t d;
  asm("cvt.rn.satfinite.e5m2x2.f16x2 %0, %1;" : "=h"(d) : "r"(h2));
  return d;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t h2) {
  return __half22float2(*reinterpret_cast<const __half2*>(&h2));
}

// Four consecutive channels in OPFMT_F16F8: hi = two words of fp16 pairs, lo8 / hi8 = one word of four e5m2 bytes each
// (byte i = channel i).  v - hi is exact in fp32, and so is its power-of-two scaling.
struct Split4F8 { uint32_t hi[2], lo8, hi8; };
__device__ __forceinline__ Split4F8 split4_f16f8(float v0, float v1, float v2, float v3) {
  Split4F8 s;
  s.hi[0] = cvt_f16x2(v0, v1);
  s.hi[1] = cvt_f16x2(v2, v3);
  const float2 f0 = unpack_f16x2(s.hi[0]), f1 = unpack_f16x2(s.hi[1]);
  constexpr float kl = (float)(1 << kLoShift);
  s.lo8 = cvt_e5m2x2((v0 - f0.x) * kl, (v1 - f0.y) * kl) | (cvt_e5m2x2((v2 - f1.x) * kl, (v3 - f1.y) * kl) << 16);
  const __half2 sc = __float2half2_rn(1.0f / (float)(1 << kHiShift));
  const __half2 g0 = __hmul2(*reinterpret_cast<const __half2*>(&s.hi[0]), sc);
  const __half2 g1 = __hmul2(*reinterpret_cast<const __half2*>(&s.hi[1]), sc);
  s.hi8 = cvt_e5m2x2_h2(*reinterpret_cast<const uint32_t*>(&g0)) | (cvt_e5m2x2_h2(*reinterpret_cast<const uint32_t*>(&g1)) << 16);
  return s;
}
// Same for OPFMT_BF16: hi / lo = two words of bf16 pairs each (lo only when wanted)
struct Split4Bf { uint32_t hi[2], lo[2]; };
__device__ __forceinline__ Split4Bf split4_bf16(float v0, float v1, float v2, float v3, bool want_lo) {
  Split4Bf s;
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
  s.hi[0] = *reinterpret_cast<const uint32_t*>(&h0);
  s.hi[1] = *reinterpret_cast<const uint32_t*>(&h1);
  s.lo[0] = s.lo[1] = 0;
  if (want_lo) {
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v0 - f0.x, v1 - f0.y);
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(v2 - f1.x, v3 - f1.y);
    s.lo[0] = *reinterpret_cast<const uint32_t*>(&l0);
    s.lo[1] = *reinterpret_cast<const uint32_t*>(&l1);
  }
  return s;
}

// Byte offsets (inside the second plane of OPFMT_F16F8) of channel c's lo8 byte; its hi8 byte is 32 bytes further.
__device__ __forceinline__ size_t p8_off(size_t row_elems /* row * C */, int c) {
  return row_elems * 2 + (size_t)(c >> 5) * 64 + (size_t)(c & 31);
}

// Stores four consecutive channels [c, c+4) (c % 4 == 0) of one row into the operand planes `hi` / `lo` (lo may be null).
__device__ __forceinline__ void store_planes4(__nv_bfloat16* hi, __nv_bfloat16* lo, int fmt, size_t row_elems, int c,
                                              float v0, float v1, float v2, float v3) {
  if (fmt == OPFMT_F16F8) {
    const Split4F8 s = split4_f16f8(v0, v1, v2, v3);
    *reinterpret_cast<uint2*>(hi + row_elems + c) = make_uint2(s.hi[0], s.hi[1]);
    uint8_t* p = reinterpret_cast<uint8_t*>(lo) + p8_off(row_elems, c);
    *reinterpret_cast<uint32_t*>(p) = s.lo8;
    *reinterpret_cast<uint32_t*>(p + 32) = s.hi8;
  } else {
    const Split4Bf s = split4_bf16(v0, v1, v2, v3, lo != nullptr);
    *reinterpret_cast<uint2*>(hi + row_elems + c) = make_uint2(s.hi[0], s.hi[1]);
    if (lo) *reinterpret_cast<uint2*>(lo + row_elems + c) = make_uint2(s.lo[0], s.lo[1]);
  }
}

// e5m2 byte -> float (e5m2 is the upper byte of an fp16)
__device__ __forceinline__ float e5m2_to_float(uint8_t b) {
  return __half2float(__ushort_as_half((unsigned short)((unsigned short)b << 8)));
}

}  // namespace sparkcodec
