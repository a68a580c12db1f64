// CUDA-core verification kernel for the dense contractions: same operands (operand planes in either format), same
// tap tables and the same fused epilogue as tc_gemm.cu, but plain FFMA: on (hi + lo) values for OPFMT_BF16, and on the
// three products the tensor cores form (hi*Whi + lo8*Whi8 + hi8*Wlo8) for OPFMT_F16F8.
// It exists to bisect the tcgen05 path in tests (sparkcodec_set_impl / sparkcodec_op_conv impl=1);
// the product path never selects it.
#include "gemm_params.cuh"

namespace sparkcodec {
namespace {

constexpr int TM = 64, TN = 32, TK = 16;

__global__ void __launch_bounds__(256)
conv_gemm_simt_kernel(const ConvGemmParams p, const __nv_bfloat16* __restrict__ a_hi,
                      const __nv_bfloat16* __restrict__ a_lo, const __nv_bfloat16* __restrict__ w_hi,
                      const __nv_bfloat16* __restrict__ w_lo, int kt, int m_tiles_per_utt, int fmt) {
  __shared__ __align__(16) float sA[TM][TK];
  __shared__ __align__(16) float sW[TN][TK];
  // OPFMT_F16F8: the e5m2 parts (still carrying their power-of-two scales, which cancel in the products)
  __shared__ __align__(16) float sAl[TM][TK], sAh[TM][TK], sWh[TN][TK], sWl[TN][TK];
  const bool f8 = fmt == OPFMT_F16F8;
  const int tid = threadIdx.x;
  const int b = blockIdx.x / m_tiles_per_utt;
  const int l0 = (blockIdx.x % m_tiles_per_utt) * TM;
  const int n0 = blockIdx.y * TN;
  const int ph = n0 / p.taps.cols_per_phase;
  const int ntaps = p.taps.ntaps[ph];
  const size_t ldw = (size_t)kt * p.c_in;

  const int r = tid >> 2;          // output row within the tile
  const int cg = (tid & 3) * 8;    // first of this thread's 8 output columns
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  // loader mapping: A tile 64 rows x 16 ch, 4 ch per thread; W tile 32 rows x 16 k, threads 0..127
  const int ar = tid >> 2, ac = (tid & 3) * 4;
  const int wr = (tid & 127) >> 2, wc = (tid & 3) * 4;

  for (int j = 0; j < ntaps; ++j) {
    const int shift = p.taps.shift[ph][j];
    for (int k0 = 0; k0 < p.c_in; k0 += TK) {
      {
        const int l = l0 + ar + shift;
        float v[4] = {0.f, 0.f, 0.f, 0.f}, vl[4] = {0.f, 0.f, 0.f, 0.f}, vh[4] = {0.f, 0.f, 0.f, 0.f};
        if (l >= 0 && l < p.L) {
          const size_t off = ((size_t)b * p.L + l) * p.c_in + k0 + ac;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (f8) {
              v[i] = __half2float(reinterpret_cast<const __half*>(a_hi)[off + i]);
              const uint8_t* q = reinterpret_cast<const uint8_t*>(a_lo) + p8_off(off - (k0 + ac), k0 + ac + i);
              vl[i] = e5m2_to_float(q[0]);
              vh[i] = e5m2_to_float(q[32]);
            } else {
              v[i] = __bfloat162float(a_hi[off + i]);
              if (a_lo) v[i] += __bfloat162float(a_lo[off + i]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { sA[ar][ac + i] = v[i]; sAl[ar][ac + i] = vl[i]; sAh[ar][ac + i] = vh[i]; }
      }
      if (tid < 128) {
        const size_t off = (size_t)(n0 + wr) * ldw + (size_t)j * p.c_in + k0 + wc;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v, vh = 0.f, vl = 0.f;
          if (f8) {
            v = __half2float(reinterpret_cast<const __half*>(w_hi)[off + i]);
            const int kk = j * p.c_in + k0 + wc + i;
            const uint8_t* q = reinterpret_cast<const uint8_t*>(w_lo) + p8_off((size_t)(n0 + wr) * ldw, kk);
            vh = e5m2_to_float(q[0]);
            vl = e5m2_to_float(q[32]);
          } else {
            v = __bfloat162float(w_hi[off + i]);
            if (w_lo) v += __bfloat162float(w_lo[off + i]);
          }
          sW[wr][wc + i] = v; sWh[wr][wc + i] = vh; sWl[wr][wc + i] = vl;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        const float a = sA[r][k];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(a, sW[cg + i][k], acc[i]);
        if (f8) {
          const float al = sAl[r][k], ah = sAh[r][k];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(al, sWh[cg + i][k], fmaf(ah, sWl[cg + i][k], acc[i]));
        }
      }
      __syncthreads();
    }
  }
  const int l = l0 + r;
  if (l < p.L) epilogue_store8(p, acc, b, ((size_t)b * p.L + l) * (size_t)p.n_total, n0 + cg);
}

}  // namespace

int launch_conv_gemm_simt(const GemmWeights& w, const OpBuf& a, int batch, int L, const Epilogue& ep, int precision,
                          cudaStream_t stream) {
  ConvGemmParams p;
  SC_TRY(fill_params(w, batch, L, ep, precision, &p));
  if (w.n_total % TN || w.taps.cols_per_phase % TN || w.c_in % TK) {
    set_error("simt gemm: unsupported shape n=%d c_in=%d", w.n_total, w.c_in);
    return SPARKCODEC_EINVAL;
  }
  const bool f32 = is_split(precision);
  const int mt = (L + TM - 1) / TM;
  dim3 grid(batch * mt, w.n_total / TN);
  const int terms = terms_for(precision);
  if (a.fmt != op_fmt_for(precision)) {
    set_error("simt gemm: operand planes are not in the format of this precision mode");
    return SPARKCODEC_EINVAL;
  }
  conv_gemm_simt_kernel<<<grid, 256, 0, stream>>>(p, a.hi, f32 ? a.lo : nullptr, w.hi_for(terms),
                                                 f32 ? w.lo_for(terms) : nullptr, w.kt, mt, a.fmt);
  SC_LAUNCH_CHECK();
  return 0;
}

}  // namespace sparkcodec
