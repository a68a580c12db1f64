// Kernels of the speaker half of BiCodec.tokenize (reference sparktts/models/bicodec.py:162-167):
//   mel spectrogram (torchaudio MelSpectrogram as built by bicodec.py:191-211)  -> spk_frames / spk_gemm / spk_magnitude
//   ECAPA-TDNN trunk up to its `latent` output (speaker/ecapa_tdnn.py:28-214)  -> spk_gemm (conv + ReLU + folded BatchNorm),
//                                                                                  spk_mean_rows, spk_se_apply
//   perceiver resampler (speaker/perceiver_encoder.py:254-350)                  -> spk_gemm, spk_attention, spk_geglu, spk_rmsnorm
//   FSQ quantise (fsq/finite_scalar_quantization.py:101-141, residual_fsq.py:213-283) -> spk_fsq_quantize
// Once per prompt and ~3 GFLOP per 6 s clip, against 590 GFLOP per 10 s of decoded audio: this path is written for
// fidelity (full fp32 FMA arithmetic, so that the integer tokens equal the reference's away from rounding boundaries),
// not for the tensor cores.  All activations are channels-last fp32 (batch, rows, channels).
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace sparkcodec {
namespace {

// frames[b, t, i] = wav[b, reflect(t * hop - win / 2 + i)], i in [0, win): the win_length samples under the window of
// the centred STFT frame t (torch.stft(center=True, pad_mode="reflect") pads n_fft / 2 on both sides and zero-pads the
// window to n_fft around its centre, so only these samples meet a non-zero window value).
__global__ void spk_frames_kernel(const float* __restrict__ wav, int n, int frames, int hop, int win,
                                  float* __restrict__ out, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % win);
    const size_t bt = i / win;
    const int t = (int)(bt % frames);
    const size_t b = bt / frames;
    long long s = (long long)t * hop - win / 2 + k;
    if (s < 0) s = -s;
    if (s >= n) s = 2LL * (n - 1) - s;
    s = s < 0 ? 0 : (s >= n ? n - 1 : s);      // (clips shorter than the padding: clamp instead of reading outside)
    out[i] = __ldg(wav + b * (size_t)n + s);
  }
}

// |re + i im| of the (rows, 2 * bins) DFT output [re | im]
__global__ void spk_magnitude_kernel(const float* __restrict__ spec, int bins, float* __restrict__ out, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / bins;
    const int k = (int)(i % bins);
    const float re = spec[row * 2 * bins + k], im = spec[row * 2 * bins + bins + k];
    out[i] = hypotf(re, im);
  }
}

// Generic fp32 convolution / linear layer as a tiled FFMA GEMM (64 x 64 tile, 16-wide K steps, 4 x 4 per thread):
//   y[b, r, n] = post( sum_j sum_c (x [+ x2])[b, r + shift_j, c] * w[n, j * K + c] + bias[n] )      rows outside [0, rows) = 0
//   post(v) = act( affine( relu?(v) ) ) (+ res[b, r, n])
template <bool HAS_X2>
__global__ void __launch_bounds__(256) spk_gemm_kernel(SpkGemm g) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int M = g.batch * g.rows;
  float acc[4][4] = {};
  // loader mapping: 4 elements per thread of each 64 x 16 tile: row = tid / 4, k = (tid % 4) * 4 + 0..3
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int gm = m0 + lr;
  const int b = gm < M ? gm / g.rows : 0, r = gm < M ? gm % g.rows : 0;
  const int gn = n0 + lr;
  for (int j = 0; j < g.ntaps; ++j) {
    const int rr = r + g.shift[j];
    const bool row_ok = gm < M && rr >= 0 && rr < g.rows;
    const size_t xrow = ((size_t)b * g.rows + (row_ok ? rr : 0));
    for (int k0 = 0; k0 < g.K; k0 += 16) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kk = k0 + lk + q;
        float a = 0.f, w = 0.f;
        if (row_ok && kk < g.K) {
          a = __ldg(g.x + xrow * g.ldx + kk);
          if (HAS_X2) a += __ldg(g.x2 + xrow * g.ldx2 + kk);
        }
        if (gn < g.N && kk < g.K) w = __ldg(g.w + (size_t)gn * g.ldw + (size_t)j * g.K + kk);
        As[lk + q][lr] = a;
        Ws[lk + q][lr] = w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][q] = fmaf(av[i], wv[q], acc[i][q]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int n = n0 + tx * 4 + q;
      if (n >= g.N) continue;
      float v = acc[i][q];
      if (g.bias) v += __ldg(g.bias + n);
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.scale) v = fmaf(v, __ldg(g.scale + n), __ldg(g.shift_v + n));
      if (g.act == 1) v = 1.0f / (1.0f + expf(-v));
      if (g.res) v += g.res[(size_t)m * g.ldres + n];
      g.y[(size_t)m * g.ldy + n] = v;
    }
  }
}

// mean over rows: x (batch, rows, C) with row stride ldx -> out (batch, C)      (SE_Connect: x.mean(dim=2))
__global__ void spk_mean_rows_kernel(const float* __restrict__ x, int rows, int C, int ldx, float* __restrict__ out) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += x[((size_t)b * rows + r) * ldx + c];
  out[(size_t)b * C + c] = s / (float)rows;
}

// out[b, r, c] = x[b, r, c] + u[b, r, c] * s[b, c]                              (SE_Res2Block: x + SE(...))
__global__ void spk_se_apply_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ u, int ldu,
                                    const float* __restrict__ s, int rows, int C, float* __restrict__ out, int ldo,
                                    size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t row = i / C, b = row / rows;
    out[row * ldo + c] = x[row * ldx + c] + u[row * ldu + c] * s[b * C + c];
  }
}

// strided 2-D copy: dst[row, 0..cols) = src[row, 0..cols)
__global__ void spk_copy_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int cols,
                                     size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / cols;
    const int c = (int)(i % cols);
    dst[row * ldd + c] = src[row * lds + c];
  }
}

// rows of `src` (batch, src_rows, C) -> rows [row_off, row_off + src_rows) of dst (batch, dst_rows, C); a batch stride of
// 0 broadcasts one (src_rows, C) block to every utterance (the perceiver's learned latents)
__global__ void spk_place_rows_kernel(const float* __restrict__ src, size_t src_batch_stride, int src_rows, int C,
                                      float* __restrict__ dst, int dst_rows, int row_off, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t rr = i / C;
    const int r = (int)(rr % src_rows);
    const size_t b = rr / src_rows;
    dst[(b * dst_rows + row_off + r) * C + c] = src[b * src_batch_stride + (size_t)r * C + c];
  }
}

// Cross attention of the perceiver (perceiver_encoder.py:137-177, 254-294): one warp per (utterance, head, query).
//   q (batch, nq, heads * 64), kv (batch, nk, 2 * heads * 64) = [k | v], out (batch, nq, heads * 64)
// sim = q . k * 64^-0.5 -> softmax over the nk keys -> weighted sum of v.  Scores are kept in shared memory so the
// softmax is the plain three-pass form (max, exp / sum, normalise) of the reference.
constexpr int kAttWarps = 4;
__global__ void __launch_bounds__(kAttWarps * 32) spk_attention_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                                       int nq, int nk, int heads, float scale,
                                                                       float* __restrict__ out, int total_warps) {
  extern __shared__ float s_att[];                      // [kAttWarps][nk]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int wid = blockIdx.x * kAttWarps + wib;
  if (wid >= total_warps) return;
  const int i = wid % nq, h = (wid / nq) % heads, b = wid / (nq * heads);
  const int dq = heads * 64;
  float* sc = s_att + (size_t)wib * nk;
  const float* qp = q + ((size_t)b * nq + i) * dq + h * 64;
  const float q0 = qp[lane], q1 = qp[lane + 32];
  const float* kb = kv + (size_t)b * nk * 2 * dq + h * 64;
  float mx = -INFINITY;
  for (int j = 0; j < nk; ++j) {
    const float* kp = kb + (size_t)j * 2 * dq;
    float d = fmaf(q0, kp[lane], q1 * kp[lane + 32]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    d *= scale;
    if (lane == 0) sc[j] = d;
    mx = fmaxf(mx, d);
  }
  __syncwarp();
  float sum = 0.f;
  for (int j = lane; j < nk; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncwarp();
  const float inv = 1.0f / sum;
  float a0 = 0.f, a1 = 0.f;
  const float* vb = kb + dq;
  for (int j = 0; j < nk; ++j) {
    const float p = sc[j] * inv;
    const float* vp = vb + (size_t)j * 2 * dq;
    a0 = fmaf(p, vp[lane], a0);
    a1 = fmaf(p, vp[lane + 32], a1);
  }
  float* op = out + ((size_t)b * nq + i) * dq + h * 64;
  op[lane] = a0;
  op[lane + 32] = a1;
}

// GEGLU (perceiver_encoder.py:225-228): h (rows, 2 * inner) = [x | gate] -> out (rows, inner) = gelu(gate) * x, erf GELU
__global__ void spk_geglu_kernel(const float* __restrict__ h, int inner, float* __restrict__ out, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / inner;
    const int c = (int)(i % inner);
    const float x = h[row * 2 * inner + c], gate = h[row * 2 * inner + inner + c];
    out[i] = 0.5f * gate * (1.0f + erff(gate * 0.70710678118654752f)) * x;
  }
}

// RMSNorm (perceiver_encoder.py:190-207): F.normalize(x, dim=-1) * sqrt(dim) * gamma.  One warp per row.
__global__ void spk_rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, int dim, int rows,
                                   float* __restrict__ out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < dim; c += 32) { const float v = x[(size_t)row * dim + c]; s = fmaf(v, v, s); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float k = sqrtf((float)dim) / fmaxf(sqrtf(s), 1e-12f);
  for (int c = lane; c < dim; c += 32) out[(size_t)row * dim + c] = x[(size_t)row * dim + c] * k * __ldg(gamma + c);
}

// FSQ quantise of one latent row (one warp per row): z = project_in(x); bounded = tanh(z + shift) * half_l - offset;
// code level = round(bounded) + L / 2; index = sum level_j * basis_j          (finite_scalar_quantization.py:101-141)
// margin = distance of the closest coordinate to a rounding boundary (tests use it to tell a near-tie from an error).
__global__ void spk_fsq_quantize_kernel(const float* __restrict__ x, int dim, int rows, const float* __restrict__ w_in,
                                        const float* __restrict__ b_in, const int* __restrict__ levels, int n_levels,
                                        int* __restrict__ idx_out, float* __restrict__ margin_out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  int index = 0, basis = 1;
  float margin = 1.0f;
  for (int j = 0; j < n_levels; ++j) {
    float s = 0.f;
    for (int c = lane; c < dim; c += 32) s = fmaf(x[(size_t)row * dim + c], __ldg(w_in + (size_t)j * dim + c), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float z = s + __ldg(b_in + j);
    const int L = levels[j];
    const float half_l = (float)(L - 1) * (1.0f + 1e-3f) / 2.0f;
    const float offset = (L % 2 == 0) ? 0.5f : 0.0f;
    const float shift = atanhf(offset / half_l);
    const float bounded = tanhf(z + shift) * half_l - offset;
    const float q = rintf(bounded);                       // torch.round: half to even
    index += ((int)q + L / 2) * basis;
    basis *= L;
    margin = fminf(margin, fabsf((bounded - floorf(bounded)) - 0.5f));
  }
  if (lane == 0) {
    idx_out[row] = index;
    if (margin_out) margin_out[row] = margin;
  }
}

inline int grid_for(size_t total) { return (int)std::min<size_t>((total + 255) / 256, 148 * 32); }

}  // namespace

int launch_spk_frames(const float* wav, int batch, int n, int frames, int hop, int win, float* out, cudaStream_t s) {
  const size_t total = (size_t)batch * frames * win;
  spk_frames_kernel<<<grid_for(total), 256, 0, s>>>(wav, n, frames, hop, win, out, total);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_spk_magnitude(const float* spec, size_t rows, int bins, float* out, cudaStream_t s) {
  const size_t total = rows * bins;
  spk_magnitude_kernel<<<grid_for(total), 256, 0, s>>>(spec, bins, out, total);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_spk_gemm(const SpkGemm& g, cudaStream_t s) {
  if (g.ntaps < 1 || g.ntaps > 8) { set_error("spk_gemm: 1..8 taps"); return SPARKCODEC_EINVAL; }
  const int M = g.batch * g.rows;
  if (M == 0 || g.N == 0) return 0;
  dim3 grid((g.N + 63) / 64, (M + 63) / 64);
  if (g.x2) spk_gemm_kernel<true><<<grid, 256, 0, s>>>(g);
  else spk_gemm_kernel<false><<<grid, 256, 0, s>>>(g);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_spk_mean_rows(const float* x, int batch, int rows, int C, int ldx, float* out, cudaStream_t s) {
  spk_mean_rows_kernel<<<dim3((C + 127) / 128, batch), 128, 0, s>>>(x, rows, C, ldx, out);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_spk_se_apply(const float* x, int ldx, const float* u, int ldu, const float* sc, int batch, int rows, int C,
                        float* out, int ldo, cudaStream_t s) {
  const size_t total = (size_t)batch * rows * C;
  spk_se_apply_kernel<<<grid_for(total), 256, 0, s>>>(x, ldx, u, ldu, sc, rows, C, out, ldo, total);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_spk_copy_cols(const float* src, int lds, float* dst, int ldd, size_t rows, int cols, cudaStream_t s) {
  const size_t total = rows * cols;
  spk_copy_cols_kernel<<<grid_for(total), 256, 0, s>>>(src, lds, dst, ldd, cols, total);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_spk_place_rows(const float* src, size_t src_batch_stride, int batch, int src_rows, int C, float* dst,
                          int dst_rows, int row_off, cudaStream_t s) {
  const size_t total = (size_t)batch * src_rows * C;
  spk_place_rows_kernel<<<grid_for(total), 256, 0, s>>>(src, src_batch_stride, src_rows, C, dst, dst_rows, row_off, total);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_spk_attention(const float* q, const float* kv, int batch, int nq, int nk, int heads, int dim_head, float* out,
                         cudaStream_t s) {
  if (dim_head != 64) { set_error("spk_attention: dim_head must be 64"); return SPARKCODEC_EINVAL; }
  const size_t smem = (size_t)kAttWarps * nk * sizeof(float);
  if (smem > 200 * 1024) { set_error("spk_attention: %d keys do not fit shared memory (reference clips are 6 s = 333 keys)", nk); return SPARKCODEC_EINVAL; }
  static PerDevice done;
  if (smem > 48 * 1024 && done.here().load(std::memory_order_relaxed) < (int)smem) {
    SC_CUDA(cudaFuncSetAttribute(spk_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    done.here().store(200 * 1024, std::memory_order_relaxed);
  }
  const int warps = batch * heads * nq;
  spk_attention_kernel<<<(warps + kAttWarps - 1) / kAttWarps, kAttWarps * 32, smem, s>>>(q, kv, nq, nk, heads,
                                                                                          1.0f / sqrtf((float)dim_head), out, warps);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_spk_geglu(const float* h, size_t rows, int inner, float* out, cudaStream_t s) {
  const size_t total = rows * inner;
  spk_geglu_kernel<<<grid_for(total), 256, 0, s>>>(h, inner, out, total);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_spk_rmsnorm(const float* x, const float* gamma, int dim, int rows, float* out, cudaStream_t s) {
  spk_rmsnorm_kernel<<<(rows * 32 + 127) / 128, 128, 0, s>>>(x, gamma, dim, rows, out);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_spk_fsq_quantize(const float* x, int dim, int rows, const float* w_in, const float* b_in, const int* levels,
                            int n_levels, int* idx_out, float* margin_out, cudaStream_t s) {
  spk_fsq_quantize_kernel<<<(rows * 32 + 127) / 128, 128, 0, s>>>(x, dim, rows, w_in, b_in, levels, n_levels, idx_out, margin_out);
  SC_LAUNCH_CHECK();
  return 0;
}

}  // namespace sparkcodec
