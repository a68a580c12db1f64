// HBM-streaming and token kernels of the BiCodec detokenize path (sm_100a):
//   * split / merge            fp32 <-> bf16 hi/lo operand planes
//   * vq_embed / vq_zq         FactorizedVectorQuantize.detokenize (+ folded linear_pre, x3)
//   * fsq_project              FSQ index -> level codes -> project_out -> (c*N + n) flattening
//   * small_linear             tiny per-utterance Linear layers (speaker project, AdaLN scale/shift)
//   * dwconv_ln                depthwise k=7 conv + LayerNorm / AdaLayerNorm, smem halo staging
//   * head                     Snake -> Conv1d(C->1, k=7) -> tanh, warp-shuffle channel reduction
#include <algorithm>

#include "common.cuh"
#include "gemm_params.cuh"
#include "tc_ptx.cuh"

namespace sparkcodec {
namespace {

// four consecutive channels at flat element index idx (idx % 4 == 0; every channel count here is a multiple of 32, so
// the 32-channel groups of OPFMT_F16F8 are also groups of the flat index)
__device__ __forceinline__ void store_op4(const OpBuf& o, size_t idx, float4 v) {
  store_planes4(o.hi, o.lo, o.fmt, idx & ~(size_t)31, (int)(idx & 31), v.x, v.y, v.z, v.w);
}

// ------------------------------------------------------------------------------- split / merge
__global__ void split_kernel(const float4* __restrict__ x, OpBuf out, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    store_op4(out, i * 4, __ldg(x + i));
}
__global__ void split_rows_kernel(const float* __restrict__ x, size_t src_batch_stride, OpBuf out, int rows, int c4,
                                  int rows_total, int row_off, size_t n4) {
  const size_t per_b = (size_t)rows * c4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / per_b, r = i % per_b;     // r = row * c4 + quad
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + b * src_batch_stride) + r);
    store_op4(out, ((b * rows_total + row_off) * (size_t)c4 + r) * 4, v);
  }
}
__global__ void merge_kernel(OpBuf in, float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v;
    if (in.fmt == OPFMT_F16F8) {
      v = __half2float(reinterpret_cast<const __half*>(in.hi)[i]);
      v += e5m2_to_float(reinterpret_cast<const uint8_t*>(in.lo)[(i >> 5) * 64 + (i & 31)]) * (1.0f / (float)(1 << kLoShift));
    } else {
      v = __bfloat162float(in.hi[i]);
      if (in.lo) v += __bfloat162float(in.lo[i]);
    }
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------ tokens
__device__ __forceinline__ long long load_token(const void* p, int dtype, size_t i) {
  return dtype == SPARKCODEC_I64 ? static_cast<const long long*>(p)[i]
                                 : (long long)static_cast<const int*>(p)[i];
}
// err[0] = 1 + which (0 semantic, 1 global), err[1] = flat position, err[2..3] = value
__device__ __forceinline__ void report_bad_token(int* err, int which, size_t pos, long long v) {
  if (atomicCAS(err, 0, 1 + which) == 0) {
    err[1] = (int)pos;
    err[2] = (int)(v & 0xffffffffll);
    err[3] = (int)(v >> 32);
  }
}

// THE codebook gather (factorized_vector_quantize.py:163-164 embed_code = F.embedding(idx, codebook)): pointer to row
// `id` of the (codebook_size, cdim) table.  vq_embed_kernel (product), vq_zq_kernel and the `codebook_rows` tap all
// go through this function and through load_token, so the bit-exact tap checks the product's own indexing.
__device__ __forceinline__ const float* codebook_row(const float* __restrict__ codebook, long long id, int cdim) {
  return codebook + (size_t)id * cdim;
}
// tap: rows[tok, :] = codebook[idx[tok], :]
__global__ void vq_rows_kernel(const void* __restrict__ sem, int sem_dtype, size_t n_tok, int codebook_size, int cdim,
                               const float* __restrict__ codebook, float* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n_tok * cdim) return;
  const size_t tok = i / cdim;
  long long id = load_token(sem, sem_dtype, tok);
  if (id < 0 || id >= codebook_size) id = 0;
  out[i] = __ldg(codebook_row(codebook, id, cdim) + (int)(i % cdim));
}

// x0[tok, c] = mat[c, :] . codebook[idx[tok], :] + vec[c]   (mat/vec fold out_project, linear_pre and
// the first SamplingBlock's x3: a purely linear chain, factorized_vector_quantize.py:157 ->
// feat_decoder.py:87 -> samper.py:98).  One thread per (token, 4 channels).
__global__ void vq_embed_kernel(const void* __restrict__ sem, int sem_dtype, size_t n_tok, int codebook_size,
                                int cdim, const float* __restrict__ codebook, const float* __restrict__ mat,
                                const float* __restrict__ vec, int c_out, OpBuf out, int* err) {
  const int groups = c_out / 4;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n_tok * groups) return;
  const size_t tok = i / groups;
  const int c = (int)(i % groups) * 4;
  long long id = load_token(sem, sem_dtype, tok);
  if (id < 0 || id >= codebook_size) {
    if (c == 0) report_bad_token(err, 0, tok, id);
    id = 0;
  }
  const float* e = codebook_row(codebook, id, cdim);
  float acc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float a = __ldg(vec + c + q);
    for (int j = 0; j < cdim; ++j) a = fmaf(__ldg(mat + (size_t)(c + q) * cdim + j), __ldg(e + j), a);
    acc[q] = a;
  }
  store_op4(out, tok * c_out + c, make_float4(acc[0], acc[1], acc[2], acc[3]));
}

// z_q[tok, :] = W_out . codebook[idx] + b_out (test tap; factorized_vector_quantize.py:154-158)
__global__ void vq_zq_kernel(const void* __restrict__ sem, int sem_dtype, size_t n_tok, int codebook_size, int cdim,
                             const float* __restrict__ codebook, const float* __restrict__ w,
                             const float* __restrict__ bias, int d_model, float* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n_tok * d_model) return;
  const size_t tok = i / d_model;
  const int c = (int)(i % d_model);
  long long id = load_token(sem, sem_dtype, tok);
  if (id < 0 || id >= codebook_size) id = 0;
  const float* e = codebook_row(codebook, id, cdim);
  // same association order as a 1x1 conv over 8 input channels: bias added last
  float a = 0.f;
  for (int j = 0; j < cdim; ++j) a = fmaf(__ldg(w + (size_t)c * cdim + j), __ldg(e + j), a);
  out[i] = a + __ldg(bias + c);
}

// Semantic tokenize tail (encode side): z_e = in_project(project(x)) folded into one (cdim x C)
// map, F.normalize, then the nearest normalised code: argmin_k |c_k|^2 - 2 e.c_k, lowest index on ties
// (factorized_vector_quantize.py:147-152, 169-187: dist = |e|^2 - 2 e c^T + |c|^2, indices = (-dist).max(1)[1]).
// One warp owns kVqFrames frames: the lanes split the channels for the projection and the codes for the search.
// The 8 warps of a CTA walk the table together in kVqChunk-code pieces staged in shared memory (two 16 B planes
// per code so that consecutive lanes read consecutive 16 B: conflict-free), so every table byte is fetched from
// L2 once per 64 frames instead of once per 8.  Also returns the margin between the best and second-best score so
// that callers can tell a real disagreement from a numerical near-tie.
constexpr int kVqFrames = 8;
constexpr int kVqDim = 8;
constexpr int kVqChunk = 512;
__global__ void __launch_bounds__(256) vq_search_kernel(const float* __restrict__ x, size_t n_frames, int c,
                                                        const float* __restrict__ mat, const float* __restrict__ vec,
                                                        const float* __restrict__ codes_n,
                                                        const float* __restrict__ codes_sq, int codebook_size,
                                                        long long* __restrict__ idx_out,
                                                        float* __restrict__ margin_out) {
  __shared__ float4 s_c0[kVqChunk], s_c1[kVqChunk];
  __shared__ float s_sq[kVqChunk];
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t f0 = warp * kVqFrames;      // may be >= n_frames in the last CTA: such warps only help with staging
  float e[kVqFrames][kVqDim];
#pragma unroll
  for (int f = 0; f < kVqFrames; ++f) {
    const size_t fr = f0 + f < n_frames ? f0 + f : n_frames - 1;
    float acc[kVqDim];
#pragma unroll
    for (int j = 0; j < kVqDim; ++j) acc[j] = 0.f;
    for (int ch = lane; ch < c; ch += 32) {
      const float v = __ldg(x + fr * c + ch);
#pragma unroll
      for (int j = 0; j < kVqDim; ++j) acc[j] = fmaf(__ldg(mat + (size_t)j * c + ch), v, acc[j]);
    }
    float nrm = 0.f;
#pragma unroll
    for (int j = 0; j < kVqDim; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      acc[j] += __ldg(vec + j);
      nrm = fmaf(acc[j], acc[j], nrm);
    }
    const float inv = 1.0f / fmaxf(sqrtf(nrm), 1e-12f);   // F.normalize: x / max(||x||, eps)
#pragma unroll
    for (int j = 0; j < kVqDim; ++j) e[f][j] = acc[j] * inv;
  }
  float best[kVqFrames], second[kVqFrames];
  int arg[kVqFrames];
#pragma unroll
  for (int f = 0; f < kVqFrames; ++f) { best[f] = INFINITY; second[f] = INFINITY; arg[f] = 0x7fffffff; }
  for (int k0 = 0; k0 < codebook_size; k0 += kVqChunk) {
    const int nk = min(kVqChunk, codebook_size - k0);
    __syncthreads();   // the previous chunk has been consumed
    for (int i = threadIdx.x; i < nk; i += blockDim.x) {
      s_c0[i] = __ldg(reinterpret_cast<const float4*>(codes_n + (size_t)(k0 + i) * kVqDim));
      s_c1[i] = __ldg(reinterpret_cast<const float4*>(codes_n + (size_t)(k0 + i) * kVqDim + 4));
      s_sq[i] = __ldg(codes_sq + k0 + i);
    }
    __syncthreads();
    for (int i = lane; i < nk; i += 32) {   // ascending code index per lane: `<` keeps the lowest index on ties
      const float4 c0 = s_c0[i], c1 = s_c1[i];
      const float sq = s_sq[i];
#pragma unroll
      for (int f = 0; f < kVqFrames; ++f) {
        float d = e[f][0] * c0.x;
        d = fmaf(e[f][1], c0.y, d); d = fmaf(e[f][2], c0.z, d); d = fmaf(e[f][3], c0.w, d);
        d = fmaf(e[f][4], c1.x, d); d = fmaf(e[f][5], c1.y, d); d = fmaf(e[f][6], c1.z, d); d = fmaf(e[f][7], c1.w, d);
        const float sc = fmaf(-2.0f, d, sq);
        if (sc < best[f]) { second[f] = best[f]; best[f] = sc; arg[f] = k0 + i; }
        else second[f] = fminf(second[f], sc);
      }
    }
  }
#pragma unroll
  for (int f = 0; f < kVqFrames; ++f) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best[f], o), os = __shfl_xor_sync(0xffffffffu, second[f], o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg[f], o);
      if (ob < best[f] || (ob == best[f] && oa < arg[f])) {
        second[f] = fminf(best[f], os);
        best[f] = ob; arg[f] = oa;
      } else {
        second[f] = fminf(second[f], ob);
      }
    }
    if (lane == 0 && f0 + f < n_frames) {
      idx_out[f0 + f] = arg[f];
      if (margin_out) margin_out[f0 + f] = second[f] - best[f];
    }
  }
}

// FSQ: level_j = (idx / basis_j) % L_j, code_j = (level_j - L_j/2) / (L_j/2)   (exact in fp32)
// z[c] = W_po[c,:] . code + b_po[c];  flat[b, c*N + n] = z[c]   (finite_scalar_quantization.py:143-162,
// residual_fsq.py:191-199, speaker_encoder.py:107-111).  One block per (b, n), one thread per c.
// THE FSQ index decode, shared by the product kernel and the `fsq_codes` tap.
__device__ __forceinline__ void fsq_decode(long long id, const int* __restrict__ levels, int n_levels, float (&code)[8]) {
  long long basis = 1;
  for (int j = 0; j < n_levels; ++j) {
    const int L = levels[j], half = L / 2;
    const int level = (int)((id / basis) % L);
    code[j] = (float)(level - half) / (float)half;
    basis *= L;
  }
}
__device__ __forceinline__ long long fsq_total(const int* __restrict__ levels, int n_levels) {
  long long total = 1;
  for (int j = 0; j < n_levels; ++j) total *= levels[j];
  return total;
}
// tap: codes[b, n, j] (residual_fsq.py get_codes_from_indices, one quantizer, scale 1)
__global__ void fsq_codes_kernel(const void* __restrict__ glob, int glob_dtype, int n_tok, int n_levels,
                                 const int* __restrict__ levels, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tok) return;
  long long id = load_token(glob, glob_dtype, (size_t)i);
  if (id < 0 || id >= fsq_total(levels, n_levels)) id = 0;
  float code[8];
  fsq_decode(id, levels, n_levels, code);
  for (int j = 0; j < n_levels; ++j) out[(size_t)i * n_levels + j] = code[j];
}

__global__ void fsq_project_kernel(const void* __restrict__ glob, int glob_dtype, int token_num, int n_levels,
                                   const int* __restrict__ levels, const float* __restrict__ w_po,
                                   const float* __restrict__ b_po, int latent, float* __restrict__ flat, int* err) {
  const int bn = blockIdx.x, b = bn / token_num, n = bn % token_num;
  long long id = load_token(glob, glob_dtype, (size_t)bn);
  if (id < 0 || id >= fsq_total(levels, n_levels)) {
    if (threadIdx.x == 0) report_bad_token(err, 1, (size_t)bn, id);
    id = 0;
  }
  float code[8];
  fsq_decode(id, levels, n_levels, code);
  for (int c = threadIdx.x; c < latent; c += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < n_levels; ++j) a = fmaf(__ldg(w_po + c * n_levels + j), code[j], a);
    flat[(size_t)b * latent * token_num + (size_t)c * token_num + n] = a + __ldg(b_po + c);
  }
}

// y[b, n] = x[b, :] . w[n, :] + bias[n].  One warp per output column n, 8 utterances per pass, so W is
// streamed from HBM once per launch (the batch re-reads of its 4-16 KB row hit L1).
__global__ void small_linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ y, int batch, int k, int n) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float4* wr = reinterpret_cast<const float4*>(w + (size_t)warp * k);
  const int k4 = k / 4;
  for (int b0 = 0; b0 < batch; b0 += 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = lane; i < k4; i += 32) {
      const float4 wv = __ldg(wr + i);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (b0 + q < batch) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)(b0 + q) * k) + i);
          acc[q] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[q]))));
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v = acc[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && b0 + q < batch) y[(size_t)(b0 + q) * n + warp] = v + __ldg(bias + warp);
    }
  }
}

// --------------------------------------------------------------------------------- token feed
// LLM output ids -> BiCodec codes without a host round trip.  The reference decodes the generated ids to
// text and regex-matches "bicodec_semantic_(\d+)" / "bicodec_global_(\d+)" (cli/SparkTTS.py:213-228,
// runtime/triton_trtllm/model_repo/spark_tts/1/model.py:283-295); in the tokenizer those are contiguous
// added-token id ranges, so the same selection is an ORDER-PRESERVING compaction of the ids that fall into
// [semantic_base, semantic_base + codebook) resp. [global_base, global_base + global_size).
// One CTA per utterance; 256 ids per round, warp ballots + a block prefix over the 8 warp totals.
__global__ void __launch_bounds__(256)
extract_codes_kernel(const void* __restrict__ ids, int id_dtype, int n_tokens, long long sem_base, int sem_size,
                     long long glob_base, int glob_size, int* __restrict__ sem_out, int* __restrict__ sem_len,
                     int* __restrict__ glob_out, int max_global, int* __restrict__ glob_len) {
  __shared__ int s_cnt[2][8];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int run[2] = {0, 0};     // codes written so far (same value in every thread)
  for (int t0 = 0; t0 < n_tokens; t0 += 256) {
    const int t = t0 + threadIdx.x;
    long long id = -1;
    if (t < n_tokens) id = load_token(ids, id_dtype, (size_t)b * n_tokens + t);
    const bool is_s = id >= sem_base && id < sem_base + sem_size;
    const bool is_g = id >= glob_base && id < glob_base + glob_size;
    const unsigned ms = __ballot_sync(0xffffffffu, is_s), mg = __ballot_sync(0xffffffffu, is_g);
    if (lane == 0) { s_cnt[0][warp] = __popc(ms); s_cnt[1][warp] = __popc(mg); }
    __syncthreads();
    int before[2] = {0, 0}, total[2] = {0, 0};
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < warp) { before[0] += s_cnt[0][w]; before[1] += s_cnt[1][w]; }
      total[0] += s_cnt[0][w]; total[1] += s_cnt[1][w];
    }
    const unsigned lt = (1u << lane) - 1u;
    if (is_s) sem_out[(size_t)b * n_tokens + run[0] + before[0] + __popc(ms & lt)] = (int)(id - sem_base);
    if (is_g) {
      const int pos = run[1] + before[1] + __popc(mg & lt);
      if (pos < max_global) glob_out[(size_t)b * max_global + pos] = (int)(id - glob_base);
    }
    run[0] += total[0]; run[1] += total[1];
    __syncthreads();
  }
  if (threadIdx.x == 0) { sem_len[b] = run[0]; glob_len[b] = run[1]; }
}

// ------------------------------------------------------------------- depthwise conv + LayerNorm
// x (batch, rows, C) fp32 -> [dwconv k=7 pad 3 over rows] -> LayerNorm over C (biased variance, eps)
// -> * scale[b] + shift[b] -> fp32 and/or operand planes.
// (vocos.py:69-75 ConvNeXtBlock head, :105-110 AdaLayerNorm, :328-334 backbone norms.)
//
// One persistent CTA per SM walks 32-row tiles of one utterance with a three-deep pipeline of BULK async copies:
// a tile with its halo is one contiguous span of memory, so one elected thread moves it with a single
// cp.async.bulk (TMA, mbarrier completion) while the other warps normalise tile i; tiles i+1 and i+2 are in
// flight.  Halo rows outside the utterance (the conv's zero padding) are zero-filled by the CTA.  A warp owns 4 CONSECUTIVE output rows: a lane holds C/32 channels as float4s,
// keeps its 7 x C/32 depthwise taps in REGISTERS for the whole kernel and reads each of the 10 input rows it
// needs once (7.5 LDS.128 per output row instead of 21 + 21 weight loads); the row statistics are
// warp-shuffle reductions, four rows interleaved.  Nothing but the tile itself is read in the main loop,
// so the tiny L1 left next to 117 KB of shared memory is not in the way.
constexpr int kLnRows = 32;        // output rows per tile
constexpr int kLnWarpRows = 4;     // consecutive rows per warp (8 warps x 4 = 32)
constexpr int kLnStages = 3;       // tiles in shared memory

template <int NV, bool DW>
__global__ void __launch_bounds__(256, 1)
dwconv_ln_kernel(const float* __restrict__ x, int rows, const float* __restrict__ dw_w /* [7][C] */,
                 const float* __restrict__ dw_b, const float* __restrict__ scale, const float* __restrict__ shift,
                 int ss_stride, float eps, float* __restrict__ out_f32, OpBuf out_op, int tiles_per_utt,
                 int total_tiles) {
  constexpr int C = NV * 128;
  constexpr int HALO = DW ? 3 : 0;
  constexpr int TAPS = DW ? 7 : 1;
  constexpr int TROWS = kLnRows + 2 * HALO;
  constexpr int C4 = C / 4;
  extern __shared__ __align__(16) float s_buf[];   // [kLnStages][TROWS][C]
  const uint32_t s_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_buf));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  __shared__ uint64_t s_bar[kLnStages];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kLnStages; ++i) mbar_init(smem_u32(&s_bar[i]), 1);
    fence_barrier_init();
  }
  __syncthreads();
  // rows [r0 - HALO, r0 + kLnRows + HALO) of utterance b -> buffer `buf`; only rows inside the utterance are copied
  auto issue = [&](int tile, int buf) {
    const int b = tile / tiles_per_utt;
    const int r0 = (tile % tiles_per_utt) * kLnRows;
    const int ra = max(r0 - HALO, 0), rb = min(r0 + kLnRows + HALO, rows);
    const int first = ra - (r0 - HALO);                   // staged row index of utterance row ra
    const uint32_t dst = s_base + (uint32_t)((buf * TROWS + first) * C * 4);
    if (first > 0 || rb - (r0 - HALO) < TROWS) {          // edge tile: zero the staged rows outside the utterance
      for (int i = threadIdx.x; i < TROWS * C4; i += blockDim.x) {
        const int rr = i / C4;
        if (rr < first || rr >= rb - (r0 - HALO))
          reinterpret_cast<float4*>(s_buf)[buf * TROWS * C4 + i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)(rb - ra) * C * 4;
      const uint32_t bar = smem_u32(&s_bar[buf]);
      mbar_expect_tx(bar, bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(x + ((size_t)b * rows + ra) * C), "r"(bytes), "r"(bar)
                   : "memory");
    }
  };

  int tile = blockIdx.x, buf = 0;
  uint32_t phase = 0;

  // depthwise taps / bias of this lane's channels: registers for the whole kernel (model constants: loaded before the
  // PDL wait, so the loads overlap the previous kernel's tail)
  float4 wt[TAPS][NV], bias[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = v * 128 + lane * 4;
    bias[v] = DW ? __ldg(reinterpret_cast<const float4*>(dw_b + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < TAPS; ++j)
      wt[j][v] = DW ? __ldg(reinterpret_cast<const float4*>(dw_w + j * C + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
  griddep_wait();                 // PDL (common.cuh): x, scale and shift are outputs of earlier kernels
  griddep_launch_dependents();
#pragma unroll
  for (int i = 0; i < kLnStages - 1; ++i)   // prologue: kLnStages - 1 tiles in flight
    if (tile + i * (int)gridDim.x < total_tiles) issue(tile + i * gridDim.x, i);

  for (; tile < total_tiles; tile += gridDim.x) {
    const int nxt = tile + (kLnStages - 1) * gridDim.x;
    const int b = tile / tiles_per_utt;
    const int r0 = (tile % tiles_per_utt) * kLnRows;
    // per-utterance affine of this lane's channels (AdaLayerNorm: scale/shift depend on the utterance)
    float4 g[NV], h[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = v * 128 + lane * 4;
      g[v] = __ldg(reinterpret_cast<const float4*>(scale + (size_t)b * ss_stride + c));
      h[v] = __ldg(reinterpret_cast<const float4*>(shift + (size_t)b * ss_stride + c));
    }
    // the buffer of tile - 1 (everyone left it at the barrier that closed the previous iteration) takes tile + 2
    if (nxt < total_tiles) issue(nxt, (buf + kLnStages - 1) % kLnStages);
    mbar_wait(smem_u32(&s_bar[buf]), phase);
    __syncthreads();   // (edge tiles: the zero-fill of this buffer by other threads is complete and visible)
    const float* s_x = s_buf + buf * TROWS * C;
    const int rw = warp * kLnWarpRows;                 // first output row of this warp inside the tile
    if (r0 + rw < rows) {
      float4 y[kLnWarpRows][NV];
#pragma unroll
      for (int q = 0; q < kLnWarpRows; ++q)
#pragma unroll
        for (int v = 0; v < NV; ++v) y[q][v] = bias[v];
      // input row (rw + j) of the staged tile feeds output row q through tap j - q
#pragma unroll
      for (int j = 0; j < kLnWarpRows + TAPS - 1; ++j) {
        float4 xin[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) xin[v] = *reinterpret_cast<const float4*>(s_x + (rw + j) * C + v * 128 + lane * 4);
#pragma unroll
        for (int q = 0; q < kLnWarpRows; ++q) {
          const int t = j - q;
          if (t >= 0 && t < TAPS) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              if (DW) {
                y[q][v].x = fmaf(wt[t][v].x, xin[v].x, y[q][v].x); y[q][v].y = fmaf(wt[t][v].y, xin[v].y, y[q][v].y);
                y[q][v].z = fmaf(wt[t][v].z, xin[v].z, y[q][v].z); y[q][v].w = fmaf(wt[t][v].w, xin[v].w, y[q][v].w);
              } else {
                y[q][v] = xin[v];
              }
            }
          }
        }
      }
      float sum[kLnWarpRows], sq[kLnWarpRows];
#pragma unroll
      for (int q = 0; q < kLnWarpRows; ++q) {
        sum[q] = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) sum[q] += (y[q][v].x + y[q][v].y) + (y[q][v].z + y[q][v].w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < kLnWarpRows; ++q) sum[q] += __shfl_xor_sync(0xffffffffu, sum[q], o);
#pragma unroll
      for (int q = 0; q < kLnWarpRows; ++q) {
        const float mean = sum[q] * (1.0f / C);
        sq[q] = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          y[q][v].x -= mean; y[q][v].y -= mean; y[q][v].z -= mean; y[q][v].w -= mean;
          sq[q] += (y[q][v].x * y[q][v].x + y[q][v].y * y[q][v].y) + (y[q][v].z * y[q][v].z + y[q][v].w * y[q][v].w);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < kLnWarpRows; ++q) sq[q] += __shfl_xor_sync(0xffffffffu, sq[q], o);
#pragma unroll
      for (int q = 0; q < kLnWarpRows; ++q) {
        const int r = r0 + rw + q;
        if (r < rows) {
          const float rstd = rsqrtf(sq[q] * (1.0f / C) + eps);
          const size_t row_off = ((size_t)b * rows + r) * C;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const int c = v * 128 + lane * 4;
            float4 o;
            o.x = fmaf(y[q][v].x * rstd, g[v].x, h[v].x); o.y = fmaf(y[q][v].y * rstd, g[v].y, h[v].y);
            o.z = fmaf(y[q][v].z * rstd, g[v].z, h[v].z); o.w = fmaf(y[q][v].w * rstd, g[v].w, h[v].w);
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + row_off + c) = o;
            if (out_op.hi) store_op4(out_op, row_off + c, o);
          }
        }
      }
    }
    // everyone is done with this buffer before the next-but-one tile streams into it (the generic-proxy
    // accesses are ordered before the async-proxy bulk copy by the fence + barrier)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (++buf == kLnStages) { buf = 0; phase ^= 1u; }
  }
}

// ------------------------------------------------------------------------------------ waveform head
// wav[b, l] = tanh(bias + sum_j sum_c w[j, c] * snake(x[b, l + j - 3, c]))     (wave_generator.py:77-81)
// A warp walks a run of kHeadRun consecutive samples of one utterance.  The rows of a run are contiguous in
// memory, so the warp streams them into its own shared-memory ring with coalesced 16 B cp.async copies, 3 granules
// (12 rows, 4.5 KB) ahead of the row it is working on; 20 warps per SM keep ~90 KB in flight (the first version kept
// 8 rows per warp in registers: 36 KB per SM, latency bound).
// Rows outside the utterance are zero-filled by cp.async; snake(0) == 0, so that IS the conv's zero padding.
// Every lane owns C/32 channels (conflict-free LDS) and Snakes each element once.
//
// The kernel is bound by instruction issue (ncu: 75 % issue-active at 59 % of the DRAM peak), so it works on TWO ROWS
// per step with packed fp32x2 arithmetic (FMUL2 / FFMA2, op_split.cuh): a register pair holds the Snake'd values of one
// channel in rows 2t and 2t + 1, and tap j adds w[j] * (s[2t], s[2t + 1]) -- the weight is a broadcast scalar operand
// -- into the running sums of the ADJACENT outputs (2t - j, 2t + 1 - j).  Even taps hit pairs that start at an even
// output, odd taps pairs that start at an odd one, so there are two sets of four pair accumulators:
//   E[k] = outputs (2k, 2k + 1)   <- taps 0, 2, 4, 6 of step t go to E[t], E[t-1], E[t-2], E[t-3]
//   O[k] = outputs (2k + 1, 2k + 2) <- taps 1, 3, 5    of step t go to O[t-1], O[t-2], O[t-3]
// After step t: out[2t - 6] = E[t-3].x + O[t-4].y and out[2t - 5] = E[t-3].y + O[t-3].x are complete.  The loop body
// is unrolled over four steps (two granules), so every accumulator index is a compile-time constant (no register
// moves): 21 FFMA2 per two rows instead of 42 FFMA; ncu counts 49 warp instructions per row all in (prologue, copies,
// transposes), 85 for the round-1 kernel, and the kernel now runs at 95 % of the copy peak under ncu.
// The per-lane partial sums of 32 consecutive samples are transposed through a padded per-warp smem tile (one
// st.shared + one ld.shared per sample instead of a 5-step shuffle tree per sample); lane L ends up with sample L, so
// the stores are coalesced 128 B lines.
// Run length and ring depth (same-box A/B, 64 x 160000 rows, profiles/r2_head_ab.txt): a ring of 4 granules leaves room
// for 5 blocks = 20 warps per SM (8 granules: 12 warps) -- with half the instructions per row the kernel needs warps
// to hide latency more than bytes in flight (90 KB per SM are enough) -- and a run of 256 samples halves the
// per-warp prologue (27 parameter loads) and the 6-row halo per run: 1.04 ms (128 / 8) -> 0.86 ms (256 / 4).
#ifndef SC_HEAD_RUN
#define SC_HEAD_RUN 256
#endif
#ifndef SC_HEAD_RING
#define SC_HEAD_RING 4
#endif
constexpr int kHeadRun = SC_HEAD_RUN;
constexpr int kHeadWarps = 4;
constexpr int kHeadGranRows = 4;     // rows per cp.async group (4 rows x 96 ch = 3 x 32 float4 for C = 96)
constexpr int kHeadRingGran = SC_HEAD_RING;     // granules in the ring (all but one in flight + the one being read)

// (s0, s1) = snake of the pair x with per-channel alpha / 1 / alpha.  EXACT: range-reduced sine (gemm_params.cuh
// snake_f, same operations per element); otherwise the SFU sine on the raw argument (snake_fast).
template <bool EXACT>
__device__ __forceinline__ f32x2 snake_pair(f32x2 x, float a, float inv) {
  f32x2 t = mul2(pk2(a), x);
  if (EXACT) {
    const f32x2 k = add2(fma2(t, pk2(0.15915494309189535f), pk2(12582912.0f)), pk2(-12582912.0f));
    t = fma2(k, pk2(-6.2831854820251465f), t);
    t = fma2(k, pk2(1.7484555e-7f), t);
  }
  float t0, t1;
  upk2(t, t0, t1);
  const f32x2 sn = pk2(__sinf(t0), __sinf(t1));
  return fma2(mul2(pk2(inv), sn), sn, x);
}

template <int NCH, bool EXACT>
__global__ void __launch_bounds__(kHeadWarps * 32)
head_kernel(const float* __restrict__ x, int rows, const float* __restrict__ alpha,
            const float* __restrict__ inv_alpha, const float* __restrict__ w /* [7][C] */, float bias,
            float* __restrict__ wav, int runs_per_utt, int total_runs) {
  constexpr int C = NCH * 32;
  constexpr int GRAN_F4 = kHeadGranRows * C / 4;          // float4s per granule (multiple of 32 for C % 32 == 0)
  constexpr int RING_ROWS = kHeadGranRows * kHeadRingGran;
  constexpr int N_GRAN = (kHeadRun + 6 + kHeadGranRows - 1) / kHeadGranRows;   // granules of a run incl. halo
  static_assert(kHeadGranRows == 4, "a granule is two steps of two rows");
  extern __shared__ __align__(16) float s_head[];         // per warp: ring [RING_ROWS][C] | part [32][33]
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = s_head + (size_t)wib * (RING_ROWS * C + 32 * 33);
  float(*sp)[33] = reinterpret_cast<float(*)[33]>(ring + RING_ROWS * C);
  const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(ring));
  const int run = blockIdx.x * kHeadWarps + wib;
  if (run >= total_runs) return;
  const int b = run / runs_per_utt;
  const int l0 = (run % runs_per_utt) * kHeadRun;
  const float* xb = x + (size_t)b * rows * C;

  float a[NCH], ia[NCH], wt[7][NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    a[k] = __ldg(alpha + lane + 32 * k);
    ia[k] = __ldg(inv_alpha + lane + 32 * k);
#pragma unroll
    for (int j = 0; j < 7; ++j) wt[j][k] = __ldg(w + j * C + lane + 32 * k);
  }
  griddep_wait();                 // PDL (common.cuh): x is the previous kernel's output
  griddep_launch_dependents();
  // granule g holds rows l0 - 3 + 4g .. +3 of the utterance.  Per lane and copy: row inside the granule and element
  // offset inside the row are loop invariants; the address is 32-bit element arithmetic + one widening multiply-add
  // (an utterance has < 2^31 elements: checked by the launcher).
  int g_row[GRAN_F4 / 32], g_col[GRAN_F4 / 32];
#pragma unroll
  for (int i = 0; i < GRAN_F4 / 32; ++i) {
    const int f = i * 32 + lane;                           // float4 index inside the granule
    g_row[i] = (f * 4) / C;
    g_col[i] = (f * 4) % C;
  }
  auto issue = [&](int g) {
    if (g < N_GRAN) {
      const uint32_t dst = ring_s + (uint32_t)(((g % kHeadRingGran) * GRAN_F4 + lane) * 16);
      const int r0 = l0 - 3 + g * kHeadGranRows;
#pragma unroll
      for (int i = 0; i < GRAN_F4 / 32; ++i) {
        const int r = r0 + g_row[i];
        const bool ok = (unsigned)r < (unsigned)rows;
        const int off = (ok ? r * C : 0) + g_col[i];       // (a zero-size copy still gets a valid address)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)(i * 32 * 16)), "l"(xb + off),
                     "r"(ok ? 16 : 0)
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // (empty groups keep the wait counts uniform)
  };
#pragma unroll
  for (int g = 0; g < kHeadRingGran - 1; ++g) issue(g);

  // Row ri of the run (row 0 = sample l0 - 3) feeds output sample ri - j through tap j.  Step t = rows 2t, 2t + 1.
  f32x2 E[4], O[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { E[i] = 0ull; O[i] = 0ull; }
  const int l_end = min(l0 + kHeadRun, rows);
  constexpr int N_GRAN2 = (N_GRAN + 1) / 2 * 2;
#pragma unroll 1
  for (int g0 = 0; g0 < N_GRAN2; g0 += 2) {
#pragma unroll
    for (int gg = 0; gg < 2; ++gg) {
      const int g = g0 + gg;
      if (g < N_GRAN) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kHeadRingGran - 2) : "memory");   // granule g has landed
        __syncwarp();
        const float* gp = ring + (g % kHeadRingGran) * (kHeadGranRows * C);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          constexpr int kP = 4;
          const int tt = gg * 2 + u;              // compile-time step index inside the 4-step block (= t mod 4)
          const int t = g * 2 + u;                // step of the run
          f32x2 sv[NCH];
#pragma unroll
          for (int k = 0; k < NCH; ++k)
            sv[k] = snake_pair<EXACT>(pk2(gp[(2 * u) * C + lane + 32 * k], gp[(2 * u + 1) * C + lane + 32 * k]), a[k], ia[k]);
#pragma unroll
          for (int m = 0; m < 4; ++m) {           // even taps j = 2m -> E[t - m]
            const int e = ((tt - m) % kP + kP) % kP;
#pragma unroll
            for (int k = 0; k < NCH; ++k) E[e] = fma2(pk2(wt[2 * m][k]), sv[k], E[e]);
          }
#pragma unroll
          for (int m = 0; m < 3; ++m) {           // odd taps j = 2m + 1 -> O[t - m - 1]
            const int o = ((tt - m - 1) % kP + kP) % kP;
#pragma unroll
            for (int k = 0; k < NCH; ++k) O[o] = fma2(pk2(wt[2 * m + 1][k]), sv[k], O[o]);
          }
          // outputs 2t - 6 = E[t-3].x + O[t-4].y and 2t - 5 = E[t-3].y + O[t-3].x are complete
          const int e3 = ((tt - 3) % kP + kP) % kP, o3 = e3, o4 = ((tt - 4) % kP + kP) % kP;
          float ex, ey, o3x, o3y, o4x, o4y;
          upk2(E[e3], ex, ey);
          upk2(O[o3], o3x, o3y);
          upk2(O[o4], o4x, o4y);
          (void)o3y; (void)o4x;
          const int so = 2 * t - 6;               // first of the two finished samples (even)
          if (so >= 0 && so < kHeadRun) {
            sp[so & 31][lane] = ex + o4y;
            sp[(so + 1) & 31][lane] = ey + o3x;
            if (((so + 1) & 31) == 31) {   // 32 samples done: transpose-reduce through shared memory, lane L gets sample L
              __syncwarp();
              float tot = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) tot += sp[lane][j];
              __syncwarp();
              const int l = l0 + so + 1 - 31 + lane;
              if (l < l_end) wav[(size_t)b * rows + l] = tanhf(tot + bias);
            }
          }
          E[e3] = 0ull;    // becomes E[t + 1]
          O[o4] = 0ull;    // becomes O[t]  (first written by tap 1 of step t + 1)
        }
        __syncwarp();            // all lanes are done reading this granule's slot before it is refilled
        issue(g + kHeadRingGran - 1);
      }
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------ launchers
int launch_split(const float* x, OpBuf out, size_t n, cudaStream_t s) {
  if (n % 4) { set_error("split: element count must be a multiple of 4"); return SPARKCODEC_EINVAL; }
  const size_t n4 = n / 4;
  const int grid = (int)std::min<size_t>((n4 + 255) / 256, 148 * 16);
  split_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(x), out, n4);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_split_rows(const float* x, size_t src_batch_stride, OpBuf out, int batch, int rows, int c, int rows_total,
                      int row_off, cudaStream_t s) {
  if (c % 4 || src_batch_stride % 4) { set_error("split_rows: channel count / stride must be multiples of 4"); return SPARKCODEC_EINVAL; }
  const size_t n4 = (size_t)batch * rows * (c / 4);
  if (n4 == 0) return 0;
  const int grid = (int)std::min<size_t>((n4 + 255) / 256, 148 * 16);
  split_rows_kernel<<<grid, 256, 0, s>>>(x, src_batch_stride, out, rows, c / 4, rows_total, row_off, n4);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_merge(const OpBuf& in, float* out, size_t n, cudaStream_t s) {
  const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
  merge_kernel<<<grid, 256, 0, s>>>(in, out, n);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_extract_codes(const void* ids, int id_dtype, int batch, int n_tokens, long long sem_base, int sem_size,
                         long long glob_base, int glob_size, int* sem_out, int* sem_len, int* glob_out, int max_global,
                         int* glob_len, cudaStream_t s) {
  extract_codes_kernel<<<batch, 256, 0, s>>>(ids, id_dtype, n_tokens, sem_base, sem_size, glob_base, glob_size, sem_out,
                                             sem_len, glob_out, max_global, glob_len);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_vq_embed(const void* sem, int sem_dtype, int batch, int frames, int, int, int codebook_size,
                    int codebook_dim, const float* codebook, const float* mat, const float* vec, int c_out,
                    OpBuf out, int* err_flag, cudaStream_t s) {
  const size_t n_tok = (size_t)batch * frames, n = n_tok * (c_out / 4);
  vq_embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sem, sem_dtype, n_tok, codebook_size, codebook_dim,
                                                             codebook, mat, vec, c_out, out, err_flag);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_vq_zq(const void* sem, int sem_dtype, int n_tok, int codebook_size, int codebook_dim,
                 const float* codebook, const float* w, const float* bias, int d_model, float* out, cudaStream_t s) {
  const size_t n = (size_t)n_tok * d_model;
  vq_zq_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sem, sem_dtype, (size_t)n_tok, codebook_size, codebook_dim,
                                                          codebook, w, bias, d_model, out);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_vq_rows(const void* sem, int sem_dtype, int n_tok, int codebook_size, int codebook_dim,
                   const float* codebook, float* out, cudaStream_t s) {
  const size_t n = (size_t)n_tok * codebook_dim;
  vq_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sem, sem_dtype, (size_t)n_tok, codebook_size, codebook_dim,
                                                            codebook, out);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_fsq_codes(const void* glob, int glob_dtype, int n_tok, int n_levels, const int* levels, float* out,
                     cudaStream_t s) {
  fsq_codes_kernel<<<(n_tok + 127) / 128, 128, 0, s>>>(glob, glob_dtype, n_tok, n_levels, levels, out);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_vq_search(const float* x, size_t n_frames, int c, const float* mat, const float* vec, const float* codes_n,
                     const float* codes_sq, int codebook_size, int codebook_dim, long long* idx_out, float* margin_out,
                     cudaStream_t s) {
  if (codebook_dim != kVqDim) { set_error("vq_search: codebook_dim must be %d", kVqDim); return SPARKCODEC_EINVAL; }
  const size_t warps = (n_frames + kVqFrames - 1) / kVqFrames;
  vq_search_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(x, n_frames, c, mat, vec, codes_n, codes_sq,
                                                              codebook_size, idx_out, margin_out);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_fsq_project(const void* glob, int glob_dtype, int batch, int token_num, int n_levels, const int* levels,
                       const float* w_po, const float* b_po, int latent, float* flat_out, int* err_flag,
                       cudaStream_t s) {
  fsq_project_kernel<<<batch * token_num, 128, 0, s>>>(glob, glob_dtype, token_num, n_levels, levels, w_po, b_po,
                                                      latent, flat_out, err_flag);
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_small_linear(const float* x, const float* w, const float* bias, float* y, int batch, int k, int n,
                        cudaStream_t s) {
  if (k % 4) { set_error("small_linear: K must be a multiple of 4"); return SPARKCODEC_EINVAL; }
  small_linear_kernel<<<(n + 7) / 8, 256, 0, s>>>(x, w, bias, y, batch, k, n);
  SC_LAUNCH_CHECK();
  return 0;
}

template <int NV>
static int launch_dwconv_ln_t(const float* x, int batch, int rows, const float* dw_w, const float* dw_b,
                              const float* scale, const float* shift, int ss, float eps, float* out_f32,
                              OpBuf out_op, cudaStream_t s) {
  const int tiles = (rows + kLnRows - 1) / kLnRows, total = batch * tiles;
  const bool dw = dw_w != nullptr;
  const size_t smem = kLnStages * (size_t)(kLnRows + (dw ? 6 : 0)) * NV * 128 * sizeof(float);
  if (smem > 226 * 1024) { set_error("dwconv_ln: %d channels do not fit the shared-memory pipeline", NV * 128); return SPARKCODEC_EINVAL; }
  static PerDevice sms, done_dw, done_ln;
  int num_sms = sms.here().load(std::memory_order_relaxed);
  if (!num_sms) {
    int dev = 0;
    SC_CUDA(cudaGetDevice(&dev));
    SC_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    sms.here().store(num_sms, std::memory_order_relaxed);
  }
  const int grid = std::min(total, num_sms);   // one persistent CTA per SM
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = add_pdl_attr(attr, 0);
  if (dw) {
    if (!done_dw.here().load(std::memory_order_relaxed)) {
      SC_CUDA(cudaFuncSetAttribute(dwconv_ln_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      done_dw.here().store(1, std::memory_order_relaxed);
    }
    SC_CUDA(cudaLaunchKernelEx(&cfg, dwconv_ln_kernel<NV, true>, x, rows, dw_w, dw_b, scale, shift, ss, eps, out_f32,
                               out_op, tiles, total));
  } else {
    if (!done_ln.here().load(std::memory_order_relaxed)) {
      SC_CUDA(cudaFuncSetAttribute(dwconv_ln_kernel<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      done_ln.here().store(1, std::memory_order_relaxed);
    }
    SC_CUDA(cudaLaunchKernelEx(&cfg, dwconv_ln_kernel<NV, false>, x, rows, dw_w, dw_b, scale, shift, ss, eps, out_f32,
                               out_op, tiles, total));
  }
  SC_LAUNCH_CHECK();
  return 0;
}
int launch_dwconv_ln(const float* x, int batch, int rows, int c, const float* dw_w, const float* dw_b,
                     const float* scale, const float* shift, int ss_batch_stride, float eps, float* out_f32,
                     OpBuf out_op, cudaStream_t s) {
  switch (c) {
    case 128: return launch_dwconv_ln_t<1>(x, batch, rows, dw_w, dw_b, scale, shift, ss_batch_stride, eps, out_f32, out_op, s);
    case 256: return launch_dwconv_ln_t<2>(x, batch, rows, dw_w, dw_b, scale, shift, ss_batch_stride, eps, out_f32, out_op, s);
    case 384: return launch_dwconv_ln_t<3>(x, batch, rows, dw_w, dw_b, scale, shift, ss_batch_stride, eps, out_f32, out_op, s);
    case 512: return launch_dwconv_ln_t<4>(x, batch, rows, dw_w, dw_b, scale, shift, ss_batch_stride, eps, out_f32, out_op, s);
    default: set_error("dwconv_ln: unsupported channel count %d (128/256/384/512)", c); return SPARKCODEC_EINVAL;
  }
}

template <int NCH, bool EXACT>
static int launch_head_t(const float* x, int batch, int rows, const float* alpha, const float* inv_alpha, const float* w,
                         float bias, float* wav, cudaStream_t s) {
  const int runs = (rows + kHeadRun - 1) / kHeadRun, total = batch * runs;
  const int grid = (total + kHeadWarps - 1) / kHeadWarps, blk = kHeadWarps * 32;
  if ((long long)rows * NCH * 32 >= (1ll << 31)) { set_error("head: %d rows x %d channels overflow the 32-bit row offsets", rows, NCH * 32); return SPARKCODEC_EINVAL; }
  const size_t smem = (size_t)kHeadWarps * (kHeadGranRows * kHeadRingGran * NCH * 32 + 32 * 33) * sizeof(float);
  static PerDevice done;
  if (!done.here().load(std::memory_order_relaxed)) {
    SC_CUDA(cudaFuncSetAttribute(head_kernel<NCH, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done.here().store(1, std::memory_order_relaxed);
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(blk);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = add_pdl_attr(attr, 0);
  SC_CUDA(cudaLaunchKernelEx(&cfg, head_kernel<NCH, EXACT>, x, rows, alpha, inv_alpha, w, bias, wav, runs, total));
  SC_LAUNCH_CHECK();
  return 0;
}

int launch_head(const float* x, int batch, int rows, int c, const float* alpha, const float* inv_alpha,
                const float* w, float bias, float* wav, int exact_sin, int, cudaStream_t s) {
#define SC_HEAD(NCH)                                                                                             \
  return exact_sin ? launch_head_t<NCH, true>(x, batch, rows, alpha, inv_alpha, w, bias, wav, s)                 \
                   : launch_head_t<NCH, false>(x, batch, rows, alpha, inv_alpha, w, bias, wav, s);
  switch (c) {
    case 32: SC_HEAD(1)
    case 64: SC_HEAD(2)
    case 96: SC_HEAD(3)
    case 128: SC_HEAD(4)
    default: set_error("head: unsupported channel count %d (32/64/96/128)", c); return SPARKCODEC_EINVAL;
  }
#undef SC_HEAD
}

}  // namespace sparkcodec
