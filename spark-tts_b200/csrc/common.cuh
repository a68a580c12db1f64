// Shared declarations for libsparkcodec (B200 / sm_100a BiCodec detokenize path).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <atomic>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/sparkcodec.h"

namespace sparkcodec {

// ---- error plumbing --------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local int64_t* g_launch_counter;   // points at the active handle's counter (or null)

#define SC_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::sparkcodec::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,           \
                              cudaGetErrorString(_e));                                     \
      return SPARKCODEC_ECUDA;                                                             \
    }                                                                                      \
  } while (0)

#define SC_TRY(expr)                                                                       \
  do {                                                                                     \
    int _r = (expr);                                                                       \
    if (_r != 0) return _r;                                                                \
  } while (0)

#define SC_LAUNCH_CHECK()                                                                  \
  do {                                                                                     \
    if (::sparkcodec::g_launch_counter) ++*::sparkcodec::g_launch_counter;                 \
    SC_CUDA(cudaGetLastError());                                                           \
  } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------
// The four kernels that make up a pass (conv_gemm_tc, resunit_fused, dwconv_ln, head) are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's CTAs may become resident as soon as every CTA of
// the previous one has passed its own griddep_wait() (SMs it does not fill, or CTAs that have exited), so barrier
// initialisation, tensor-memory allocation, tensor-map prefetches, parameter loads and the launch latency itself
// overlap the previous kernel's tail.  EVERY thread executes griddep_wait() before the kernel touches anything another
// kernel may have written or may still be reading (it returns once all prerequisite grids have completed and their
// memory is visible; a no-op for a kernel launched without the attribute), so completion stays transitive along the
// chain and buffer reuse (ping-pong planes, the in-place residual stream) is as safe as with plain stream order.
// SPARKCODEC_PDL=0 launches everything with plain stream order.
#ifdef __CUDACC__
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
// appends the PDL attribute to attr[n] (room for it is the caller's business); returns the new attribute count
inline int add_pdl_attr(cudaLaunchAttribute* attr, int n) {
  if (!pdl_enabled()) return n;
  attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[n].val.programmaticStreamSerializationAllowed = 1;
  return n + 1;
}

// ---- activations -------------------------------------------------------------------------------
// All activations are channels-last (batch, rows, channels).  A dense-contraction operand is stored as two
// 2-byte-per-element planes (the "OP" format), in one of two formats:
//   OPFMT_BF16  : hi = bf16(x), lo = bf16(x - hi).  bf16 mode multiplies hi*Whi only (lo == null); the three-term
//                 fp32 mode hi*Whi + lo*Whi + hi*Wlo (kind::f16 MMAs, ~2^-16 relative per product).
//   OPFMT_F16F8 : hi = fp16(x) (11 significant bits); the second plane holds, per group of 32 channels, 64 bytes
//                 [ e5m2((x - hi) * 2^kLoShift) x 32 | e5m2(hi * 2^-kHiShift) x 32 ].  The two-term fp32 mode multiplies
//                 hi*Whi with kind::f16 MMAs and adds BOTH cross terms lo*Whi + hi*Wlo with kind::f8f6f4 MMAs (K = 32,
//                 twice the rate) against a weight plane laid out [ e5m2(Whi * 2^-kLoShift) | e5m2(Wlo * 2^kHiShift) ]:
//                 the cross terms are 2^-12 of the product, so their 3-bit significands cost ~2^-15 relative -- 74 dB
//                 waveform SNR in the oracle simulation (tools/sim_split_precision.py) against a 60 dB bar, for two
//                 thirds of the three-term tensor time.  Plane sizes, TMA boxes and K stepping are identical in both formats.
enum OpFmt { OPFMT_BF16 = 0, OPFMT_F16F8 = 1 };
constexpr int kLoShift = 4;   // activations: lo * 2^4 ; weights: hi * 2^-4   (e5m2 normal range 2^-14 .. 2^15)
constexpr int kHiShift = 8;   // activations: hi * 2^-8; weights: lo * 2^8
struct OpBuf {
  __nv_bfloat16* hi = nullptr;   // bf16 or fp16 bits
  __nv_bfloat16* lo = nullptr;   // null in bf16 mode
  int fmt = OPFMT_BF16;
};
// MMA terms of the fp32 precision mode: 2 (default, OPFMT_F16F8) or 3 (OPFMT_BF16); env SPARKCODEC_FP32_TERMS
int fp32_terms();
// tensor-core products per MAC of a precision mode: BF16 1, FP32 fp32_terms(), FP32X3 3
inline int terms_for(int precision) {
  return precision == SPARKCODEC_PREC_BF16 ? 1 : (precision == SPARKCODEC_PREC_FP32X3 ? 3 : fp32_terms());
}
inline bool is_split(int precision) { return precision != SPARKCODEC_PREC_BF16; }   // two operand planes
inline int op_fmt_for(int precision) { return terms_for(precision) == 2 ? OPFMT_F16F8 : OPFMT_BF16; }

constexpr int kMaxPhases = 8;   // polyphase branches of a transposed conv (= stride)
constexpr int kMaxTaps = 7;     // taps per branch (k=7 convs; <=3 for the transposed convs)
constexpr int kBlockM = 128;    // rows of one accumulator tile (UMMA M)

// Tap table of one dense contraction seen as D[b,l,n] = sum_j sum_c A[b, l+shift[ph][j], c] * W[n, j*C_in + c]
struct TapTable {
  int n_phase = 1;
  int cols_per_phase = 0;                  // N_total / n_phase (= C_out)
  int ntaps[kMaxPhases] = {0};
  int shift[kMaxPhases][kMaxTaps] = {{0}};
};

// Device-resident packed weights of one dense contraction.
struct GemmWeights {
  int c_in = 0, n_total = 0, kt = 0;       // W is (n_total, kt*c_in) K-major
  TapTable taps;
  __nv_bfloat16* w_hi = nullptr;           // OPFMT_BF16 planes
  __nv_bfloat16* w_lo = nullptr;
  __nv_bfloat16* w_h16 = nullptr;          // OPFMT_F16F8 planes: fp16(W) and the packed e5m2 plane [Whi | Wlo]
  __nv_bfloat16* w_p8 = nullptr;
  float* bias = nullptr;                   // (n_total) (phase-replicated for transposed convs)
  int block_n = 0;                         // N tile of the tcgen05 kernel (divides cols_per_phase)
  CUtensorMap tmap_hi[2], tmap_lo[2];      // (K, N) boxes (bk, block_n) for bk = 64 ([0]) and 32 ([1])
  CUtensorMap tmap_h16[2], tmap_p8[2];
  // planes a kernel with NTERMS terms reads
  const __nv_bfloat16* hi_for(int nterms) const { return nterms == 2 ? w_h16 : w_hi; }
  const __nv_bfloat16* lo_for(int nterms) const { return nterms == 2 ? w_p8 : w_lo; }
  bool has_bk64 = false;                   // c_in % 64 == 0
};

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_SNAKE = 2 };

struct Epilogue {
  const float* rowbias = nullptr;    // (batch, n_total) added per utterance (d_vector), optional
  const float* residual = nullptr;   // (batch, L, n_total) fp32, optional (may alias out_f32)
  int act = ACT_NONE;
  const float* alpha = nullptr;      // (n_total) snake alpha of the NEXT layer's input snake
  const float* inv_alpha = nullptr;  // (n_total) 1/(alpha+1e-9)
  float* out_f32 = nullptr;          // pre-activation value (residual stream), optional
  OpBuf out_op;                      // post-activation operand planes, optional
};

// ---- kernels / launchers (defined in the .cu files) --------------------------------------------
int tma_init();   // resolves cuTensorMapEncodeTiled through the runtime (no link-time libcuda dep)
int make_weight_tmaps(GemmWeights& w);

// D = conv(A) with fused epilogue.  A: (batch, L, c_in) operand planes.
int launch_conv_gemm_tc(const GemmWeights& w, const OpBuf& a, int batch, int L, const Epilogue& ep,
                        int precision, int num_sms, cudaStream_t stream);
int launch_conv_gemm_simt(const GemmWeights& w, const OpBuf& a, int batch, int L, const Epilogue& ep,
                          int precision, cudaStream_t stream);

// Fused ResidualUnit (k7 dilated conv -> Snake -> 1x1 conv -> + x -> Snake) for C in {96, 192}; x is updated
// in place, `out` (optional) receives snake_out(x) as operand planes.  ru_fused.cu
bool resunit_fusable(const GemmWeights& c7, const GemmWeights& c1, int* dil);
int launch_resunit_fused(const GemmWeights& c7, const GemmWeights& c1, const OpBuf& a, int batch, int L,
                         const float* alpha_mid, const float* inv_mid, float* x, const float* alpha_out,
                         const float* inv_out, OpBuf out, int precision, int num_sms, cudaStream_t stream);

// streaming / token kernels
// One-time per-DEVICE launch state (function attributes are per context, so a process that drives several GPUs must
// set them on each; relaxed atomics because two host threads may race to the same value).
constexpr int kMaxDevices = 64;
struct PerDevice {
  std::atomic<int> v[kMaxDevices];
  PerDevice() { for (auto& x : v) x.store(0, std::memory_order_relaxed); }
  std::atomic<int>& here() {
    int d = 0;
    cudaGetDevice(&d);
    return v[(d >= 0 && d < kMaxDevices) ? d : 0];
  }
};

int launch_split(const float* x, OpBuf out, size_t n, cudaStream_t s);                 // fp32 -> hi/lo
int launch_merge(const OpBuf& in, float* out, size_t n, cudaStream_t s);               // hi(+lo) -> fp32
// rows (batch, rows, c) fp32 with batch stride `src_batch_stride` floats -> planes of a (batch, rows_total, c)
// tensor at row offset `row_off`
int launch_split_rows(const float* x, size_t src_batch_stride, OpBuf out, int batch, int rows, int c, int rows_total,
                      int row_off, cudaStream_t s);
int launch_extract_codes(const void* ids, int id_dtype, int batch, int n_tokens, long long sem_base, int sem_size,
                         long long glob_base, int glob_size, int* sem_out, int* sem_len, int* glob_out, int max_global,
                         int* glob_len, cudaStream_t s);
int launch_vq_embed(const void* sem, int sem_dtype, int batch, int frames, int t0, int rows,
                    int codebook_size, int codebook_dim, const float* codebook, const float* mat,
                    const float* vec, int c_out, OpBuf out, int* err_flag, cudaStream_t s);
int launch_vq_search(const float* x, size_t n_frames, int c, const float* mat, const float* vec, const float* codes_n,
                     const float* codes_sq, int codebook_size, int codebook_dim, long long* idx_out, float* margin_out,
                     cudaStream_t s);                                                       // encode side: nearest code
int launch_vq_zq(const void* sem, int sem_dtype, int n_tok, int codebook_size, int codebook_dim,
                 const float* codebook, const float* w, const float* bias, int d_model, float* out,
                 cudaStream_t s);
int launch_vq_rows(const void* sem, int sem_dtype, int n_tok, int codebook_size, int codebook_dim,
                   const float* codebook, float* out, cudaStream_t s);                      // tap: codebook[idx]
int launch_fsq_codes(const void* glob, int glob_dtype, int n_tok, int n_levels, const int* levels, float* out,
                     cudaStream_t s);                                                        // tap: FSQ level codes
int launch_fsq_project(const void* glob, int glob_dtype, int batch, int token_num, int n_levels,
                       const int* levels, const float* w_po, const float* b_po, int latent,
                       float* flat_out, int* err_flag, cudaStream_t s);
int launch_small_linear(const float* x, const float* w, const float* bias, float* y, int batch, int k,
                        int n, cudaStream_t s);
// (depthwise k=7 conv +) LayerNorm over channels, affine per utterance (scale/shift with batch stride)
int launch_dwconv_ln(const float* x, int batch, int rows, int c, const float* dw_w, const float* dw_b,
                     const float* scale, const float* shift, int ss_batch_stride, float eps,
                     float* out_f32, OpBuf out_op, cudaStream_t s);
int launch_head(const float* x, int batch, int rows, int c, const float* alpha, const float* inv_alpha,
                const float* w, float bias, float* wav, int exact_sin, int crop_rows, cudaStream_t s);

// ---- speaker half of tokenize (speaker_kernels.cu): plain fp32 FFMA kernels, channels-last activations ----
struct SpkGemm {
  const float* x = nullptr; int ldx = 0;       // input rows (batch * rows, K) with row stride ldx
  const float* x2 = nullptr; int ldx2 = 0;     // optional second input added element-wise on load
  int batch = 1, rows = 0, K = 0;
  int ntaps = 1; int shift[8] = {0};           // conv taps: row shifts (rows outside [0, rows) read as zero)
  const float* w = nullptr; int ldw = 0;       // W[n][j * K + c]
  const float* bias = nullptr;
  int relu = 0;                                // ReLU after the bias ...
  const float* scale = nullptr; const float* shift_v = nullptr;   // ... then the folded BatchNorm affine (optional)
  int act = 0;                                 // 1: sigmoid
  const float* res = nullptr; int ldres = 0;   // optional residual added last
  float* y = nullptr; int ldy = 0; int N = 0;
};
int launch_spk_frames(const float* wav, int batch, int n, int frames, int hop, int win, float* out, cudaStream_t s);
int launch_spk_magnitude(const float* spec, size_t rows, int bins, float* out, cudaStream_t s);
int launch_spk_gemm(const SpkGemm& g, cudaStream_t s);
int launch_spk_mean_rows(const float* x, int batch, int rows, int C, int ldx, float* out, cudaStream_t s);
int launch_spk_se_apply(const float* x, int ldx, const float* u, int ldu, const float* sc, int batch, int rows, int C,
                        float* out, int ldo, cudaStream_t s);
int launch_spk_copy_cols(const float* src, int lds, float* dst, int ldd, size_t rows, int cols, cudaStream_t s);
int launch_spk_place_rows(const float* src, size_t src_batch_stride, int batch, int src_rows, int C, float* dst,
                          int dst_rows, int row_off, cudaStream_t s);
int launch_spk_attention(const float* q, const float* kv, int batch, int nq, int nk, int heads, int dim_head, float* out,
                         cudaStream_t s);
int launch_spk_geglu(const float* h, size_t rows, int inner, float* out, cudaStream_t s);
int launch_spk_rmsnorm(const float* x, const float* gamma, int dim, int rows, float* out, cudaStream_t s);
int launch_spk_fsq_quantize(const float* x, int dim, int rows, const float* w_in, const float* b_in, const int* levels,
                            int n_levels, int* idx_out, float* margin_out, cudaStream_t s);

// ---- small host helpers ------------------------------------------------------------------------
inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
  uint32_t r = u + 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(r >> 16);
}
inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

}  // namespace sparkcodec
