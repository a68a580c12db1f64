/*
 * sparkcodec.h -- C ABI of the B200-native BiCodec detokenize path (libsparkcodec.so).
 *
 * Plain C: pointers, sizes and ints only; no torch / CUDA types in any signature (streams are
 * passed as void* = cudaStream_t).  Every function returns 0 on success or a negative
 * SPARKCODEC_E* code; sparkcodec_last_error() returns the thread-local message.
 *
 * The reference (arghyasur1991/Spark-TTS) is pure Python and has no FFI for this path; each entry
 * point below names the reference interface it replaces (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 */
#ifndef SPARKCODEC_H_
#define SPARKCODEC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPARKCODEC_ABI_VERSION 1

#if defined(__GNUC__)
#define SPARKCODEC_API __attribute__((visibility("default")))
#else
#define SPARKCODEC_API
#endif

enum {
  SPARKCODEC_OK = 0,
  SPARKCODEC_EINVAL = -1,    /* bad argument / shape / dtype            -> ValueError   */
  SPARKCODEC_EINDEX = -2,    /* token id outside its codebook           -> IndexError   */
  SPARKCODEC_ECUDA = -3,     /* CUDA runtime / driver failure           -> RuntimeError */
  SPARKCODEC_ESTATE = -4,    /* call order (e.g. detokenize before finalize)            */
  SPARKCODEC_ENOMEM = -5,    /* workspace too small / allocation failed                 */
  SPARKCODEC_EMISSING = -6   /* a checkpoint tensor the path needs was never set        */
};

/* token dtypes accepted for either input (ONNX contract: semantic int64, global int32;
 * the CLI passes int64 for both, Triton int32 for both -- export_sparktts_onnx.py:819-840,
 * cli/SparkTTS.py:213-234, runtime/triton_trtllm/model_repo/vocoder/config.pbtxt:28-46) */
enum { SPARKCODEC_I32 = 0, SPARKCODEC_I64 = 1 };

/* arithmetic of the dense contractions (activations and I/O are fp32 either way):
 *   FP32 : error-compensated split products on tcgen05, fp32 accumulate: an fp16 x fp16 main term plus both cross
 *          terms as e5m2 x e5m2 products at twice the rate (default), or three bf16 products
 *          (env SPARKCODEC_FP32_TERMS=3; sparkcodec_fp32_terms() reports which)
 *          (parity bound vs the reference fp32: max-abs 1e-3, SNR >= 60 dB)
 *   BF16 : single bf16 product, fp32 accumulate (looser stated bound)            */
enum { SPARKCODEC_PREC_FP32 = 0, SPARKCODEC_PREC_BF16 = 1,
       /* FP32 with the three-term bf16 split whatever the process default (~2^-17 per product, 1.5 x the tensor time):
        * what the encode side (tokenize) asks for, where a rounding flips a discrete token */
       SPARKCODEC_PREC_FP32X3 = 2 };

/* which implementation of the dense contractions runs: tcgen05 tensor cores (product path: one kernel
 * per conv, the narrow ResidualUnits as ONE fused kernel each), the CUDA-core verification kernel with
 * the same operands/epilogue (tests only), or tcgen05 with the ResidualUnit fusion switched off
 * (tests / A-B timing only). */
enum { SPARKCODEC_IMPL_TC = 0, SPARKCODEC_IMPL_SIMT = 1, SPARKCODEC_IMPL_TC_UNFUSED = 2 };

/* Mirrors the `audio_tokenizer` section of BiCodec/config.yaml consumed by
 * BiCodec.load_from_checkpoint (sparktts/models/bicodec.py:81-88). */
typedef struct sparkcodec_config {
  int32_t d_model;               /* quantizer.input_dim = prenet in/out = decoder.input_channel */
  int32_t codebook_size;         /* quantizer.codebook_size */
  int32_t codebook_dim;          /* quantizer.codebook_dim  */
  int32_t fsq_num_levels;        /* len(speaker_encoder.fsq_levels), <= 8 */
  int32_t fsq_levels[8];
  int32_t token_num;             /* speaker_encoder.token_num  */
  int32_t latent_dim;            /* speaker_encoder.latent_dim */
  int32_t vocos_dim;             /* prenet.vocos_dim */
  int32_t vocos_intermediate_dim;
  int32_t vocos_num_layers;      /* prenet.vocos_num_layers (AdaLN-conditioned backbone) */
  int32_t downsample_layers;     /* ConvNeXt layers of each downsample backbone (2) */
  int32_t num_downsample;        /* len(prenet.sample_ratios), all ratios must be 1 */
  int32_t dec_channels;          /* decoder.channels */
  int32_t num_upsample;          /* len(decoder.rates), <= 8 */
  int32_t rates[8];
  int32_t kernel_sizes[8];
} sparkcodec_config;

typedef struct sparkcodec_handle sparkcodec_handle;

SPARKCODEC_API const char* sparkcodec_last_error(void);
SPARKCODEC_API int sparkcodec_abi_version(void);

/* Replaces the module construction half of BiCodec.load_from_checkpoint
 * (sparktts/models/bicodec.py:69-98).  `device` is the CUDA ordinal. */
SPARKCODEC_API int sparkcodec_create(const sparkcodec_config* cfg, int device, sparkcodec_handle** out);
SPARKCODEC_API int sparkcodec_destroy(sparkcodec_handle* h);

/* Replaces load_state_dict (bicodec.py:100-101): hand over one checkpoint tensor by its
 * model.safetensors key (e.g. "decoder.model.1.block.1.weight_v").  `data` is fp32 HOST memory,
 * copied during the call.  Keys that the detokenize path does not use are ignored (return 0). */
SPARKCODEC_API int sparkcodec_set_tensor(sparkcodec_handle* h, const char* key, const float* data,
                          const int64_t* shape, int ndim);

/* Replaces model.eval() + remove_weight_norm() (bicodec.py:108-109, 213-221): folds weight-norm
 * (dim 0), re-lays every dense weight out for the tcgen05 kernels (K-major bf16 hi/lo planes,
 * polyphase split of the transposed convolutions), uploads, and frees the host copies.
 * Returns SPARKCODEC_EMISSING (message names the key) if a needed tensor was never set. */
SPARKCODEC_API int sparkcodec_finalize(sparkcodec_handle* h);

/* Device workspace needed by one detokenize call of (batch, frames). The library never allocates
 * per call; the caller (torch caching allocator) owns the workspace. */
SPARKCODEC_API int sparkcodec_workspace_bytes(sparkcodec_handle* h, int batch, int frames, size_t* bytes);

/* Replaces BiCodec.detokenize (sparktts/models/bicodec.py:171-189) == the ONNX `bicodec_vocoder`
 * graph (export_sparktts_onnx.py:267-312):
 *   semantic  : device ptr, (batch, frames) row-major, dtype sem_dtype, ids in [0, codebook_size)
 *   global    : device ptr, (batch, token_num) row-major [(B,1,N) and (B,N) are the same bytes]
 *   wav_out   : device ptr, (batch, hop*frames) fp32 [== (B,1,hop*T) contiguous]
 * Asynchronous on `stream`; inputs are borrowed until the stream reaches the end of the call.
 * Token ids are validated on the device: the first error is latched in the handle and reported by
 * sparkcodec_check_tokens() (the reference raises IndexError on CPU / device-asserts on CUDA). */
SPARKCODEC_API int sparkcodec_detokenize(sparkcodec_handle* h, const void* semantic, int sem_dtype,
                          const void* global_tokens, int glob_dtype, int batch, int frames,
                          int precision, void* workspace, size_t workspace_bytes, float* wav_out,
                          void* stream);

/* The two halves of detokenize, split at the prenet -> WaveGenerator boundary (bicodec.py:185-187).
 * They exist for time-sharding (BASELINE config 5): a rank runs sparkcodec_prenet on its own frame
 * window (tokens are replicated, so the window simply carries `prenet_halo` extra frames each side and
 * the caller crops), exchanges `wavegen_halo` boundary rows of x with its neighbours over NCCL, and
 * runs sparkcodec_wavegen on the widened window.  Zero padding is applied at the edges of whatever
 * window is passed, which is the true utterance edge only for the first/last shard; rows within the
 * halo of a shard edge are discarded by the caller.
 *   x_out / x_in : device, (batch, frames, d_model) fp32 channels-last = prenet(z_q, d) + d          */
SPARKCODEC_API int sparkcodec_prenet(sparkcodec_handle* h, const void* semantic, int sem_dtype,
                      const void* global_tokens, int glob_dtype, int batch, int frames, int precision,
                      void* workspace, size_t workspace_bytes, float* x_out, void* stream);
SPARKCODEC_API int sparkcodec_wavegen(sparkcodec_handle* h, const float* x_in, int batch, int frames, int precision,
                       void* workspace, size_t workspace_bytes, float* wav_out, void* stream);
/* The same WaveGenerator call with its input STAGED in pieces, so that a time-sharded rank can overlap the NCCL
 * halo exchange with the interior rows: sparkcodec_wavegen_stage converts rows [row_offset, row_offset + rows) of
 * the (batch, frames_total, d_model) window into the kernels' operand format inside `workspace` (x_rows is
 * (batch, rows, d_model) fp32 with `src_batch_stride` floats between utterances, so a slice of a wider tensor
 * needs no copy); the rank stages its own rows while the neighbours' halo rows are in flight, stages the halos when
 * they have landed, and sparkcodec_wavegen_staged then runs the WaveGenerator over the whole window.  All calls
 * of one window must pass the same (batch, frames_total, precision, workspace) and the workspace must hold the
 * whole batch (sparkcodec_workspace_bytes(batch, frames_total)); the result is bit-identical to
 * sparkcodec_wavegen on the concatenated rows. */
SPARKCODEC_API int sparkcodec_wavegen_stage(sparkcodec_handle* h, const float* x_rows, int batch, int rows,
                             int64_t src_batch_stride, int frames_total, int row_offset, int precision,
                             void* workspace, size_t workspace_bytes, void* stream);
SPARKCODEC_API int sparkcodec_wavegen_staged(sparkcodec_handle* h, int batch, int frames_total, int precision,
                              void* workspace, size_t workspace_bytes, float* wav_out, void* stream);
/* Receptive-field half-widths in token frames: prenet (57 for the released config) and
 * WaveGenerator (conservative ceil; SURVEY.md §8e measured 9.82 frames). */
SPARKCODEC_API int sparkcodec_halo_frames(sparkcodec_handle* h, int* prenet_halo, int* wavegen_halo);

/* Synchronises `stream` and returns SPARKCODEC_EINDEX if any token id seen since the last check
 * was out of range (message holds which input, position and value). */
SPARKCODEC_API int sparkcodec_check_tokens(sparkcodec_handle* h, void* stream);

/* Token feed, the step immediately BEFORE the path.  The reference decodes the LLM's
 * generated ids to text on the host and regex-matches "bicodec_semantic_(\d+)" / "bicodec_global_(\d+)"
 * (cli/SparkTTS.py:213-228, runtime/triton_trtllm/model_repo/spark_tts/1/model.py:283-295).  Those tokens
 * are contiguous added-token id ranges of the tokenizer, so the same selection is an order-preserving
 * compaction on the device:
 *   token_ids     : device, (batch, n_tokens) generated ids (prompt already trimmed), dtype id_dtype
 *   semantic_out  : device int32 (batch, n_tokens): row b holds semantic_len[b] codes in generation order
 *   global_out    : device int32 (batch, max_global): row b holds min(global_len[b], max_global) codes
 * `*_base` = tokenizer id of "<|bicodec_semantic_0|>" / "<|bicodec_global_0|>".  Stateless, asynchronous; it has no
 * handle, so it runs on the CALLER's current CUDA device (the one the pointers and the stream belong to).  Every
 * entry point that takes a handle switches to the handle's device for the call and restores the caller's. */
SPARKCODEC_API int sparkcodec_extract_codes(const void* token_ids, int id_dtype, int batch, int n_tokens,
                             int64_t semantic_base, int codebook_size, int64_t global_base, int global_size,
                             int32_t* semantic_out, int32_t* semantic_len, int32_t* global_out, int max_global,
                             int32_t* global_len, void* stream);

/* Semantic half of BiCodec.tokenize, the encode-side mirror of the path (SURVEY.md section 8f-4;
 * sparktts/models/bicodec.py:151-169: z = encoder(feat.transpose(1,2)); quantizer.tokenize(z);
 * feat_encoder.py:79-90, factorized_vector_quantize.py:147-152,169-187).  Available only when the checkpoint
 * handed to sparkcodec_set_tensor carried `encoder.*` and `quantizer.in_project.*` (else SPARKCODEC_ESTATE).
 *   feat        : device, (batch, frames, d_model) fp32 -- the wav2vec2 feature mix the reference feeds
 *   tokens_out  : device int64 (batch, frames): index of the nearest L2-normalised code, lowest index on ties
 *   margin_out  : optional device fp32 (batch, frames): second-best minus best distance, so a caller can tell
 *                 a numerical near-tie from a disagreement (NULL to skip)
 * Uses the same workspace as sparkcodec_detokenize(batch, frames).  (The speaker half is
 * sparkcodec_tokenize_speaker below.) */
SPARKCODEC_API int sparkcodec_tokenize_semantic(sparkcodec_handle* h, const float* feat, int batch, int frames,
                                 int precision, void* workspace, size_t workspace_bytes, int64_t* tokens_out,
                                 float* margin_out, void* stream);

/* Speaker half of BiCodec.tokenize (sparktts/models/bicodec.py:162-167: mel = mel_transformer(ref_wav);
 * global_tokens = speaker_encoder.tokenize(mel^T); speaker/speaker_encoder.py:100-105 = ECAPA-TDNN latent
 * (speaker/ecapa_tdnn.py:196-208) -> perceiver resampler (speaker/perceiver_encoder.py:335-350) -> ResidualFSQ indices
 * (fsq/residual_fsq.py:213-283)).  Available only when the checkpoint carried `speaker_encoder.speaker_encoder.*`,
 * `speaker_encoder.perceiver_sampler.*` and `speaker_encoder.quantizer.project_in.*` (else SPARKCODEC_ESTATE).
 *   ref_wav     : device, (batch, n_samples) fp32 -- the reference clip (BiCodecTokenizer.get_ref_clip: 6 s at 16 kHz)
 *   tokens_out  : device int32 (batch, token_num) [== the reference's (B, 1, token_num)], ids in [0, prod(fsq_levels))
 *   margin_out  : optional device fp32 (batch, token_num): distance of the closest pre-rounding FSQ coordinate to a
 *                 rounding boundary, so a caller can tell a numerical near-tie from a disagreement (NULL to skip)
 * The mel transform is torchaudio's MelSpectrogram as bicodec.py:191-211 builds it (power 1, slaney scale and norm,
 * centred reflect-padded STFT, periodic Hann window); its parameters default to the released config.yaml
 * `mel_params` (16 kHz, n_fft 1024, win 640, hop 320, 128 mels, 10 Hz .. Nyquist) and can be set BEFORE finalize
 * with sparkcodec_set_mel_params (f_max < 0 = Nyquist).  Everything runs in fp32 FMA arithmetic. */
SPARKCODEC_API int sparkcodec_set_mel_params(sparkcodec_handle* h, int sample_rate, int n_fft, int win_length,
                              int hop_length, int n_mels, float f_min, float f_max);
SPARKCODEC_API int sparkcodec_speaker_workspace_bytes(sparkcodec_handle* h, int batch, int n_samples, size_t* bytes);
SPARKCODEC_API int sparkcodec_tokenize_speaker(sparkcodec_handle* h, const float* ref_wav, int batch, int n_samples,
                                void* workspace, size_t workspace_bytes, int32_t* tokens_out, float* margin_out,
                                void* stream);

/* ---- test / profiling hooks (not needed by a drop-in user) --------------------------------- */

/* sparkcodec_tokenize_speaker that also copies an intermediate ("mel" (T, n_mels), "ecapa_latent" (T, 1536),
 * "perceiver" (token_num, latent_dim)) as fp32 (batch, rows, channels) into tap_out. */
SPARKCODEC_API int sparkcodec_tokenize_speaker_tap(sparkcodec_handle* h, const float* ref_wav, int batch, int n_samples,
                                    void* workspace, size_t workspace_bytes, int32_t* tokens_out, const char* tap,
                                    float* tap_out, size_t tap_capacity, int64_t* tap_shape, void* stream);


/* Selects tcgen05 (default) or the CUDA-core verification kernels for the dense contractions. */
SPARKCODEC_API int sparkcodec_set_impl(sparkcodec_handle* h, int impl);

/* Runs detokenize but also copies the activation named `tap` (oracle tap names, e.g. "d_vector",
 * "prenet.downsample.0", "decoder.model.1.block.2") as fp32 (batch, rows, channels) into tap_out
 * (device, capacity tap_capacity floats); writes its rows/channels to tap_shape[0..1]. */
SPARKCODEC_API int sparkcodec_detokenize_tap(sparkcodec_handle* h, const void* semantic, int sem_dtype,
                              const void* global_tokens, int glob_dtype, int batch, int frames,
                              int precision, void* workspace, size_t workspace_bytes, float* wav_out,
                              const char* tap, float* tap_out, size_t tap_capacity,
                              int64_t* tap_shape, void* stream);

/* Stand-alone dense convolution through the same packing + kernels the model uses.
 *   kind 0: Conv1d weight (C_out, C_in, k), dilation `param`, padding (k-1)/2*dilation
 *   kind 1: ConvTranspose1d weight (C_in, C_out, k), stride `param`, padding (k-stride)/2
 * x_dev (batch, L, C_in) fp32 channels-last; y_dev (batch, L_out, C_out) fp32.
 * act: 0 none, 1 GELU(erf), 2 snake(alpha_host[C_out]); residual_dev optional (same shape as y). */
SPARKCODEC_API int sparkcodec_op_conv(int device, int kind, const float* w_host, const int64_t* wshape,
                       const float* bias_host, int param, int batch, int L, const float* x_dev,
                       float* y_dev, const float* residual_dev, int act, const float* alpha_host,
                       int precision, int impl, void* stream);

/* Host-side weight re-layout, exposed so it can be checked without a GPU.
 * Fills w_hi/w_lo (uint16 bf16 bit patterns, (n_total, kt*c_in) row-major), shifts (n_phase*kt,
 * unused = INT32_MIN) and returns kt / n_phase / n_total through the out params. */
SPARKCODEC_API int sparkcodec_pack_conv(int kind, const float* w_host, const int64_t* wshape, int param,
                         uint16_t* w_hi, uint16_t* w_lo, size_t w_capacity, int32_t* shifts,
                         int32_t* ntaps, int32_t* kt, int32_t* n_phase, int32_t* n_total);

/* Profile mode: records a CUDA-event pair around every kernel launch of subsequent calls (adds launch
 * gaps; never on while a throughput number is taken).  sparkcodec_profile_read synchronises and writes
 * one line per launch: "name\tms\talgorithmic_flops\talgorithmic_bytes\n" (buf may be NULL to query
 * the size).  Enabling/disabling clears the records. */
SPARKCODEC_API int sparkcodec_profile(sparkcodec_handle* h, int enable);
SPARKCODEC_API int sparkcodec_profile_read(sparkcodec_handle* h, char* buf, size_t cap, size_t* needed);

/* Number of this library's kernel launches issued since the handle was created (bench.py's
 * `gpu_launches`). */
SPARKCODEC_API int sparkcodec_launch_count(sparkcodec_handle* h, int64_t* count);

/* Test hook, no GPU needed: like sparkcodec_pack_conv, but returns the weight planes of the two-term fp32 mode --
 * w_h16 = fp16(W) bits and w_p8 = per row and group of 32 K values, 64 bytes [ e5m2(fp16(W) * 2^-4) x 32 |
 * e5m2((W - fp16(W)) * 2^8) x 32 ], both (n_total, kt*c_in) x 2 bytes.  EINVAL when kt*c_in is not a multiple of 32. */
SPARKCODEC_API int sparkcodec_pack_conv_f16f8(int kind, const float* w_host, const int64_t* wshape, int param,
                                              uint16_t* w_h16, uint16_t* w_p8, size_t w_capacity);

/* Tensor-core products per multiply-accumulate of the FP32 precision mode in this process: 2 (fp16 main term + two
 * e5m2 cross terms = 2 bf16-MMA equivalents of tensor time) or 3 (bf16 x 3).  bench.py's `tensor_work_factor`. */
SPARKCODEC_API int sparkcodec_fp32_terms(void);

/* Test hook, no GPU needed: the N tile width the tcgen05 conv kernel uses for a layer with `n_total` output columns in
 * polyphase branches of `cols_per_phase` columns when a launch has `m_tiles` 128-row tiles on a device with `num_sms`
 * SMs: the packed width (largest of 256/192/128/96/64 that divides cols_per_phase), narrowed to 128/96/64 while the
 * launch would keep at most half of the SMs busy.  Returns the width, or EINVAL when no width divides cols_per_phase. */
SPARKCODEC_API int sparkcodec_tile_width(int n_total, int cols_per_phase, int m_tiles, int num_sms);

#ifdef __cplusplus
}
#endif
#endif /* SPARKCODEC_H_ */
