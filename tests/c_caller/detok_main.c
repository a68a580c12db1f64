/* A reference-independent C caller of libsparkcodec.so: no Python, no torch -- only the C ABI of
 * include/sparkcodec.h plus the CUDA runtime for device buffers.  Used by tests/test_gpu_c_caller.py.
 *
 *   detok_main <model.bin> <tokens.bin> <out.f32> [fp32|bf16]
 *
 * model.bin : sparkcodec_config (raw struct), then records { u32 key_len, key bytes, u32 ndim, i64 dims[ndim],
 *             f32 data[prod(dims)] } until end of file (the tensors of BiCodec/model.safetensors by key).
 * tokens.bin: i32 batch, i32 frames, i64 semantic[batch*frames], i32 global[batch*token_num]
 *             (the dtypes of the ONNX bicodec_vocoder contract, export_sparktts_onnx.py:819-840).
 * out.f32   : batch * hop * frames floats = BiCodec.detokenize's (B, 1, hop*T) waveform (bicodec.py:171-189).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "sparkcodec.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ != 0) {                                                                  \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, sparkcodec_last_error());        \
      return 2;                                                                      \
    }                                                                                \
  } while (0)
#define CUDA(call)                                                                   \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      fprintf(stderr, "%s -> %s\n", #call, cudaGetErrorString(e_));                  \
      return 3;                                                                      \
    }                                                                                \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s model.bin tokens.bin out.f32 [fp32|bf16]\n", argv[0]);
    return 1;
  }
  const int precision = (argc > 4 && strcmp(argv[4], "bf16") == 0) ? SPARKCODEC_PREC_BF16 : SPARKCODEC_PREC_FP32;
  if (sparkcodec_abi_version() != SPARKCODEC_ABI_VERSION) {
    fprintf(stderr, "ABI mismatch\n");
    return 1;
  }

  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  sparkcodec_config cfg;
  if (fread(&cfg, sizeof(cfg), 1, f) != 1) { fprintf(stderr, "short model file\n"); return 1; }
  sparkcodec_handle* h = NULL;
  CHECK(sparkcodec_create(&cfg, 0, &h));
  int n_tensors = 0;
  for (;;) {
    uint32_t key_len, ndim;
    if (fread(&key_len, 4, 1, f) != 1) break;
    char key[512];
    int64_t dims[4];
    if (key_len >= sizeof(key) || fread(key, 1, key_len, f) != key_len || fread(&ndim, 4, 1, f) != 1 || ndim > 4 ||
        fread(dims, 8, ndim, f) != ndim) {
      fprintf(stderr, "corrupt record %d\n", n_tensors);
      return 1;
    }
    key[key_len] = 0;
    size_t n = 1;
    for (uint32_t i = 0; i < ndim; ++i) n *= (size_t)dims[i];
    float* data = (float*)malloc(n * sizeof(float) + 4);
    if (!data || fread(data, sizeof(float), n, f) != n) { fprintf(stderr, "short tensor %s\n", key); return 1; }
    CHECK(sparkcodec_set_tensor(h, key, data, dims, (int)ndim));
    free(data);
    ++n_tensors;
  }
  fclose(f);
  CHECK(sparkcodec_finalize(h));

  f = fopen(argv[2], "rb");
  if (!f) { perror(argv[2]); return 1; }
  int32_t bt[2];
  if (fread(bt, 4, 2, f) != 2) return 1;
  const int B = bt[0], T = bt[1];
  const size_t n_sem = (size_t)B * T, n_glob = (size_t)B * cfg.token_num;
  int64_t* sem = (int64_t*)malloc(n_sem * 8 + 8);
  int32_t* glob = (int32_t*)malloc(n_glob * 4 + 4);
  if (fread(sem, 8, n_sem, f) != n_sem || fread(glob, 4, n_glob, f) != n_glob) { fprintf(stderr, "short token file\n"); return 1; }
  fclose(f);

  int hop = 1;
  for (int i = 0; i < cfg.num_upsample; ++i) hop *= cfg.rates[i];
  const size_t n_wav = (size_t)B * hop * T;
  size_t ws_bytes = 0;
  CHECK(sparkcodec_workspace_bytes(h, B, T, &ws_bytes));
  void *d_sem, *d_glob, *d_ws;
  float* d_wav;
  CUDA(cudaMalloc(&d_sem, n_sem * 8));
  CUDA(cudaMalloc(&d_glob, n_glob * 4));
  CUDA(cudaMalloc(&d_ws, ws_bytes));
  CUDA(cudaMalloc((void**)&d_wav, n_wav * sizeof(float)));
  CUDA(cudaMemcpy(d_sem, sem, n_sem * 8, cudaMemcpyHostToDevice));
  CUDA(cudaMemcpy(d_glob, glob, n_glob * 4, cudaMemcpyHostToDevice));
  cudaStream_t stream;
  CUDA(cudaStreamCreate(&stream));
  CHECK(sparkcodec_detokenize(h, d_sem, SPARKCODEC_I64, d_glob, SPARKCODEC_I32, B, T, precision, d_ws, ws_bytes, d_wav,
                              (void*)stream));
  CHECK(sparkcodec_check_tokens(h, (void*)stream));   /* synchronises; EINDEX for an out-of-range id */
  float* wav = (float*)malloc(n_wav * sizeof(float));
  CUDA(cudaMemcpy(wav, d_wav, n_wav * sizeof(float), cudaMemcpyDeviceToHost));
  int64_t launches = 0;
  CHECK(sparkcodec_launch_count(h, &launches));

  f = fopen(argv[3], "wb");
  if (!f || fwrite(wav, sizeof(float), n_wav, f) != n_wav) { perror(argv[3]); return 1; }
  fclose(f);
  printf("tensors=%d batch=%d frames=%d samples=%zu workspace=%zu launches=%lld\n", n_tensors, B, T, n_wav, ws_bytes,
         (long long)launches);
  cudaFree(d_sem); cudaFree(d_glob); cudaFree(d_ws); cudaFree(d_wav);
  cudaStreamDestroy(stream);
  CHECK(sparkcodec_destroy(h));
  return 0;
}
