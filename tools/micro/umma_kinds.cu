// Micro-benchmark: back-to-back tcgen05.mma issue rate (cta_group::1, M = 128) of kind::f16 (K = 16) and kind::f8f6f4
// (e5m2, K = 32) as a function of N, and of the two kinds interleaved the way the two-term fp32 mode issues them
// (2 x f16 then 2 x f8 per 32-channel chunk, or in longer runs).  Operands resident in shared memory (no TMA traffic).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../spark-tts_b200/csrc -o umma_kinds umma_kinds.cu
#include <cstdio>
#include "tc_ptx.cuh"
using namespace sparkcodec;

namespace sparkcodec { void set_error(const char*, ...) {} thread_local int64_t* g_launch_counter = nullptr; }

// MODE 0: f16 only; 1: f8 only; 2: f16 f16 f8 f8 repeating; 3: 8 x f16 then 8 x f8 repeating; 4: f8 with A from tensor memory
template <int N, int BK, int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t barp = smem_u32(&bar);
  for (int i = threadIdx.x; i < (4 * 128 + 4 * 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(barp, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&tslot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    constexpr uint32_t id16 = make_idesc<N, 0>(), id8 = make_idesc<N, 1>();
    const uint32_t a0 = base, b0 = base + 4 * 128 * BK * 2;
    uint64_t ad[4], bd[4];
    for (int q = 0; q < 4; ++q) {
      ad[q] = make_smem_desc<BK>(a0 + (uint32_t)q * 128 * BK * 2);
      bd[q] = make_smem_desc<BK>(b0 + (uint32_t)q * N * BK * 2);
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i += 16) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const uint32_t d = tmem + (uint32_t)((u >> 2) & 1) * N;
        const uint64_t a = ad[u & 3] + 2 * (u & 1), b = bd[u & 3] + 2 * (u & 1);
        const bool f8 = MODE == 1 || MODE == 4 || (MODE == 2 && (u & 2)) || (MODE == 3 && (u & 8));
        if (MODE == 4) umma_f8_ts(d, tmem + 2 * N + 8 * (u & 3), b, id8, 1u);
        else if (f8) umma_f8(d, a, b, id8, 1u);
        else umma_bf16(d, a, b, id16, 1u);
      }
    }
    umma_commit(barp);
    mbar_wait(barp, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int N, int BK, int MODE>
void run(long long* d) {
  const int iters = 16000;
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(rate_kernel<N, BK, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"f16 only", "f8 only", "2xf16,2xf8", "8xf16,8xf8", "f8 A in TMEM"};
  rate_kernel<N, BK, MODE><<<148, 128, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d rows=%3dB %-14s: %7.1f cycles/MMA %s\n", N, BK * 2, names[MODE], (double)c / iters,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}
template <int N, int BK>
void all(long long* d) { run<N, BK, 0>(d); run<N, BK, 1>(d); run<N, BK, 2>(d); run<N, BK, 3>(d); run<N, BK, 4>(d); }
int main() {
  long long* d; cudaMalloc(&d, 64);
  all<64, 32>(d); all<96, 32>(d); all<128, 32>(d); all<192, 32>(d); all<256, 32>(d);
  all<96, 64>(d); all<192, 64>(d); all<256, 64>(d);
  return 0;
}
