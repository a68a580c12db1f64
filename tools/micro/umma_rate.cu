// Micro-benchmark: back-to-back tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) issue rate as a function
// of N and of the smem swizzle mode, operands resident in shared memory (no TMA traffic).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../spark-tts_b200/csrc -o umma_rate umma_rate.cu
#include <cstdio>
#include "tc_ptx.cuh"
using namespace sparkcodec;

namespace sparkcodec { void set_error(const char*, ...) {} thread_local int64_t* g_launch_counter = nullptr; }

template <int N, int BK>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int distinct) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t barp = smem_u32(&bar);
  // zero operands (values irrelevant for timing)
  for (int i = threadIdx.x; i < (128 + 256) * 128 * 4 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(barp, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&tslot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc<N>();
    const uint32_t a0 = base, b0 = base + 4 * 128 * BK * 2;   // up to 4 distinct A tiles then B tiles
    uint64_t ad[4], bd[4];
    for (int q = 0; q < 4; ++q) {
      ad[q] = make_smem_desc<BK>(a0 + (uint32_t)(q % distinct) * 128 * BK * 2);
      bd[q] = make_smem_desc<BK>(b0 + (uint32_t)(q % distinct) * N * BK * 2);
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) umma_bf16(tmem + (uint32_t)(u >> 2) * N, ad[u & 3] + 2 * (u & 1), bd[u & 3] + 2 * (u & 1), idesc, 1u);
    }
    umma_commit(barp);
    mbar_wait(barp, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int N, int BK>
void run(long long* d, int distinct) {
  const int iters = 20000;
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(rate_kernel<N, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148}) {
    rate_kernel<N, BK><<<grid, 128, smem>>>(d, iters, distinct);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d swizzle=%3dB distinct=%d grid=%3d : %7.1f cycles/MMA  (%5.0f flop/clk/SM)  %s\n", N, BK * 2, distinct, grid,
           (double)c / iters, 2.0 * 128 * N * 16 * iters / (double)c, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<64, 32>(d, 1); run<96, 32>(d, 1); run<128, 32>(d, 1); run<192, 32>(d, 1); run<256, 32>(d, 1);
  run<96, 64>(d, 1); run<192, 64>(d, 1); run<256, 64>(d, 1);
  run<96, 32>(d, 4); run<192, 32>(d, 4); run<192, 64>(d, 4); run<256, 64>(d, 4);
  return 0;
}
