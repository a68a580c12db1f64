// Micro-benchmark: cost of mbarrier primitives on sm_100a (single thread, cycles per operation).
#include <cstdio>
#include "tc_ptx.cuh"
using namespace sparkcodec;
namespace sparkcodec { void set_error(const char*, ...) {} thread_local int64_t* g_launch_counter = nullptr; }

__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void k(long long* out) {
  __shared__ uint64_t bar[4];
  __shared__ uint64_t pp[2];
  const uint32_t b0 = smem_u32(&bar[0]);
  if (threadIdx.x == 0) { mbar_init(b0, 1); mbar_init(smem_u32(&pp[0]), 1); mbar_init(smem_u32(&pp[1]), 1); fence_barrier_init(); }
  __syncthreads();
  const int N = 2000;
  if (threadIdx.x == 0) {
    // (a) try_wait on a phase that is already complete (parity 1 of a fresh barrier = "previous phase")
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < N; ++i) acc += mbar_try_wait(b0, 1u);
    long long t1 = clock64();
    out[0] = (t1 - t0); out[7] = acc;
    // (b) test_wait, same
    t0 = clock64();
    for (int i = 0; i < N; ++i) acc += mbar_test_wait(b0, 1u);
    t1 = clock64();
    out[1] = (t1 - t0); out[7] += acc;
    // (c) arrive + try_wait on own barrier (count 1): completes a phase each iteration
    t0 = clock64();
    for (int i = 0; i < N; ++i) { mbar_arrive(b0); while (!mbar_try_wait(b0, (uint32_t)i & 1u)) {} }
    t1 = clock64();
    out[2] = (t1 - t0);
  }
  __syncthreads();
  // (d) ping-pong between two warps through two mbarriers: round-trip latency of arrive -> peer's try_wait wake-up
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) { mbar_arrive(smem_u32(&pp[0])); while (!mbar_try_wait(smem_u32(&pp[1]), (uint32_t)i & 1u)) {} }
    out[3] = clock64() - t0;
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < N; ++i) { while (!mbar_try_wait(smem_u32(&pp[0]), (uint32_t)i & 1u)) {} mbar_arrive(smem_u32(&pp[1])); }
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  k<<<1, 64>>>(d); cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("try_wait (ready)      : %6.1f cycles\n", h[0] / 2000.0);
  printf("test_wait (ready)     : %6.1f cycles\n", h[1] / 2000.0);
  printf("arrive + try_wait self: %6.1f cycles\n", h[2] / 2000.0);
  printf("ping-pong round trip  : %6.1f cycles (two arrive->wake hops)\n", h[3] / 2000.0);
  return 0;
}
