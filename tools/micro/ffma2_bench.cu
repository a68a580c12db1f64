// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a, alone and mixed
// with ALU work.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float a[16]; unsigned long long p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
  unsigned long long ss = pk(s, s), cc = pk(0.5f, 0.25f);
  unsigned int z = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.5f);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, cc);
    } else if (MODE == 2) {   // 16 FFMA + 16 LOP
#pragma unroll
      for (int i = 0; i < 16; ++i) { a[i] = fmaf(a[i], s, 0.5f); z = (z ^ (z << 1)) + i; }
    } else {                  // 8 FFMA2 + 16 LOP
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, cc); z = (z ^ (z << 1)) + i; z = (z ^ (z << 1)) + i + 8; }
    }
  }
  float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
  for (int i = 0; i < 8; ++i) r += (float)(p[i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + z;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  const char* names[4] = {"16 FFMA", "8 FFMA2", "16 FFMA + 32 ALU", "8 FFMA2 + 32 ALU"};
  for (int m = 0; m < 4; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (m == 0) k<0><<<148 * 4, 256>>>(d, iters, 0.999f);
      if (m == 1) k<1><<<148 * 4, 256>>>(d, iters, 0.999f);
      if (m == 2) k<2><<<148 * 4, 256>>>(d, iters, 0.999f);
      if (m == 3) k<3><<<148 * 4, 256>>>(d, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16 * iters * 148.0 * 4 * 256;
    printf("%-20s %8.3f ms  %7.1f TFLOP/s (fp32 fma flops)\n", names[m], ms, fl / ms / 1e9);
  }
  return 0;
}
