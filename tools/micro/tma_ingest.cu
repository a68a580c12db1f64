// Micro-benchmark: TMA operation throughput / bytes per SM-cycle when every SM streams the SAME small
// (L2-resident) weight matrix over and over (the weight re-streaming pattern of the conv kernels).
// Variants: number of producer threads, tensor map in kernel-parameter space vs global memory, box size.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../spark-tts_b200/csrc -o tma_ingest tma_ingest.cu
#include <cstdio>
#include <vector>
#include "tc_ptx.cuh"
using namespace sparkcodec;
namespace sparkcodec { void set_error(const char*, ...) {} thread_local int64_t* g_launch_counter = nullptr; }

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int STAGES = 12;

// NP producer warps (one elected thread each) fill disjoint stage subsets; one consumer thread releases stages.
template <int NP>
__global__ void __launch_bounds__(32 * (NP + 1), 1) ingest(const __grid_constant__ CUtensorMap tm_param, const CUtensorMap* tm_glob,
                                                           int box_rows, int k_chunks, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const uint32_t stage_bytes = box_rows * 64;
  const CUtensorMap* tm = tm_glob ? tm_glob : &tm_param;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    fence_barrier_init();
    prefetch_tmap(tm);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp < NP) {
    if (elect_one()) {
      uint32_t ph = 0;
      int kc = 0;
      for (int i = warp; i < iters; i += NP) {
        const int s = i % STAGES;
        ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
        mbar_expect_tx(smem_u32(&full[s]), stage_bytes);
        tma_load_2d(base + s * stage_bytes, tm, smem_u32(&full[s]), kc * 32, 0);
        if (++kc == k_chunks) kc = 0;
      }
    }
  } else if (threadIdx.x == 32 * NP) {
    uint32_t s = 0, ph = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      mbar_wait(smem_u32(&full[s]), ph);
      mbar_arrive(smem_u32(&empty[s]));
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}

// Pure TMA throughput: one thread issues `batch` loads per mbarrier phase into distinct smem slots and waits
// once per batch (no per-stage consumer handshake).
__global__ void __launch_bounds__(32, 1) burst(const __grid_constant__ CUtensorMap tm, int box_rows, int k_chunks, int batch,
                                               int rounds, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  const uint32_t stage_bytes = box_rows * 64;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); prefetch_tmap(&tm); }
  __syncwarp();
  if (threadIdx.x == 0) {
    int kc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      mbar_expect_tx(smem_u32(&bar), stage_bytes * batch);
      for (int i = 0; i < batch; ++i) {
        tma_load_2d(base + (uint32_t)i * stage_bytes, &tm, smem_u32(&bar), kc * 32, 0);
        if (++kc == k_chunks) kc = 0;
      }
      mbar_wait(smem_u32(&bar), (uint32_t)r & 1u);
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_enc enc = (PFN_enc)fn;
  const int K = 7 * 192, N = 192;                      // a (192, 1344) bf16 weight matrix = 516 KB, L2 resident
  __nv_bfloat16* w; cudaMalloc(&w, (size_t)K * N * 2); cudaMemset(w, 0, (size_t)K * N * 2);
  long long* d; cudaMalloc(&d, 64);
  CUtensorMap* tmg; cudaMalloc(&tmg, sizeof(CUtensorMap));
  const int iters = 24000;
  cudaFuncSetAttribute(ingest<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(ingest<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(ingest<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(burst, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int box_rows : {48, 96, 192}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N}, strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, es[2] = {1, 1};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int grid : {1, 148}) {
      const int batch = 192 * 1024 / (box_rows * 64) > 32 ? 32 : 192 * 1024 / (box_rows * 64), rounds = 400;
      burst<<<grid, 32, 200 * 1024>>>(tm, box_rows, K / 32, batch, rounds, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      printf("burst box %3d x 64 B batch=%2d grid=%3d : %6.1f B/clk/SM  (%6.0f cycles/box incl. 1 wait per batch) %s\n", box_rows, batch,
             grid, (double)rounds * batch * box_rows * 64 / c, (double)c / rounds / batch, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  for (int box_rows : {48, 192}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N}, strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, es[2] = {1, 1};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaMemcpy(tmg, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    for (int glob = 0; glob < 2; ++glob)
      for (int np : {1, 2, 4})
        for (int grid : {1, 148}) {
          const CUtensorMap* g = glob ? tmg : nullptr;
          if (np == 1) ingest<1><<<grid, 64, 200 * 1024>>>(tm, g, box_rows, K / 32, iters, d);
          if (np == 2) ingest<2><<<grid, 96, 200 * 1024>>>(tm, g, box_rows, K / 32, iters, d);
          if (np == 4) ingest<4><<<grid, 160, 200 * 1024>>>(tm, g, box_rows, K / 32, iters, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("box %3d x 64 B (%5d B) map=%s producers=%d grid=%3d : %6.1f B/clk/SM  (%6.0f cycles/box) %s\n", box_rows,
                 box_rows * 64, glob ? "global" : "param ", np, grid, (double)iters * box_rows * 64 / c, (double)c / iters,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  }
  return 0;
}
