// Fused ResidualUnit of the WaveGenerator for the narrow stages (C = 96 / 192 channels), one kernel:
//
//   mid   = snake_mid( conv_k7_dil( op_in ) + b7 )          op_in = snake_in(x) as bf16 hi/lo planes
//   x     = x + conv_1x1( mid ) + b1                         fp32 residual stream, updated in place
//   op_out = snake_next(x)                                   operand planes of the next layer (optional)
//
// (reference: ResidualUnit.forward, sparktts/modules/blocks/layers.py:51-67; the Snake of the NEXT unit /
// up-sampler is applied here because it is element-wise on this unit's output.)
//
// Un-fused, the k7 conv writes `mid` to HBM and the 1x1 conv reads it back; at C <= 192 the 1x1 conv is
// purely HBM-bound (16 B per output element).  Here `mid` never leaves the SM:
//
//   TMEM acc1 (k7 result) --epilogue warps: +b7, Snake, bf16 hi/lo split--> smem K-major SWIZZLE_64B chunks
//        --tcgen05.mma (A = mid chunk, B = W1 chunk, split-K over the chunks)--> TMEM acc2
//        --epilogue warps: + residual slab (TMA) + b1, store x, Snake, store operand planes.
//
// Roles (352 threads, one persistent CTA per SM):
//   warp 0 / one lane : TMA producer: activation halo tiles (one per K group, shared by the 7 taps through
//                       row-offset descriptors), W7 tap tiles and W1 tiles through one weight ring
//   warp 1 / one lane : tcgen05.mma issuer (k7 conv of tile i, then the 1x1 conv of tile i - SKEW)
//   warps 2..9        : epilogue: mid stage of tile i, final stage of tile i - SKEW
//   warp 10 / one lane: residual TMA producer (fp32 128 x 32 slabs, SWIZZLE_128B)
// C = 96 keeps two acc1 and two acc2 buffers in TMEM and skews the 1x1 conv by one tile (SKEW = 1) so the
// tensor pipe never waits for the mid stage; C = 192 (acc1 + acc2 = 384 columns) runs un-skewed.
#include "gemm_params.cuh"
#include "tc_ptx.cuh"

namespace sparkcodec {
namespace {

constexpr int kRuHaloRowsMax = 184;                 // 128 + 6 * 9, multiple of 8
constexpr int kRuAChunkBytes = 12 * 1024;           // 184 rows x 64 B, rounded up to 1024
constexpr int kRuMidChunkBytes = kBlockM * 64;      // 128 rows x 32 bf16 (one plane)
constexpr int kRuSlabBytes = kBlockM * 128;         // 128 rows x 32 fp32
constexpr int kRuEpiWarps = 8;
constexpr int kRuEpiThreads = kRuEpiWarps * 32;
constexpr int kRuResWarp = 2 + kRuEpiWarps;
constexpr int kRuThreads = (3 + kRuEpiWarps) * 32;

struct RuParams {
  int batch, L, dil, halo_rows;
  int m_tiles_per_utt, num_tiles;
  const float *bias7, *alpha_mid, *inv_mid;   // [C]
  const float *bias1, *alpha_out, *inv_out;   // [C]; alpha_out/inv_out unused when out_hi == null
  float* x;                                   // (batch, L, C) fp32, read as residual and overwritten
  __nv_bfloat16 *out_hi, *out_lo;             // (batch, L, C) operand planes of the next layer (optional)
};

template <int C, int NTERMS>
struct RuCfg {
  static constexpr int kPlanes = NTERMS == 3 ? 2 : 1;
  static constexpr int kChunks = C / 32;                           // K chunks of 32 channels
  static constexpr int G = NTERMS == 3 ? 1 : (C == 96 ? 3 : 2);   // K chunks per ring stage
  static constexpr int kGroups = kChunks / G;
  static constexpr int kAStage = G * kPlanes * kRuAChunkBytes;
  static constexpr int kWChunk = C * 64;                           // C rows x 64 B (one plane)
  static constexpr int kWStage = G * kPlanes * kWChunk;
  static constexpr int kMidSlot = kPlanes * kRuMidChunkBytes;
  static constexpr int NB1 = C <= 96 ? 2 : 1;                      // acc1 / acc2 buffers in TMEM
  static constexpr int NB2 = NB1;
  static constexpr int SKEW = NB1 - 1;
  // ring depths (227 KB budget)
  static constexpr int SA = (NTERMS == 3) ? (C <= 96 ? 3 : 2) : (C <= 96 ? 2 : 3);
  static constexpr int SW = (C <= 96) ? 4 : 3;
  static constexpr int SM = (NTERMS == 3) ? (C <= 96 ? 4 : 3) : 4;
  static constexpr int SR = (NTERMS == 3 && C > 96) ? 3 : 2;
  static constexpr int kParBytes = 3 * C * 4;
  static constexpr int kNumBars = 2 * SA + 2 * SW + 2 * SM + NB1 + 2 * NB2 + 2 * SR;
  static constexpr int kSmemBytes = SA * kAStage + SW * kWStage + SM * kMidSlot + SR * kRuSlabBytes + kParBytes +
                                    kNumBars * 8 + 16 + 1024 /* alignment */;
  static_assert(kChunks % G == 0, "K groups must tile the channels");
  static_assert(SM >= G + SKEW, "mid ring too shallow");
  static_assert((NB1 + NB2) * C <= 512, "accumulators must fit TMEM");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
  static_assert(kWChunk % 1024 == 0, "weight chunks must stay 1024 B aligned");
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }

template <int C, int NTERMS>
__global__ void __launch_bounds__(kRuThreads, 1)
resunit_fused_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                     const __grid_constant__ CUtensorMap tm_w7_hi, const __grid_constant__ CUtensorMap tm_w7_lo,
                     const __grid_constant__ CUtensorMap tm_w1_hi, const __grid_constant__ CUtensorMap tm_w1_lo,
                     const __grid_constant__ CUtensorMap tm_res, const RuParams p) {
  using Cfg = RuCfg<C, NTERMS>;
  constexpr int SA = Cfg::SA, SW = Cfg::SW, SM = Cfg::SM, SR = Cfg::SR, NB1 = Cfg::NB1, NB2 = Cfg::NB2;
  constexpr int G = Cfg::G, SKEW = Cfg::SKEW;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t w_base = a_base + SA * Cfg::kAStage;
  const uint32_t mid_base = w_base + SW * Cfg::kWStage;
  const uint32_t res_base = mid_base + SM * Cfg::kMidSlot;
  const uint32_t par_base = res_base + SR * kRuSlabBytes;
  const uint32_t bar_base = par_base + Cfg::kParBytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (SA + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * SA + SW + s); };
  auto mid_full = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + s); };
  auto mid_empty = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + SM + s); };
  auto acc1_full = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + 2 * SM + s); };
  auto acc2_full = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + 2 * SM + NB1 + s); };
  auto acc2_empty = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + 2 * SM + NB1 + NB2 + s); };
  auto res_full = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + 2 * SM + NB1 + 2 * NB2 + s); };
  auto res_empty = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + 2 * SM + NB1 + 2 * NB2 + SR + s); };
  const uint32_t tmem_slot = bar_base + 8u * Cfg::kNumBars;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* s_par = reinterpret_cast<float*>(smem_raw + (par_base - smem_u32(smem_raw)));   // [bias7 | alpha_mid | inv_mid]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_my = ((int)blockIdx.x < p.num_tiles) ? (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_par[i] = __ldg(p.bias7 + i);
    s_par[C + i] = __ldg(p.alpha_mid + i);
    s_par[2 * C + i] = __ldg(p.inv_mid + i);
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a_hi);
    prefetch_tmap(&tm_w7_hi);
    prefetch_tmap(&tm_w1_hi);
    if (NTERMS == 3) {
      prefetch_tmap(&tm_a_lo);
      prefetch_tmap(&tm_w7_lo);
      prefetch_tmap(&tm_w1_lo);
    }
    for (int s = 0; s < SA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < SW; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < SM; ++s) { mbar_init(mid_full(s), kRuEpiThreads); mbar_init(mid_empty(s), 1); }
    for (int s = 0; s < NB1; ++s) mbar_init(acc1_full(s), 1);
    for (int s = 0; s < NB2; ++s) { mbar_init(acc2_full(s), 1); mbar_init(acc2_empty(s), kRuEpiThreads); }
    for (int s = 0; s < SR; ++s) { mbar_init(res_full(s), 1); mbar_init(res_empty(s), kRuEpiThreads); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      uint32_t as = 0, aph = 0, ws = 0, wph = 0;
      const uint32_t a_tx = (uint32_t)(G * Cfg::kPlanes) * (uint32_t)p.halo_rows * 64u;
      for (int it = 0; it < n_my + SKEW; ++it) {
        if (it < n_my) {
          const int tile = blockIdx.x + it * gridDim.x;
          const int b = tile / p.m_tiles_per_utt;
          const int row0 = (tile % p.m_tiles_per_utt) * kBlockM - 3 * p.dil;
          for (int kg = 0; kg < Cfg::kGroups; ++kg) {
            mbar_wait(a_empty(as), aph ^ 1u);
            mbar_expect_tx(a_full(as), a_tx);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const uint32_t sa = a_base + as * Cfg::kAStage + (uint32_t)(g * Cfg::kPlanes) * kRuAChunkBytes;
              tma_load_3d(sa, &tm_a_hi, a_full(as), (kg * G + g) * 32, row0, b);
              if (NTERMS == 3) tma_load_3d(sa + kRuAChunkBytes, &tm_a_lo, a_full(as), (kg * G + g) * 32, row0, b);
            }
            if (++as == SA) { as = 0; aph ^= 1u; }
            for (int j = 0; j < 7; ++j) {
              mbar_wait(w_empty(ws), wph ^ 1u);
              mbar_expect_tx(w_full(ws), Cfg::kWStage);
#pragma unroll
              for (int g = 0; g < G; ++g) {
                const uint32_t sw = w_base + ws * Cfg::kWStage + (uint32_t)(g * Cfg::kPlanes) * Cfg::kWChunk;
                tma_load_2d(sw, &tm_w7_hi, w_full(ws), j * C + (kg * G + g) * 32, 0);
                if (NTERMS == 3) tma_load_2d(sw + Cfg::kWChunk, &tm_w7_lo, w_full(ws), j * C + (kg * G + g) * 32, 0);
              }
              if (++ws == SW) { ws = 0; wph ^= 1u; }
            }
          }
        }
        if (it >= SKEW) {
          for (int kg = 0; kg < Cfg::kGroups; ++kg) {
            mbar_wait(w_empty(ws), wph ^ 1u);
            mbar_expect_tx(w_full(ws), Cfg::kWStage);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const uint32_t sw = w_base + ws * Cfg::kWStage + (uint32_t)(g * Cfg::kPlanes) * Cfg::kWChunk;
              tma_load_2d(sw, &tm_w1_hi, w_full(ws), (kg * G + g) * 32, 0);
              if (NTERMS == 3) tma_load_2d(sw + Cfg::kWChunk, &tm_w1_lo, w_full(ws), (kg * G + g) * 32, 0);
            }
            if (++ws == SW) { ws = 0; wph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc<C>();
      uint32_t as = 0, aph = 0, ws = 0, wph = 0, ms = 0, mph = 0;
      for (int it = 0; it < n_my + SKEW; ++it) {
        if (it < n_my) {
          // ---- k7 (dilated) conv of tile `it` -> acc1[it % NB1].  The buffer is free: the 1x1 conv of the
          // tile that last used it was issued earlier in program order and waited for every mid chunk, i.e.
          // for the epilogue warps to have drained it.
          const uint32_t d1 = tmem_base + (uint32_t)(it % NB1) * C;
          for (int kg = 0; kg < Cfg::kGroups; ++kg) {
            mbar_wait(a_full(as), aph);
            const uint32_t sa = a_base + as * Cfg::kAStage;
            for (int j = 0; j < 7; ++j) {
              mbar_wait(w_full(ws), wph);
              tc_fence_after();
              const uint32_t sw = w_base + ws * Cfg::kWStage;
              const uint32_t a_off = (uint32_t)(j * p.dil) * 64u;   // tap j starts j*dil rows into the halo tile
#pragma unroll
              for (int g = 0; g < G; ++g) {
                const uint64_t a_hi = make_smem_desc<32>(sa + (uint32_t)(g * Cfg::kPlanes) * kRuAChunkBytes + a_off);
                const uint64_t w_hi = make_smem_desc<32>(sw + (uint32_t)(g * Cfg::kPlanes) * Cfg::kWChunk);
#pragma unroll
                for (int k = 0; k < 2; ++k) umma_bf16(d1, a_hi + 2 * k, w_hi + 2 * k, idesc, (kg | j | g | k) != 0);
                if (NTERMS == 3) {
                  const uint64_t a_lo = make_smem_desc<32>(sa + (uint32_t)(g * Cfg::kPlanes + 1) * kRuAChunkBytes + a_off);
                  const uint64_t w_lo = make_smem_desc<32>(sw + (uint32_t)(g * Cfg::kPlanes + 1) * Cfg::kWChunk);
#pragma unroll
                  for (int k = 0; k < 2; ++k) umma_bf16(d1, a_lo + 2 * k, w_hi + 2 * k, idesc, 1u);
#pragma unroll
                  for (int k = 0; k < 2; ++k) umma_bf16(d1, a_hi + 2 * k, w_lo + 2 * k, idesc, 1u);
                }
              }
              umma_commit(w_empty(ws));
              if (++ws == SW) { ws = 0; wph ^= 1u; }
            }
            umma_commit(a_empty(as));
            if (++as == SA) { as = 0; aph ^= 1u; }
          }
          umma_commit(acc1_full(it % NB1));
        }
        if (it >= SKEW) {
          // ---- 1x1 conv of tile jt: split-K over the mid chunks as the epilogue warps publish them
          const int jt = it - SKEW;
          const uint32_t d2 = tmem_base + (uint32_t)(NB1 + jt % NB2) * C;
          mbar_wait(acc2_empty(jt % NB2), ((uint32_t)(jt / NB2) & 1u) ^ 1u);
          tc_fence_after();
          for (int kg = 0; kg < Cfg::kGroups; ++kg) {
            mbar_wait(w_full(ws), wph);
            const uint32_t sw = w_base + ws * Cfg::kWStage;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              mbar_wait(mid_full(ms), mph);
              tc_fence_after();
              const uint32_t sm = mid_base + ms * Cfg::kMidSlot;
              const uint64_t a_hi = make_smem_desc<32>(sm);
              const uint64_t w_hi = make_smem_desc<32>(sw + (uint32_t)(g * Cfg::kPlanes) * Cfg::kWChunk);
#pragma unroll
              for (int k = 0; k < 2; ++k) umma_bf16(d2, a_hi + 2 * k, w_hi + 2 * k, idesc, (kg | g | k) != 0);
              if (NTERMS == 3) {
                const uint64_t a_lo = make_smem_desc<32>(sm + kRuMidChunkBytes);
                const uint64_t w_lo = make_smem_desc<32>(sw + (uint32_t)(g * Cfg::kPlanes + 1) * Cfg::kWChunk);
#pragma unroll
                for (int k = 0; k < 2; ++k) umma_bf16(d2, a_lo + 2 * k, w_hi + 2 * k, idesc, 1u);
#pragma unroll
                for (int k = 0; k < 2; ++k) umma_bf16(d2, a_hi + 2 * k, w_lo + 2 * k, idesc, 1u);
              }
              umma_commit(mid_empty(ms));
              if (++ms == SM) { ms = 0; mph ^= 1u; }
            }
            umma_commit(w_empty(ws));
            if (++ws == SW) { ws = 0; wph ^= 1u; }
          }
          umma_commit(acc2_full(jt % NB2));
        }
      }
    }
  } else if (warp == kRuResWarp) {
    // ================================ residual TMA producer ================================
    if (elect_one()) {
      prefetch_tmap(&tm_res);
      uint32_t rs = 0, rph = 0;
      for (int jt = 0; jt < n_my; ++jt) {
        const int tile = blockIdx.x + jt * gridDim.x;
        const int b = tile / p.m_tiles_per_utt;
        const int l0 = (tile % p.m_tiles_per_utt) * kBlockM;
        for (int c = 0; c < C; c += 32) {
          mbar_wait(res_empty(rs), rph ^ 1u);
          mbar_expect_tx(res_full(rs), kRuSlabBytes);
          tma_load_3d(res_base + rs * kRuSlabBytes, &tm_res, res_full(rs), c, l0, b);
          if (++rs == SR) { rs = 0; rph ^= 1u; }
        }
      }
    }
  } else {
    // ================================ epilogue warps ================================
    const int group = warp & 3;                 // TMEM lane quarter this warp may read
    const int row_in_tile = group * 32 + lane;
    const int ew = warp - 2;                    // 0..7
    const int half = ew >> 2;                   // which 16 of a chunk's 32 columns this warp drains from TMEM
    const int q4 = lane & 7, rsub = lane >> 3;  // final stage, phase 2: column quad, row within a 4-row group
    uint32_t ms = 0, mph = 0, rs = 0, rph = 0;
    for (int it = 0; it < n_my + SKEW; ++it) {
      if (it < n_my) {
        // ---- mid stage of tile `it`: acc1 -> + b7 -> Snake -> bf16 hi/lo -> K-major SWIZZLE_64B smem chunks
        mbar_wait(acc1_full(it % NB1), (uint32_t)(it / NB1) & 1u);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(group * 32) << 16) + (uint32_t)(it % NB1) * C;
#pragma unroll 1
        for (int c = 0; c < C; c += 32) {
          mbar_wait(mid_empty(ms), mph ^ 1u);
          uint32_t r[16];
          tmem_ld_x16(t_row + c + 16 * half, r);
          tmem_ld_wait();
          const int n0 = c + 16 * half;
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_par + n0 + 4 * q);
            const float4 a4 = *reinterpret_cast<const float4*>(s_par + C + n0 + 4 * q);
            const float4 i4 = *reinterpret_cast<const float4*>(s_par + 2 * C + n0 + 4 * q);
            const float v0 = snake_f(__uint_as_float(r[4 * q + 0]) + b4.x, a4.x, i4.x);
            const float v1 = snake_f(__uint_as_float(r[4 * q + 1]) + b4.y, a4.y, i4.y);
            const float v2 = snake_f(__uint_as_float(r[4 * q + 2]) + b4.z, a4.z, i4.z);
            const float v3 = snake_f(__uint_as_float(r[4 * q + 3]) + b4.w, a4.w, i4.w);
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
            hi[2 * q] = pack_bf16(h0);
            hi[2 * q + 1] = pack_bf16(h1);
            if (NTERMS == 3) {
              const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
              lo[2 * q] = pack_bf16(__floats2bfloat162_rn(v0 - f0.x, v1 - f0.y));
              lo[2 * q + 1] = pack_bf16(__floats2bfloat162_rn(v2 - f1.x, v3 - f1.y));
            }
          }
          // row r of a chunk is 64 B; 16 B piece j lives at j ^ ((r >> 1) & 3) (TMA/UMMA SWIZZLE_64B)
          const uint32_t row_addr = mid_base + ms * Cfg::kMidSlot + (uint32_t)row_in_tile * 64u;
          const uint32_t sw = (uint32_t)(row_in_tile >> 1) & 3u;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t off = (((uint32_t)(2 * half + j)) ^ sw) << 4;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + off), "r"(hi[4 * j]),
                         "r"(hi[4 * j + 1]), "r"(hi[4 * j + 2]), "r"(hi[4 * j + 3])
                         : "memory");
            if (NTERMS == 3)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + kRuMidChunkBytes + off),
                           "r"(lo[4 * j]), "r"(lo[4 * j + 1]), "r"(lo[4 * j + 2]), "r"(lo[4 * j + 3])
                           : "memory");
          }
          fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
          tc_fence_before();
          mbar_arrive(mid_full(ms));
          if (++ms == SM) { ms = 0; mph ^= 1u; }
        }
      }
      if (it >= SKEW) {
        // ---- final stage of tile jt: acc2 + residual slab (+ b1) -> x, Snake -> operand planes.
        // Phase 1: each thread adds its 16 accumulator columns into ITS row of the residual slab (in place);
        // phase 2: the slab is read back transposed (8 lanes = one 128 B row) so every global access is a
        // full row segment.
        const int jt = it - SKEW;
        const int tile = blockIdx.x + jt * gridDim.x;
        const int b = tile / p.m_tiles_per_utt;
        const int l0 = (tile % p.m_tiles_per_utt) * kBlockM;
        mbar_wait(acc2_full(jt % NB2), (uint32_t)(jt / NB2) & 1u);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(group * 32) << 16) + (uint32_t)(NB1 + jt % NB2) * C;
#pragma unroll 1
        for (int c = 0; c < C; c += 32) {
          const int n = c + q4 * 4;
          const float4 bias4 = __ldg(reinterpret_cast<const float4*>(p.bias1 + n));
          float4 alpha4 = make_float4(0.f, 0.f, 0.f, 0.f), inv4 = alpha4;
          if (p.out_hi) {
            alpha4 = __ldg(reinterpret_cast<const float4*>(p.alpha_out + n));
            inv4 = __ldg(reinterpret_cast<const float4*>(p.inv_out + n));
          }
          uint32_t r[16];
          tmem_ld_x16(t_row + c + 16 * half, r);
          tmem_ld_wait();
          if (c + 32 >= C) {   // accumulator fully drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            mbar_arrive(acc2_empty(jt % NB2));
          }
          const uint32_t slab = res_base + rs * kRuSlabBytes;
          mbar_wait(res_full(rs), rph);
          {
            const uint32_t row_addr = slab + (uint32_t)row_in_tile * 128u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t addr = row_addr + ((((uint32_t)(4 * half + j)) ^ ((uint32_t)row_in_tile & 7u)) << 4);
              float4 v;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                           : "r"(addr));
              // same association as the stand-alone 1x1 kernel: (acc + residual) + bias
              v.x = __uint_as_float(r[4 * j + 0]) + v.x;
              v.y = __uint_as_float(r[4 * j + 1]) + v.y;
              v.z = __uint_as_float(r[4 * j + 2]) + v.z;
              v.w = __uint_as_float(r[4 * j + 3]) + v.w;
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                           : "memory");
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kRuEpiThreads) : "memory");
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = ew * 16 + i * 4 + rsub;
            const int l = l0 + rr;
            const uint32_t off = (uint32_t)rr * 128u + ((uint32_t)(q4 ^ (rr & 7)) << 4);
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"(slab + off));
            if (l < p.L) {
              v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
              const size_t idx = ((size_t)b * p.L + l) * (size_t)C + n;
              *reinterpret_cast<float4*>(p.x + idx) = v;
              if (p.out_hi) {
                v.x = snake_f(v.x, alpha4.x, inv4.x); v.y = snake_f(v.y, alpha4.y, inv4.y);
                v.z = snake_f(v.z, alpha4.z, inv4.z); v.w = snake_f(v.w, alpha4.w, inv4.w);
                const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
                *reinterpret_cast<uint2*>(p.out_hi + idx) = make_uint2(pack_bf16(h0), pack_bf16(h1));
                if (p.out_lo) {
                  const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
                  *reinterpret_cast<uint2*>(p.out_lo + idx) =
                      make_uint2(pack_bf16(__floats2bfloat162_rn(v.x - f0.x, v.y - f0.y)),
                                 pack_bf16(__floats2bfloat162_rn(v.z - f1.x, v.w - f1.y)));
                }
              }
            }
          }
          fence_proxy_async();   // this slab was written through the generic proxy; the next TMA load overwrites it
          mbar_arrive(res_empty(rs));
          if (++rs == SR) { rs = 0; rph ^= 1u; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int C, int NTERMS>
int launch_ru(const GemmWeights& c7, const GemmWeights& c1, const OpBuf& a, int batch, int L, const RuParams& p,
              int num_sms, cudaStream_t stream) {
  using Cfg = RuCfg<C, NTERMS>;
  CUtensorMap ta_hi, ta_lo, t_res;
  const uint64_t dims[3] = {(uint64_t)C, (uint64_t)L, (uint64_t)batch};
  const uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)L * C * 2};
  const uint32_t box[3] = {32u, (uint32_t)p.halo_rows, 1u};
  SC_TRY(encode_tmap(&ta_hi, a.hi, 3, dims, strides, box, 64, false, false));
  if (NTERMS == 3) SC_TRY(encode_tmap(&ta_lo, a.lo, 3, dims, strides, box, 64, false, false));
  else ta_lo = ta_hi;
  const uint64_t rstr[2] = {(uint64_t)C * 4, (uint64_t)L * C * 4};
  const uint32_t rbox[3] = {32u, (uint32_t)kBlockM, 1u};
  SC_TRY(encode_tmap(&t_res, p.x, 3, dims, rstr, rbox, 128, false, true));
  auto kern = resunit_fused_kernel<C, NTERMS>;
  static bool attr_done = false;   // per instantiation
  if (!attr_done) {
    SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  kern<<<grid, kRuThreads, Cfg::kSmemBytes, stream>>>(ta_hi, ta_lo, c7.tmap_hi[1], NTERMS == 3 ? c7.tmap_lo[1] : c7.tmap_hi[1],
                                                       c1.tmap_hi[1], NTERMS == 3 ? c1.tmap_lo[1] : c1.tmap_hi[1], t_res, p);
  SC_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// k7 conv (single phase, 7 taps at multiples of one dilation <= 9) followed by a 1x1 conv, both C -> C with
// C in {96, 192}: the shapes of the last two WaveGenerator stages.
bool resunit_fusable(const GemmWeights& c7, const GemmWeights& c1, int* dil) {
  const int C = c7.c_in;
  if (C != 96 && C != 192) return false;
  if (c7.n_total != C || c1.c_in != C || c1.n_total != C) return false;
  if (c7.taps.n_phase != 1 || c7.taps.ntaps[0] != 7 || c1.taps.n_phase != 1 || c1.taps.ntaps[0] != 1) return false;
  if (c1.taps.shift[0][0] != 0 || c7.block_n != C || c1.block_n != C) return false;
  const int d = c7.taps.shift[0][4] - c7.taps.shift[0][3];
  if (d < 1 || kBlockM + 6 * d > kRuHaloRowsMax) return false;
  for (int j = 0; j < 7; ++j)
    if (c7.taps.shift[0][j] != (j - 3) * d) return false;
  if (dil) *dil = d;
  return true;
}

int launch_resunit_fused(const GemmWeights& c7, const GemmWeights& c1, const OpBuf& a, int batch, int L,
                         const float* alpha_mid, const float* inv_mid, float* x, const float* alpha_out,
                         const float* inv_out, OpBuf out, int precision, int num_sms, cudaStream_t stream) {
  int dil = 0;
  if (!resunit_fusable(c7, c1, &dil)) {
    set_error("resunit_fused: unsupported layer shapes (C=%d)", c7.c_in);
    return SPARKCODEC_EINVAL;
  }
  const bool f32 = precision == SPARKCODEC_PREC_FP32;
  if (f32 && (!a.lo || (out.hi && !out.lo))) {
    set_error("resunit_fused: fp32 mode needs both operand planes");
    return SPARKCODEC_EINVAL;
  }
  RuParams p;
  p.batch = batch; p.L = L; p.dil = dil;
  p.halo_rows = (kBlockM + 6 * dil + 7) / 8 * 8;
  p.m_tiles_per_utt = (L + kBlockM - 1) / kBlockM;
  p.num_tiles = batch * p.m_tiles_per_utt;
  p.bias7 = c7.bias; p.alpha_mid = alpha_mid; p.inv_mid = inv_mid;
  p.bias1 = c1.bias; p.alpha_out = alpha_out; p.inv_out = inv_out;
  p.x = x;
  p.out_hi = out.hi;
  p.out_lo = f32 ? out.lo : nullptr;
  if (c7.c_in == 96)
    return f32 ? launch_ru<96, 3>(c7, c1, a, batch, L, p, num_sms, stream)
               : launch_ru<96, 1>(c7, c1, a, batch, L, p, num_sms, stream);
  return f32 ? launch_ru<192, 3>(c7, c1, a, batch, L, p, num_sms, stream)
             : launch_ru<192, 1>(c7, c1, a, batch, L, p, num_sms, stream);
}

}  // namespace sparkcodec
