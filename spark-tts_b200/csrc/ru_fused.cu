// Fused ResidualUnit of the WaveGenerator (C = 96 / 192 / 384 channels), one kernel:
//
//   mid   = snake_mid( conv_k7_dil( op_in ) + b7 )          op_in = snake_in(x) as bf16 hi/lo planes
//   x     = x + conv_1x1( mid ) + b1                         fp32 residual stream, updated in place
//   op_out = snake_next(x)                                   operand planes of the next layer (optional)
//
// (reference: ResidualUnit.forward, sparktts/modules/blocks/layers.py:51-67; the Snake of the NEXT unit /
// up-sampler is applied here because it is element-wise on this unit's output.)
//
// Un-fused, the k7 conv writes `mid` to HBM and the 1x1 conv reads it back; the 1x1 conv is then purely HBM-bound
// (16 B per output element).  Here `mid` never leaves the SM:
//
//   TMEM acc1 (k7 result) --mid team: tcgen05.ld, +b7, Snake, bf16 hi/lo split, tcgen05.st IN PLACE--> the 32 fp32
//        columns of a chunk become 16 columns of packed bf16 hi + 16 columns of packed bf16 lo
//        --tcgen05.mma with the A operand in TENSOR MEMORY (B = W1 chunk from smem, split-K over chunks)--> acc2
//        --final team: + residual slab (TMA) + b1 -> x slab, Snake -> operand planes, all stored with TMA.
// Keeping `mid` in TMEM leaves the whole shared memory to the operand rings: the weight ring has to cover the
// ~2k-cycle refill round trip (MMA done -> commit -> producer -> TMA from L2 -> full), see DESIGN.md.
//
// Roles (640 threads = 20 warps, i.e. five per scheduler and 96 registers per thread -- a 21st warp costs every thread
// 16 registers; one persistent CTA per SM, clusters of two CTAs):
//   warps 0..7   : mid team   (acc1 -> mid operand, two chunks in flight)
//   warps 8..15  : final team (acc2 + residual -> x, operand planes)
//   warp 16 / one lane: output side.  Residual slabs in (TMA loads) and the final stage's TMA stores out: the teams
//                       only write a chunk into shared memory and arrive on an mbarrier; this thread issues the stores,
//                       hands the staging slot back once the TMA unit has read it and refills the residual slab it just
//                       stored from -- the store latency is off the teams' critical path and the slab ring needs no
//                       "empty" barriers
//   warp 17 / 18, one lane each: TMA producers (activation halo tiles, W7 / W1 weight ring)
//   warp 19 / one lane: tcgen05.mma issuer (highest warp id: the scheduler favours it over the polling teams)
// TMEM plans (512 columns):
//   C =  96: acc1 x2 + acc2 x2 (96 columns each); the 1x1 conv of tile i-1 is issued after the k7 conv of tile i (SKEW 1)
//   C = 192: acc1 x2 (384) + ONE 96-column acc2: the 1x1 conv runs as two N halves between the halves of the next k7 conv
//   C = 384: ONE acc1 (384, the k7 conv is issued as two N = 192 MMAs per step) + two 64-column acc2 buffers: the 1x1
//            conv runs as six N chunks that ping-pong between them (the final team drains chunk n while chunk n+1 is
//            accumulated); no skew -- the tensor pipe idles only while the first mid chunks are converted.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "gemm_params.cuh"
#include "tc_ptx.cuh"

namespace sparkcodec {
namespace {

constexpr int kRuHaloRowsMax = 184;                 // 128 + 6 * 9, multiple of 8
constexpr int kRuAChunkBytes = 12 * 1024;           // 184 rows x 64 B, rounded up to 1024
constexpr int kRuSlabBytes = kBlockM * 128;         // 128 rows x 32 fp32
constexpr int kRuPlaneTile = kBlockM * 64;          // 128 rows x 32 bf16 (one operand plane of an output chunk)
// warp roles: 0..7 mid team, 8..15 final team, 16 residual loads + output stores, 17 / 18 TMA producers (activations,
// weights), 19 MMA issuer.
// The warp scheduler favours the highest warp id among eligible warps, so the single-thread roles that
// feed the tensor pipe sit ABOVE the 16 epilogue warps (which spend much of their time polling mbarriers).
// A warp may only read the TMEM lane quarter warp_id % 4: both teams start at a multiple of 4, so
// `warp & 3` enumerates the four quarters twice per team.
constexpr int kRuTeamWarps = 8;
constexpr int kRuTeamThreads = kRuTeamWarps * 32;
constexpr int kRuMidWarp0 = 0;
constexpr int kRuFinWarp0 = kRuMidWarp0 + kRuTeamWarps;
constexpr int kRuResWarp = kRuFinWarp0 + kRuTeamWarps;   // residual slabs in, x slabs + operand planes out
constexpr int kRuTmaAWarp = kRuResWarp + 1;              // activation halo tiles
constexpr int kRuTmaWWarp = kRuTmaAWarp + 1;             // W7 / W1 tiles
constexpr int kRuMmaWarp = kRuTmaWWarp + 1;
constexpr int kRuThreads = (kRuMmaWarp + 1) * 32;

struct RuParams {
  int batch, L, dil, halo_rows;
  int m_tiles_per_utt, num_tiles;
  const float *bias7, *alpha_mid, *inv_mid;   // [C]
  const float *bias1, *alpha_out, *inv_out;   // [C]; alpha_out/inv_out unused when out_hi == null
  float* x;                                   // (batch, L, C) fp32, read as residual and overwritten
  __nv_bfloat16 *out_hi, *out_lo;             // (batch, L, C) operand planes of the next layer (optional)
  long long* dbg;                             // SPARKCODEC_RU_TRACE: per-tile event clocks of CTA 0 (else null)
  int l2_prefetch;                            // issue L2 prefetches of the next tile's operand / residual rows
};
constexpr int kRuTraceEvents = 32, kRuTraceTiles = 16;

template <int C, int NTERMS, int PG>   // PG = CTAs sharing one MMA (1, or 2 = cta_group::2 pair)
struct RuCfg {
  static constexpr bool WIDE = C > 256;                            // C = 384: one acc1, k7 conv as two N halves
  static constexpr int kPlanes = NTERMS >= 2 ? 2 : 1;
  static constexpr int kChunks = C / 32;                           // K chunks of 32 channels
  static constexpr int G = (NTERMS >= 2 || WIDE) ? 1 : (C == 96 ? 3 : 2);   // K chunks per ring stage (k7 conv)
  static constexpr int kGroups = kChunks / G;
  static constexpr int kAStage = G * kPlanes * kRuAChunkBytes;
  static constexpr int KN = WIDE ? 2 : 1;                          // N halves of the k7 conv (UMMA N <= 256)
  static constexpr int N7 = C / KN;                                // N of one k7 MMA
  static constexpr int kWChunk = (C / PG) * 64;                    // C / PG rows x 64 B (one plane): a CTA pair splits the N rows
  // Taps of the k7 conv per weight-ring stage.  C = 96 in fp32 mode: a (tap, K chunk) stage is only 6 N = 96 MMAs
  // (336 tensor-pipe cycles), less than what the single issuing thread spends per stage on the barrier wait, the
  // descriptor arithmetic and the commit -- the kernel was ISSUE bound at ~67 cycles per 56-cycle MMA.  Four taps
  // per stage (groups of 4 + 3) amortise that 3.5x.
  static constexpr int TJ = (C == 96 && NTERMS >= 2 && PG == 2) ? 4 : 1;   // (pair mode: 3 KB per plane and tap)
  static constexpr int kWTap = G * kPlanes * kWChunk;              // bytes of one tap inside a stage
  static constexpr int kWStage = TJ * kWTap;
  // TMEM plan.  C <= 192: acc1 (k7 result, then the converted mid operand) is double buffered and the 1x1 conv of
  // tile i - 1 is issued AFTER (part of) the k7 conv of tile i (SKEW = 1), so the tensor pipe never waits for the
  // mid stage.  C = 96: two 96-column acc2 buffers (4 x 96 = 384 columns).  C = 192: 2 x 192 columns of acc1 leave
  // 128, so the 1x1 conv runs as TWO N halves through ONE 96-column accumulator; the halves are issued between the
  // two halves of the next tile's k7 conv, which gives the final team time to drain the first one.  C = 384: acc1
  // alone takes 384 columns (single buffer, SKEW = 0); the 1x1 conv runs as SIX 64-column N chunks through TWO
  // accumulators, so the final team drains chunk n while the tensor cores accumulate chunk n + 1.
  static constexpr bool SPLIT = C > 96;
  static constexpr int NB1 = WIDE ? 1 : 2;
  static constexpr int NH = WIDE ? 6 : (SPLIT ? 2 : 1);            // N chunks of the 1x1 conv
  static constexpr int N2 = C / NH;                                // accumulator width of the 1x1 conv
  static constexpr int NB2 = WIDE ? 2 : (SPLIT ? 1 : 2);
  static constexpr int SKEW = WIDE ? 0 : 1;
  // W1 ring stages.  C <= 192: one stage per K group, chunk stride = the W7 chunk stride.  C = 384: a chunk is only
  // N2 / PG = 32 or 64 rows, so a stage carries G1 = 6 K chunks (compact) -- 12 stages per tile instead of 72.
  // (C = 96 with four-tap stages: the whole W1 -- 3 chunks, 18 KB -- is ONE stage; as three stages it occupied the entire
  // 3-slot ring during the 1x1 phase and the next tile's first W7 stage could not be prefetched.)
  static constexpr int G1 = WIDE ? 6 : (TJ > 1 ? kChunks : G);
  static constexpr int kGroups1 = kChunks / G1;
  static constexpr int kW1Chunk = WIDE ? (N2 / PG) * 64 : kWChunk;                 // chunk stride (one plane) inside a W1 stage
  static constexpr int kW1Stage = G1 * kPlanes * (N2 / PG) * 64;                    // bytes of a W1 stage in this CTA
  // ring depths (227 KB budget)
  static constexpr int SA = (NTERMS >= 2) ? 2 : (C <= 96 ? 2 : 3);
  // Output side of the final stage.  Every 32-column chunk ends with TMA stores (x slab + operand planes) issued by
  // one thread; the team may run kOutSlots - 2 chunks ahead of the stores' shared-memory reads.  C = 96 has
  // shared memory to spare, so it gets three staging slots and a fourth residual slab (worth 1-2 %: its limit is
  // the tensor pipe itself, whose N = 96 MMAs take 56 instead of 48 cycles reading operands from shared memory,
  // plus ~20 % weight-ring waits).  C = 192 keeps the shared memory for the weight ring.  C = 384: BOTH epilogue
  // teams run the final stage on alternate 32-column chunks (the 1x1 phase is not overlapped with a k7 conv there,
  // so the drain rate is what the tensor pipe waits for): one staging slot and two residual slabs per team.
  // (C = 96 with four-tap weight stages: two staging slots, so that FOUR residual slabs still fit -- the slab of a
  // chunk comes from HBM and has to be requested two chunks ahead, a late slab cost the final team 1.6k cycles a tile.)
  static constexpr int kOutSlots = (C <= 96 && TJ == 1) ? 3 : 2;
  static constexpr int SR = (C <= 96 || WIDE) ? 4 : 3;   // WIDE: two slabs per epilogue team (slot parity = team)
  static constexpr int kParBytes = 6 * C * 4;   // bias7 | alpha_mid | inv_mid | bias1 | alpha_out | inv_out
  static constexpr int kStageOut = kOutSlots * kPlanes * kRuPlaneTile;   // staging of the operand-plane TMA stores
  static constexpr int kFixed = SA * kAStage + SR * kRuSlabBytes + kStageOut + kParBytes + 1024 /* barriers */ +
                                1024 /* alignment */;
  static constexpr int SWRaw = (227 * 1024 - kFixed) / kWStage;
  static constexpr int SW = SWRaw > 12 ? 12 : SWRaw;
  static constexpr int kNumMid = NB1 * kChunks;                    // one "chunk converted" barrier per (acc1 buffer, chunk)
  static constexpr int kNumBars = 2 * SA + 2 * SW + kNumMid + NB1 + 2 * NB2 + SR + 2 * kOutSlots;
  static constexpr int kSmemBytes = SA * kAStage + SW * kWStage + SR * kRuSlabBytes + kStageOut + kParBytes +
                                    kNumBars * 8 + 16 + 1024 /* alignment */;
  static_assert(kNumBars * 8 + 16 <= 1024, "barrier block larger than budgeted");
  static_assert(SW >= 3, "weight ring too shallow");
  static_assert(kChunks % G == 0 && kChunks % G1 == 0, "K groups must tile the channels");
  static_assert(kW1Stage <= kWStage, "a W1 stage must fit a ring slot");
  static_assert(NB1 * C + NB2 * N2 <= 512, "accumulators must fit TMEM");
  static_assert(N7 <= 256 && N7 % 16 == 0 && N2 % 16 == 0, "UMMA N range");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
  static_assert(kWChunk % 1024 == 0 && kW1Chunk % 1024 == 0 && ((N7 / PG) * 64) % 1024 == 0,
                "weight chunks must stay 1024 B aligned");
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }
// event trace of CTA 0 (debug builds of the schedule; dbg is null in normal runs)
__device__ __forceinline__ void ru_trace(const RuParams& p, int tile_it, int ev) {
  if (p.dbg && blockIdx.x == 0 && tile_it < kRuTraceTiles && ev < kRuTraceEvents - 5)
    p.dbg[tile_it * kRuTraceEvents + ev] = clock64();
}

// mbar_wait that adds the cycles spent waiting to `acc` when tracing (schedule debugging)
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool timed, long long& acc) {
  if (!timed) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}

template <int C, int NTERMS, int CL, bool PAIR>
__global__ void __launch_bounds__(kRuThreads, 1)
resunit_fused_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                     const __grid_constant__ CUtensorMap tm_w7_hi, const __grid_constant__ CUtensorMap tm_w7_lo,
                     const __grid_constant__ CUtensorMap tm_w1_hi, const __grid_constant__ CUtensorMap tm_w1_lo,
                     const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_o_hi,
                     const __grid_constant__ CUtensorMap tm_o_lo, const RuParams p) {
  static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of two");
  constexpr int PG = PAIR ? 2 : 1;
  constexpr bool EXACT = NTERMS == 3;   // range-reduced sine only where the products are accurate enough to see it
  using Cfg = RuCfg<C, NTERMS, PG>;
  constexpr int SA = Cfg::SA, SW = Cfg::SW, SR = Cfg::SR, NB1 = Cfg::NB1, NB2 = Cfg::NB2;
  constexpr int G = Cfg::G, SKEW = Cfg::SKEW, NMID = Cfg::kNumMid, NH = Cfg::NH, N2 = Cfg::N2;
  constexpr int KN = Cfg::KN, N7 = Cfg::N7, G1 = Cfg::G1, TJ = Cfg::TJ;
  constexpr int RB7 = N7 / CL;                         // W7 rows one CTA fetches per N half and stage (= TMA box rows)
  // k7 K groups issued before the first 1x1 half (skewed C = 192 schedule only)
  constexpr int KG1 = (Cfg::SPLIT && SKEW == 1) ? (Cfg::kGroups + 1) / 2 : Cfg::kGroups;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t w_base = a_base + SA * Cfg::kAStage;
  const uint32_t res_base = w_base + SW * Cfg::kWStage;
  const uint32_t out_base = res_base + SR * kRuSlabBytes;          // operand-plane staging, 1024 B aligned
  const uint32_t par_base = out_base + Cfg::kStageOut;
  const uint32_t bar_base = par_base + Cfg::kParBytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (SA + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * SA + SW + s); };
  auto mid_full = [&](int buf, int chunk) { return bar_base + 8u * (2 * SA + 2 * SW + buf * Cfg::kChunks + chunk); };
  auto acc1_full = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + NMID + s); };
  auto acc2_full = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + NMID + NB1 + s); };
  auto acc2_empty = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + NMID + NB1 + NB2 + s); };
  auto res_full = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + NMID + NB1 + 2 * NB2 + s); };
  // staging slot written by the team -> output thread / read by the TMA unit -> team
  auto out_full = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + NMID + NB1 + 2 * NB2 + SR + s); };
  auto out_empty = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SW + NMID + NB1 + 2 * NB2 + SR + Cfg::kOutSlots + s); };
  const uint32_t tmem_slot = bar_base + 8u * Cfg::kNumBars;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* s_par = reinterpret_cast<float*>(smem_raw + (par_base - smem_u32(smem_raw)));   // [bias7 | alpha_mid | inv_mid]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Tiles are dealt to CLUSTERS of CL CTAs that work on CL consecutive tiles in lock step and share the weights:
  //  * PAIR (cta_group::2): the two CTAs run ONE M = 256 MMA stream issued by the leader; each CTA stages its own
  //    128 activation rows and HALF of the weight rows of every stage (the tensor cores read the other half from
  //    the peer), so the weight traffic per SM halves and the weight ring is twice as deep in tensor-pipe time.
  //    Best where the tensor pipe is the limit (fp32 mode: 3 MMAs per MAC).
  //  * multicast (CL == 2, !PAIR): independent M = 128 MMA streams; every weight stage is fetched half by each CTA
  //    and multicast to both (half the L2 reads, same shared-memory footprint).  Best where the epilogue is the
  //    limit (bf16 mode): the CTAs are only coupled through the weight ring.
  // A trailing tile without a partner is a dummy (all rows out of range: TMA zero fill, stores masked).
  const int cl_rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const bool leader = !PAIR || cl_rank == 0;       // owner of the barriers the MMA issuer waits on
  // barrier of the pair's leader CTA / arrive on it from either CTA
  auto lead = [&](uint32_t bar) { return PAIR ? mapa_shared(bar, 0) : bar; };
  auto arrive_lead = [&](uint32_t bar) {
    if (PAIR) mbar_arrive_cluster_relaxed(mapa_shared(bar, 0));   // tensor-memory hand-offs only (mid_full, acc2_empty)
    else mbar_arrive(bar);
  };
  const int cid = (int)blockIdx.x / CL, ncl = (int)gridDim.x / CL;
  const int n_groups = (p.num_tiles + CL - 1) / CL;
  const int n_my = cid < n_groups ? (n_groups - cid + ncl - 1) / ncl : 0;
  auto tile_of = [&](int it) { return (cid + it * ncl) * CL + cl_rank; };
  auto tile_b = [&](int tile) { return tile < p.num_tiles ? tile / p.m_tiles_per_utt : p.batch; };   // batch = out of range
  auto tile_l0 = [&](int tile) { return tile < p.num_tiles ? (tile % p.m_tiles_per_utt) * kBlockM : 0; };

  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_par[i] = __ldg(p.bias7 + i);
    s_par[C + i] = __ldg(p.alpha_mid + i);
    s_par[2 * C + i] = __ldg(p.inv_mid + i);
    // the final stage is latency bound (one chunk in flight per team): no global loads inside a chunk
    s_par[3 * C + i] = __ldg(p.bias1 + i);
    s_par[4 * C + i] = p.out_hi ? __ldg(p.alpha_out + i) : 0.f;
    s_par[5 * C + i] = p.out_hi ? __ldg(p.inv_out + i) : 0.f;
  }
  if (warp == kRuTmaWWarp && lane == 0) {
    prefetch_tmap(&tm_a_hi);
    prefetch_tmap(&tm_w7_hi);
    prefetch_tmap(&tm_w1_hi);
    if (NTERMS >= 2) {
      prefetch_tmap(&tm_a_lo);
      prefetch_tmap(&tm_w7_lo);
      prefetch_tmap(&tm_w1_lo);
    }
    for (int s = 0; s < SA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < SW; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), PAIR ? 1 : CL); }
    for (int s = 0; s < NMID; ++s) mbar_init(mid_full(0, s), PG);   // one arrive per CTA (after a named barrier of the 4 warps that converted the chunk)
    for (int s = 0; s < NB1; ++s) mbar_init(acc1_full(s), 1);
    // acc2 is handed back by one thread per final-stage team and CTA (WIDE: both teams drain every accumulator)
    for (int s = 0; s < NB2; ++s) { mbar_init(acc2_full(s), 1); mbar_init(acc2_empty(s), PG * (Cfg::WIDE ? 2 : 1)); }
    for (int s = 0; s < SR; ++s) mbar_init(res_full(s), 1);
    for (int s = 0; s < Cfg::kOutSlots; ++s) { mbar_init(out_full(s), 1); mbar_init(out_empty(s), 1); }
    fence_barrier_init();
  }
  if (warp == kRuMmaWarp) {
    if (PAIR) tmem_alloc_cg2<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();   // peers' barriers are initialised before any multicast / remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // PDL (common.cuh): everything above overlapped the previous kernel's tail; from here on its outputs are read
  griddep_wait();
  griddep_launch_dependents();

  if (warp == kRuTmaAWarp) {
    // ================================ TMA producer: activation halo tiles ================================
    if (elect_one()) {
      uint32_t as = 0, aph = 0;
      const uint32_t a_tx = (uint32_t)(G * Cfg::kPlanes) * (uint32_t)p.halo_rows * 64u;
      for (int it = 0; it < n_my; ++it) {
        const int tile = tile_of(it);
        const int b = tile_b(tile);
        const int row0 = tile_l0(tile) - 3 * p.dil;
        if (p.l2_prefetch && it + 1 < n_my) {   // next tile's operand rows -> L2 (they come from HBM: written by the previous kernel)
          const int nt = tile_of(it + 1);
          const int nb = tile_b(nt), nrow0 = tile_l0(nt) - 3 * p.dil;
          for (int kc = 0; kc < Cfg::kChunks; ++kc) {
            tma_prefetch_3d(&tm_a_hi, kc * 32, nrow0, nb);
            if (NTERMS >= 2) tma_prefetch_3d(&tm_a_lo, kc * 32, nrow0, nb);
          }
        }
        for (int kg = 0; kg < Cfg::kGroups; ++kg) {
          mbar_wait(a_empty(as), aph ^ 1u);
          if (leader) mbar_expect_tx(a_full(as), PG * a_tx);   // pair: bytes of both CTAs land on the leader's barrier
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const uint32_t sa = a_base + as * Cfg::kAStage + (uint32_t)(g * Cfg::kPlanes) * kRuAChunkBytes;
            if (PAIR) {
              tma_load_3d_cg2(sa, &tm_a_hi, lead(a_full(as)), (kg * G + g) * 32, row0, b);
              if (NTERMS >= 2) tma_load_3d_cg2(sa + kRuAChunkBytes, &tm_a_lo, lead(a_full(as)), (kg * G + g) * 32, row0, b);
            } else {
              tma_load_3d(sa, &tm_a_hi, a_full(as), (kg * G + g) * 32, row0, b);
              if (NTERMS >= 2) tma_load_3d(sa + kRuAChunkBytes, &tm_a_lo, a_full(as), (kg * G + g) * 32, row0, b);
            }
          }
          if (++as == SA) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == kRuTmaWWarp) {
    // ================================ TMA producer: weights ================================
    // One ring for the W7 tap tiles of tile `it` and the W1 tiles of tile `it - SKEW`, in the order the MMA
    // warp consumes them.  Pair mode: this CTA fetches rows [cl_rank * C / 2, +C / 2) of every chunk into its
    // OWN shared memory and completes the bytes on the leader's barrier.  Multicast mode: it fetches the same
    // rows but multicasts them into BOTH CTAs' full-height chunks.
    if (elect_one()) {
      uint32_t ws = 0, wph = 0;
      constexpr uint16_t mask = (uint16_t)((1u << CL) - 1u);
      // W7 chunk of one stage: KN N halves of N7 rows.  Pair: this CTA holds rows [h N7 + rank RB7, +RB7) of half h at
      // chunk offset h RB7 rows (the tensor cores read the other RB7 rows of the half from the peer).  Multicast:
      // it fetches the same rows but writes them at their absolute row offset into BOTH CTAs' full-height chunks.
      auto load_w = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int k0) {
#pragma unroll
        for (int h = 0; h < KN; ++h) {
          const int row = h * N7 + cl_rank * RB7;
          if (PAIR) tma_load_2d_cg2(dst + (uint32_t)(h * RB7) * 64u, map, lead(bar), k0, row);
          else if (CL > 1) tma_load_2d_mc(dst + (uint32_t)row * 64u, map, bar, k0, row, mask);
          else tma_load_2d(dst + (uint32_t)row * 64u, map, bar, k0, row);
        }
      };
      const uint32_t mc1_off = (uint32_t)(cl_rank * (N2 / CL)) * 64u;
      auto load_w1 = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int k0, int half) {
        const int row = half * N2 + cl_rank * (N2 / CL);
        if (PAIR) tma_load_2d_cg2(dst, map, lead(bar), k0, row);
        else if (CL > 1) tma_load_2d_mc(dst + mc1_off, map, bar, k0, row, mask);
        else tma_load_2d(dst, map, bar, k0, half * N2);
      };
      auto w7_groups = [&](int g0, int g1) {
        for (int kg = g0; kg < g1; ++kg) {
          for (int j0 = 0; j0 < 7; j0 += TJ) {            // one stage = up to TJ consecutive taps of this K group
            const int nt = 7 - j0 < TJ ? 7 - j0 : TJ;
            mbar_wait(w_empty(ws), wph ^ 1u);
            if (leader) mbar_expect_tx(w_full(ws), (uint32_t)(PG * nt) * Cfg::kWTap);
            for (int t = 0; t < nt; ++t) {
#pragma unroll
              for (int g = 0; g < G; ++g) {
                const uint32_t sw = w_base + ws * Cfg::kWStage + (uint32_t)t * Cfg::kWTap + (uint32_t)(g * Cfg::kPlanes) * Cfg::kWChunk;
                load_w(sw, &tm_w7_hi, w_full(ws), (j0 + t) * C + (kg * G + g) * 32);
                if (NTERMS >= 2) load_w(sw + Cfg::kWChunk, &tm_w7_lo, w_full(ws), (j0 + t) * C + (kg * G + g) * 32);
              }
            }
            if (++ws == SW) { ws = 0; wph ^= 1u; }
          }
        }
      };
      auto w1_half = [&](int half) {
        for (int kg = 0; kg < Cfg::kGroups1; ++kg) {
          mbar_wait(w_empty(ws), wph ^ 1u);
          if (leader) mbar_expect_tx(w_full(ws), PG * Cfg::kW1Stage);
#pragma unroll
          for (int g = 0; g < G1; ++g) {
            const uint32_t sw = w_base + ws * Cfg::kWStage + (uint32_t)(g * Cfg::kPlanes) * Cfg::kW1Chunk;
            load_w1(sw, &tm_w1_hi, w_full(ws), (kg * G1 + g) * 32, half);
            if (NTERMS >= 2) load_w1(sw + Cfg::kW1Chunk, &tm_w1_lo, w_full(ws), (kg * G1 + g) * 32, half);
          }
          if (++ws == SW) { ws = 0; wph ^= 1u; }
        }
      };
      // same order as the MMA issuer.  Skewed (C <= 192): k7 groups [0, KG1) of tile it, 1x1 half 0 of tile it - 1,
      // the remaining k7 groups, 1x1 half 1 (C = 192 only).  Un-skewed (C = 384): k7 conv of tile it, then its NH
      // 1x1 chunks.
      for (int it = 0; it < n_my + SKEW; ++it) {
        if (it < n_my) w7_groups(0, KG1);
        if (SKEW == 0) {
          for (int nh = 0; nh < NH; ++nh) w1_half(nh);
          continue;
        }
        if (it >= SKEW) w1_half(0);
        if (it < n_my) w7_groups(KG1, Cfg::kGroups);
        if (NH > 1 && it >= SKEW) w1_half(1);
      }
    }
  } else if (warp == kRuMmaWarp) {
    // ================================ MMA issuer ================================
    // pair mode: the leader's thread issues the M = 256 MMAs for both CTAs
    if (leader && elect_one()) {
      constexpr int HF = NTERMS == 2 ? 0 : 1;   // kind::f16 operand format: fp16 (two-term fp32 mode) or bf16
      constexpr uint32_t idesc = PAIR ? make_idesc_cg2<N7, HF>() : make_idesc<N7, HF>();   // one k7 MMA covers N7 columns
      // cross terms of the two-term mode as e5m2 products (kind::f8f6f4, K = 32 = the 32 B step of a kind::f16 K = 16)
      constexpr uint32_t idesc8 = PAIR ? make_idesc_cg2<N7, 1>() : make_idesc<N7, 1>();
      auto mma8 = [&](uint32_t d, uint64_t a, uint64_t b) {
        if (PAIR) umma_f8_cg2(d, a, b, idesc8, 1u);
        else umma_f8(d, a, b, idesc8, 1u);
      };
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t accumulate) {
        if (PAIR) umma_bf16_cg2(d, a, b, idesc, accumulate);
        else umma_bf16(d, a, b, idesc, accumulate);
      };
      auto commit = [&](uint32_t bar) {     // pair mode: arrives on the barrier at this offset in BOTH CTAs
        if (PAIR) umma_commit_cg2(bar);
        else umma_commit(bar);
      };
      auto commit_w = [&](uint32_t bar) {   // weight slots are shared by the cluster in pair AND multicast mode
        if (PAIR) umma_commit_cg2(bar);
        else umma_commit_cl<CL>(bar);
      };
      constexpr uint32_t idesc2 = PAIR ? make_idesc_cg2<N2, HF>() : make_idesc<N2, HF>();   // the 1x1 conv is N2 wide
      constexpr uint32_t idesc82 = PAIR ? make_idesc_cg2<N2, 1>() : make_idesc<N2, 1>();
      auto mma8_ts = [&](uint32_t d, uint32_t a_tmem, uint64_t b) {
        if (PAIR) umma_f8_ts_cg2(d, a_tmem, b, idesc82, 1u);
        else umma_f8_ts(d, a_tmem, b, idesc82, 1u);
      };
      auto mma_ts = [&](uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t accumulate) {
        if (PAIR) umma_bf16_ts_cg2(d, a_tmem, b, idesc2, accumulate);
        else umma_bf16_ts(d, a_tmem, b, idesc2, accumulate);
      };
      uint32_t as = 0, aph = 0, ws = 0, wph = 0;
      const bool timed = p.dbg != nullptr && blockIdx.x == 0;
      long long wa = 0, ww = 0, ww1 = 0, wm = 0, w2 = 0;
      // ---- K groups [g0, g1) of the k7 (dilated) conv of tile `it` -> acc1[it % NB1].  The buffer is free: the 1x1
      // conv of the tile that last used it was issued earlier in program order (the tensor pipe is in order) and
      // had waited for every mid chunk, i.e. for the mid team to be done with it.
      auto k7_groups = [&](int it, int g0, int g1) {
        const uint32_t d1 = tmem_base + (uint32_t)(it % NB1) * C;
        for (int kg = g0; kg < g1; ++kg) {
          mbar_wait_t(a_full(as), aph, timed, wa);
          if (kg == 0) ru_trace(p, it, 0);
          const uint32_t sa = a_base + as * Cfg::kAStage;
          for (int j = 0; j < 7; ++j) {
            if (j % TJ == 0) {                                     // a ring stage carries TJ consecutive taps
              mbar_wait_t(w_full(ws), wph, timed, ww);
              tc_fence_after();
            }
            const uint32_t sw = w_base + ws * Cfg::kWStage + (uint32_t)(j % TJ) * Cfg::kWTap;
            const uint32_t a_off = (uint32_t)(j * p.dil) * 64u;   // tap j starts j*dil rows into the halo tile
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const uint64_t a_hi = make_smem_desc<32>(sa + (uint32_t)(g * Cfg::kPlanes) * kRuAChunkBytes + a_off);
              const uint64_t a_lo = make_smem_desc<32>(sa + (uint32_t)(g * Cfg::kPlanes + 1) * kRuAChunkBytes + a_off);
#pragma unroll
              for (int h = 0; h < KN; ++h) {   // N halves of the k7 conv: B rows [h N7, +N7) -> accumulator columns [h N7, +N7)
                // rows of half h inside a chunk: pair h * N7 / 2 (this CTA's share), otherwise h * N7
                const uint32_t hoff = (uint32_t)(h * (N7 / PG)) * 64u;
                const uint32_t dh = d1 + (uint32_t)(h * N7);
                const uint64_t w_hi = make_smem_desc<32>(sw + (uint32_t)(g * Cfg::kPlanes) * Cfg::kWChunk + hoff);
#pragma unroll
                for (int k = 0; k < 2; ++k) mma(dh, a_hi + 2 * k, w_hi + 2 * k, (kg | j | g | k) != 0);
                if (NTERMS >= 2) {
                  const uint64_t w_lo = make_smem_desc<32>(sw + (uint32_t)(g * Cfg::kPlanes + 1) * Cfg::kWChunk + hoff);
                  if (NTERMS == 2) {   // [A_lo8 | A_hi8] x [W_hi8 | W_lo8], 32 bytes each
#pragma unroll
                    for (int k = 0; k < 2; ++k) mma8(dh, a_lo + 2 * k, w_lo + 2 * k);
                  } else {
#pragma unroll
                    for (int k = 0; k < 2; ++k) mma(dh, a_lo + 2 * k, w_hi + 2 * k, 1u);
#pragma unroll
                    for (int k = 0; k < 2; ++k) mma(dh, a_hi + 2 * k, w_lo + 2 * k, 1u);
                  }
                }
              }
            }
            if (j % TJ == TJ - 1 || j == 6) {
              commit_w(w_empty(ws));
              if (++ws == SW) { ws = 0; wph ^= 1u; }
            }
          }
          commit(a_empty(as));
          if (++as == SA) { as = 0; aph ^= 1u; }
        }
        if (g1 == Cfg::kGroups) {
          commit(acc1_full(it % NB1));
          ru_trace(p, it, 1);
          if (timed && it < kRuTraceTiles) { p.dbg[it * kRuTraceEvents + 27] = wa; p.dbg[it * kRuTraceEvents + 28] = ww; }
          wa = ww = 0;
        }
      };
      // ---- N half `half` of the 1x1 conv of tile jt: A = the mid chunks in tensor memory (split-K as the mid team
      // publishes them), B = W1 rows [half * N2, +N2), D = acc2 buffer u % NB2 where u counts halves
      auto one_by_one = [&](int jt, int half) {
        const uint32_t u = (uint32_t)jt * NH + (uint32_t)half;
        const uint32_t d2 = tmem_base + (uint32_t)(NB1 * C) + (u % NB2) * N2;
        mbar_wait_t(acc2_empty(u % NB2), ((u / NB2) & 1u) ^ 1u, timed, w2);
        tc_fence_after();
        if (half == 0) ru_trace(p, jt, 2);
        const uint32_t mid_tmem = tmem_base + (uint32_t)(jt % NB1) * C;   // acc1 buffer of tile jt, converted in place
        const uint32_t mid_par = (uint32_t)(jt / NB1) & 1u;
        for (int kg = 0; kg < Cfg::kGroups1; ++kg) {
          mbar_wait_t(w_full(ws), wph, timed, ww1);
          const uint32_t sw = w_base + ws * Cfg::kWStage;
#pragma unroll
          for (int g = 0; g < G1; ++g) {
            const int kc = kg * G1 + g;
            if (half == 0) {
              mbar_wait_t(mid_full(jt % NB1, kc), mid_par, timed, wm);
              if (kc < 6) ru_trace(p, jt, 3 + kc);
            }
            tc_fence_after();
            // A operand in tensor memory: lane = row, 16 K values = 8 columns of packed bf16 pairs
            const uint32_t a_hi = mid_tmem + (uint32_t)kc * 32u;
            const uint64_t w_hi = make_smem_desc<32>(sw + (uint32_t)(g * Cfg::kPlanes) * Cfg::kW1Chunk);
#pragma unroll
            for (int k = 0; k < 2; ++k) mma_ts(d2, a_hi + 8 * k, w_hi + 2 * k, (kg | g | k) != 0);
            if (NTERMS >= 2) {
              const uint32_t a_lo = a_hi + 16u;
              const uint64_t w_lo = make_smem_desc<32>(sw + (uint32_t)(g * Cfg::kPlanes + 1) * Cfg::kW1Chunk);
              if (NTERMS == 2) {   // columns +16..23: 32 e5m2 lo values, +24..31: 32 e5m2 hi values (four per column)
#pragma unroll
                for (int k = 0; k < 2; ++k) mma8_ts(d2, a_lo + 8 * k, w_lo + 2 * k);
              } else {
#pragma unroll
                for (int k = 0; k < 2; ++k) mma_ts(d2, a_lo + 8 * k, w_hi + 2 * k, 1u);
#pragma unroll
                for (int k = 0; k < 2; ++k) mma_ts(d2, a_hi + 8 * k, w_lo + 2 * k, 1u);
              }
            }
          }
          commit_w(w_empty(ws));
          if (++ws == SW) { ws = 0; wph ^= 1u; }
        }
        commit(acc2_full(u % NB2));
        if (half == NH - 1) {
          ru_trace(p, jt, 9);
          if (timed && jt < kRuTraceTiles) {
            p.dbg[jt * kRuTraceEvents + 29] = ww1; p.dbg[jt * kRuTraceEvents + 30] = wm; p.dbg[jt * kRuTraceEvents + 31] = w2;
          }
          ww1 = wm = w2 = 0;
        }
      };
      for (int it = 0; it < n_my + SKEW; ++it) {
        if (it < n_my) k7_groups(it, 0, KG1);
        if (SKEW == 0) {   // C = 384: the tile's own 1x1 chunks follow its k7 conv (single acc1 buffer)
          for (int nh = 0; nh < NH; ++nh) one_by_one(it, nh);
          continue;
        }
        if (it >= SKEW) one_by_one(it - SKEW, 0);
        if (it < n_my && KG1 < Cfg::kGroups) k7_groups(it, KG1, Cfg::kGroups);
        if (NH > 1 && it >= SKEW) one_by_one(it - SKEW, 1);
      }
    }
  } else if (warp == kRuResWarp) {
    // ================================ output thread: residual slabs in, TMA stores out ================================
    // Walks the final stage's chunks in the teams' order (q = running chunk number of this CTA; chunk q uses residual
    // slab q % SR).  Per chunk: wait until all 256 threads of the team have written the slab + staging slot (out_full),
    // issue the stores as one bulk group, wait until the TMA unit has READ them, hand the staging slot back
    // (out_empty) and refill the slab with the residual tile of chunk q + SR.  The slab ring therefore needs no
    // "empty" barriers: the thread that frees a slab is the one that refills it.
    if (elect_one()) {
      prefetch_tmap(&tm_res);
      const bool has_out = p.out_hi != nullptr;
      const uint32_t total = (uint32_t)n_my * Cfg::kChunks;
      auto load_res = [&](uint32_t q) {
        const int tile = tile_of((int)(q / Cfg::kChunks));
        const uint32_t rs = q % SR;
        mbar_expect_tx(res_full(rs), kRuSlabBytes);
        tma_load_3d(res_base + rs * kRuSlabBytes, &tm_res, res_full(rs), (int)(q % Cfg::kChunks) * 32, tile_l0(tile), tile_b(tile));
      };
      for (uint32_t q = 0; q < (uint32_t)SR && q < total; ++q) load_res(q);
      for (uint32_t q = 0; q < total; ++q) {
        const int jt = (int)(q / Cfg::kChunks), ci = (int)(q % Cfg::kChunks);
        const int tile = tile_of(jt);
        const int b = tile_b(tile), l0 = tile_l0(tile);
        // staging slot and its use count: WIDE = one slot per team (chunks alternate between the teams)
        const int slot = Cfg::WIDE ? (ci & 1) : (int)(q % Cfg::kOutSlots);
        const uint32_t n_use = Cfg::WIDE ? (uint32_t)jt * (Cfg::kChunks / 2) + (uint32_t)(ci >> 1) : q / Cfg::kOutSlots;
        mbar_wait(out_full(slot), n_use & 1u);
        const uint32_t slab = res_base + (q % SR) * kRuSlabBytes;
        const uint32_t st_hi = out_base + (uint32_t)slot * (uint32_t)(Cfg::kPlanes * kRuPlaneTile);
        tma_store_3d(&tm_res, slab, ci * 32, l0, b);
        if (has_out) {
          tma_store_3d(&tm_o_hi, st_hi, ci * 32, l0, b);
          if (NTERMS >= 2) tma_store_3d(&tm_o_lo, st_hi + kRuPlaneTile, ci * 32, l0, b);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the TMA unit has read the slab and the slot
        mbar_arrive(out_empty(slot));
        if (q + SR < total) load_res(q + SR);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output stores have landed
    }
  } else {
    // ================================ epilogue teams ================================
    const bool fin_team = warp >= kRuFinWarp0;
    const int group = warp & 3;                 // TMEM lane quarter this warp may read
    const int row_in_tile = group * 32 + lane;

    // ---- mid stage of tile `it`: acc1 -> + b7 -> Snake -> bf16 hi/lo, written back IN PLACE: the 32 fp32 columns
    // of a chunk become 16 columns of packed hi pairs + 16 columns of packed lo pairs, which is exactly the
    // tensor-memory A-operand layout of the 1x1 conv's MMAs (lane = row, 2 K values per 32-bit column).
    // A warp converts whole chunks of its lane quarter; the `nsubs` warps of a quarter take the chunks
    // round-robin, so `nsubs` chunks are in flight.
    auto mid_stage = [&](int it, int sub, int nsubs) {
      const int buf = it % NB1;
      mbar_wait(acc1_full(buf), (uint32_t)(it / NB1) & 1u);
      tc_fence_after();
      if (warp == 0 && lane == 0) ru_trace(p, it, 10);
      const uint32_t t_row = tmem_base + ((uint32_t)(group * 32) << 16) + (uint32_t)buf * C;
#pragma unroll 1
      for (int ci = 0; ci < Cfg::kChunks; ++ci) {
        const uint32_t cg = (uint32_t)it * Cfg::kChunks + ci;   // running chunk number
        if ((int)(cg % (uint32_t)nsubs) != sub) continue;
        const int c = ci * 32;
        uint32_t r[32];
        tmem_ld_x32(t_row + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {     // 16 columns -> 8 packed hi columns + 8 packed lo columns
          uint32_t hi[8], lo[8];
          uint32_t l8[4], h8[4];   // two-term mode: four e5m2 values per column
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const int q = 4 * hf + qq;
            const float4 b4 = *reinterpret_cast<const float4*>(s_par + c + 4 * q);
            const float4 a4 = *reinterpret_cast<const float4*>(s_par + C + c + 4 * q);
            const float4 i4 = *reinterpret_cast<const float4*>(s_par + 2 * C + c + 4 * q);
            const float v0 = snake_sel<EXACT>(__uint_as_float(r[4 * q + 0]) + b4.x, a4.x, i4.x);
            const float v1 = snake_sel<EXACT>(__uint_as_float(r[4 * q + 1]) + b4.y, a4.y, i4.y);
            const float v2 = snake_sel<EXACT>(__uint_as_float(r[4 * q + 2]) + b4.z, a4.z, i4.z);
            const float v3 = snake_sel<EXACT>(__uint_as_float(r[4 * q + 3]) + b4.w, a4.w, i4.w);
            if (NTERMS == 2) {
              const Split4F8 sp = split4_f16f8(v0, v1, v2, v3);
              hi[2 * qq] = sp.hi[0];
              hi[2 * qq + 1] = sp.hi[1];
              l8[qq] = sp.lo8;
              h8[qq] = sp.hi8;
            } else {
              const __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
              hi[2 * qq] = pack_bf16(h0);
              hi[2 * qq + 1] = pack_bf16(h1);
              if (NTERMS == 3) {
                const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
                lo[2 * qq] = pack_bf16(__floats2bfloat162_rn(v0 - f0.x, v1 - f0.y));
                lo[2 * qq + 1] = pack_bf16(__floats2bfloat162_rn(v2 - f1.x, v3 - f1.y));
              }
            }
          }
          // every accumulator column of the chunk is already in registers, so the in-place stores are safe
          tmem_st_x8(t_row + c + 8 * hf, hi);
          if (NTERMS == 3) tmem_st_x8(t_row + c + 16 + 8 * hf, lo);
          if (NTERMS == 2) {   // 32 fp32 columns -> 16 of fp16 pairs + 8 of lo8 quads + 8 of hi8 quads
            tmem_st_x4(t_row + c + 16 + 4 * hf, l8);
            tmem_st_x4(t_row + c + 24 + 4 * hf, h8);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        // the 4 warps (one per lane quarter) that share `sub` converted this chunk: sync them, then ONE thread
        // arrives (a cluster-scope release arrive by every thread costs thousands of cycles in pair mode)
        asm volatile("bar.sync %0, 128;" ::"r"(2 + sub) : "memory");
        if (group == 0 && lane == 0) arrive_lead(mid_full(buf, ci));
        if (group == 0 && lane == 0 && ci < 6) ru_trace(p, it, 11 + ci);
      }
    };

    // ---- final stage: acc2 + residual slab + b1 -> x, Snake -> operand planes.
    // A thread owns 16 columns of ITS row (TMEM lane): it adds them into its row of the residual slab IN PLACE
    // (the slab then holds the new x tile in the TMA box layout), Snakes them straight from registers and
    // parks the bf16 hi/lo pairs in a staging tile (SWIZZLE_64B box layout), then arrives on the slot's `out_full`
    // mbarrier.  The OUTPUT THREAD (warp 16) issues the TMA stores of the slab and the staging tiles (no transposed
    // re-read, no per-thread global stores or address arithmetic, rows beyond the utterance and dummy tiles are clipped
    // by the TMA unit), hands the staging slot back (`out_empty`) once the TMA unit has read it and refills the slab.
    // When one thread of the team did this, every chunk paid its wait for the previous stores (~0.8 k cycles) plus the
    // issue of three TMA stores (~0.5 k) at the team's barrier.
    const int ew = warp & 7;                    // 0..7 within a team
    const int half = ew >> 2;                   // which 16 of a chunk's 32 columns this warp drains from TMEM
    // One 32-column chunk.  q = running chunk number of this CTA (residual ring position), slot / n_use = staging slot
    // and how often it has been used before (mbarrier parities), t_col = TMEM column of the chunk, hand_back = this
    // team is done with accumulator `acc_idx` after the chunk, bar_id = the team's named barrier.
    auto final_chunk = [&](int jt, int c, uint32_t q, int slot, uint32_t n_use, uint32_t t_col, bool hand_back,
                           uint32_t acc_idx, int bar_id, bool tr) {
      const bool has_out = p.out_hi != nullptr;
      const uint32_t rs_ = q % SR, rph_ = (q / SR) & 1u;
      uint32_t r[16];
      tmem_ld_x16(tmem_base + ((uint32_t)(group * 32) << 16) + t_col + 16 * half, r);
      tmem_ld_wait();
      tc_fence_before();
      if (hand_back) {
        // every thread of the team has drained its share of the accumulator: ONE thread hands the TMEM buffer back to
        // the (leader's) MMA warp -- in pair mode a cluster-scope release arrive
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kRuTeamThreads) : "memory");
        if (ew == 1 && lane == 0) arrive_lead(acc2_empty(acc_idx));
      }
      if (tr) ru_trace(p, jt, 14);
      const uint32_t slab = res_base + rs_ * kRuSlabBytes;
      const uint32_t st_hi = out_base + (uint32_t)slot * (uint32_t)(Cfg::kPlanes * kRuPlaneTile);
      const uint32_t st_lo = st_hi + kRuPlaneTile;
      mbar_wait(res_full(rs_), rph_);
      if (tr) ru_trace(p, jt, 15);
      // Plain C++ shared-memory accesses (no asm volatile): the compiler is free to issue all loads of the chunk
      // first and to interleave the 16 Snake chains; the mbarrier waits around the chunk are the compiler barriers.
      uint8_t* const slab_row = smem_raw + (slab - smem_u32(smem_raw)) + (size_t)row_in_tile * 128;
      uint8_t* const hi_row = smem_raw + (st_hi - smem_u32(smem_raw)) + (size_t)row_in_tile * 64;
      uint8_t* const lo_row = smem_raw + (st_lo - smem_u32(smem_raw)) + (size_t)row_in_tile * 64;
      const uint32_t swz64 = (uint32_t)(row_in_tile >> 1) & 3u;
      const int n0 = c + 16 * half;
      float4 v[4], b4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = *reinterpret_cast<const float4*>(slab_row + ((((uint32_t)(4 * half + j)) ^ ((uint32_t)row_in_tile & 7u)) << 4));
        b4[j] = *reinterpret_cast<const float4*>(s_par + 3 * C + n0 + 4 * j);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // same association as the stand-alone 1x1 kernel: (acc + residual) + bias
        v[j].x = (__uint_as_float(r[4 * j + 0]) + v[j].x) + b4[j].x;
        v[j].y = (__uint_as_float(r[4 * j + 1]) + v[j].y) + b4[j].y;
        v[j].z = (__uint_as_float(r[4 * j + 2]) + v[j].z) + b4[j].z;
        v[j].w = (__uint_as_float(r[4 * j + 3]) + v[j].w) + b4[j].w;
        *reinterpret_cast<float4*>(slab_row + ((((uint32_t)(4 * half + j)) ^ ((uint32_t)row_in_tile & 7u)) << 4)) = v[j];
      }
      // Snake + bf16 split into registers first; the staging slot is only needed for the stores below
      uint4 hp[2], lp[2];
      if (has_out) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {          // 8 columns = one 16 B piece of each operand plane
          uint32_t* hpp = reinterpret_cast<uint32_t*>(&hp[jj]);
          uint32_t* lpp = reinterpret_cast<uint32_t*>(&lp[jj]);
          float4 a4[2], i4[2];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            a4[h2] = *reinterpret_cast<const float4*>(s_par + 4 * C + n0 + 4 * (2 * jj + h2));
            i4[h2] = *reinterpret_cast<const float4*>(s_par + 5 * C + n0 + 4 * (2 * jj + h2));
          }
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int j = 2 * jj + h2;
            const float s0 = snake_sel<EXACT>(v[j].x, a4[h2].x, i4[h2].x), s1 = snake_sel<EXACT>(v[j].y, a4[h2].y, i4[h2].y);
            const float s2 = snake_sel<EXACT>(v[j].z, a4[h2].z, i4[h2].z), s3 = snake_sel<EXACT>(v[j].w, a4[h2].w, i4[h2].w);
            if (NTERMS == 2) {
              // lp[0] = the 16 lo8 bytes of this thread's 16 columns, lp[1] = the 16 hi8 bytes (word = quad j)
              const Split4F8 sp = split4_f16f8(s0, s1, s2, s3);
              hpp[2 * h2] = sp.hi[0];
              hpp[2 * h2 + 1] = sp.hi[1];
              reinterpret_cast<uint32_t*>(&lp[0])[j] = sp.lo8;
              reinterpret_cast<uint32_t*>(&lp[1])[j] = sp.hi8;
            } else {
              const __nv_bfloat162 h0 = __floats2bfloat162_rn(s0, s1), h1 = __floats2bfloat162_rn(s2, s3);
              hpp[2 * h2] = pack_bf16(h0);
              hpp[2 * h2 + 1] = pack_bf16(h1);
              if (NTERMS == 3) {
                const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
                lpp[2 * h2] = pack_bf16(__floats2bfloat162_rn(s0 - f0.x, s1 - f0.y));
                lpp[2 * h2 + 1] = pack_bf16(__floats2bfloat162_rn(s2 - f1.x, s3 - f1.y));
              }
            }
          }
        }
      }
      // The TMA unit has finished reading the slot's previous contents.  Waited for even when there are no operand
      // planes to stage (the last unit): without it the team could run up to SR chunks ahead of the output thread and
      // complete a SECOND phase of the slot's `out_full` barrier before the output thread has looked at the first --
      // a parity wait cannot tell two completed phases from none, and the output thread would wait forever (seen as a
      // rare bounded-wait trap after a cold start, when the first TMA stores are slow).
      mbar_wait(out_empty(slot), (n_use & 1u) ^ 1u);
      if (has_out) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const uint32_t off = (((uint32_t)(2 * half + jj)) ^ swz64) << 4;   // SWIZZLE_64B box layout
          *reinterpret_cast<uint4*>(hi_row + off) = hp[jj];
          if (NTERMS == 3) *reinterpret_cast<uint4*>(lo_row + off) = lp[jj];
          // two-term mode: the 64 B row of the packed plane is [lo8 x 32 | hi8 x 32]; jj = 0 lo8, jj = 1 hi8
          if (NTERMS == 2) *reinterpret_cast<uint4*>(lo_row + ((((uint32_t)(2 * jj + half)) ^ swz64) << 4)) = lp[jj];
        }
      }
      if (tr) ru_trace(p, jt, 16);
      fence_proxy_async();   // generic-proxy writes -> visible to the TMA unit's async-proxy reads
      // team barrier + ONE arrive (256 arrives on one mbarrier word serialise in shared memory: measured + 3-5 % at C = 96)
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kRuTeamThreads) : "memory");
      if (ew == 0 && lane == 0) mbar_arrive(out_full(slot));
      if (tr) ru_trace(p, jt, 17);
    };

    const int sub = (warp - kRuMidWarp0) >> 2;   // 0..1 mid team, 2..3 final team
    const int tm = warp >> 3;                    // team 0 / 1
    if (Cfg::WIDE) {
      // C = 384 (un-skewed, single acc1): all 16 warps convert the mid chunks of tile it (4 chunks in flight), then
      // BOTH teams drain the 1x1 accumulators on alternate 32-column chunks -- the 1x1 phase is not hidden behind a
      // k7 conv here, so the tensor pipe waits for whatever the drain does not keep up with.  Staging slot = team,
      // residual slabs of the team's parity.
      for (int it = 0; it < n_my; ++it) {
        mid_stage(it, sub, 4);
        for (int ci = tm; ci < Cfg::kChunks; ci += 2) {
          const uint32_t u = (uint32_t)it * NH + (uint32_t)(ci >> 1);
          mbar_wait(acc2_full(u % NB2), (u / NB2) & 1u);
          tc_fence_after();
          final_chunk(it, ci * 32, (uint32_t)it * Cfg::kChunks + (uint32_t)ci, tm, (uint32_t)it * (Cfg::kChunks / 2) + (uint32_t)(ci >> 1),
                      (uint32_t)(NB1 * C) + (u % NB2) * N2 + (uint32_t)(ci & 1) * 32u, true, u % NB2, 6 + tm, false);
        }
      }
    } else if (!fin_team) {
      // the teams work on different tiles at the same time: mid of tile i, final of tile i - 1
      for (int it = 0; it < n_my; ++it) mid_stage(it, sub, 2);
    } else {
      for (int jt = 0; jt < n_my; ++jt) {
#pragma unroll 1
        for (int nh = 0; nh < NH; ++nh) {            // N halves of the 1x1 conv (one accumulator each)
          const uint32_t u = (uint32_t)jt * NH + (uint32_t)nh;
          mbar_wait(acc2_full(u % NB2), (u / NB2) & 1u);
          tc_fence_after();
          if (warp == kRuFinWarp0 && lane == 0 && nh == 0) ru_trace(p, jt, 20);
#pragma unroll 1
          for (int cc = 0; cc < N2; cc += 32) {
            const int c = nh * N2 + cc;
            const uint32_t q = (uint32_t)jt * Cfg::kChunks + (uint32_t)(c / 32);
            final_chunk(jt, c, q, (int)(q % Cfg::kOutSlots), q / Cfg::kOutSlots, (uint32_t)(NB1 * C) + (u % NB2) * N2 + (uint32_t)cc,
                        cc + 32 >= N2, u % NB2, 1, C == 96 && cc == 64 && warp == kRuFinWarp0 && lane == 0);
            if (warp == kRuFinWarp0 && lane == 0 && c < 192) ru_trace(p, jt, 21 + c / 32);
          }
        }
      }
    }
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it / arrive on its barriers
  else __syncthreads();
  if (warp == kRuMmaWarp) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_cg2<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

template <int C, int NTERMS, int CL, bool PAIR>
int launch_ru(const GemmWeights& c7, const GemmWeights& c1, const OpBuf& a, int batch, int L, const RuParams& p,
              int num_sms, cudaStream_t stream) {
  using Cfg = RuCfg<C, NTERMS, PAIR ? 2 : 1>;
  CUtensorMap ta_hi, ta_lo, t_res;
  const uint64_t dims[3] = {(uint64_t)C, (uint64_t)L, (uint64_t)batch};
  const uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)L * C * 2};
  const uint32_t box[3] = {32u, (uint32_t)p.halo_rows, 1u};
  SC_TRY(encode_tmap(&ta_hi, a.hi, 3, dims, strides, box, 64, false, false));
  if (NTERMS >= 2) SC_TRY(encode_tmap(&ta_lo, a.lo, 3, dims, strides, box, 64, false, false));
  else ta_lo = ta_hi;
  const uint64_t rstr[2] = {(uint64_t)C * 4, (uint64_t)L * C * 4};
  const uint32_t rbox[3] = {32u, (uint32_t)kBlockM, 1u};
  SC_TRY(encode_tmap(&t_res, p.x, 3, dims, rstr, rbox, 128, false, true));
  // output operand planes (TMA stores): (C x L x batch) bf16, boxes of 32 columns x 128 rows, SWIZZLE_64B
  CUtensorMap to_hi = ta_hi, to_lo = ta_hi;
  if (p.out_hi) {
    const uint32_t obox[3] = {32u, (uint32_t)kBlockM, 1u};
    SC_TRY(encode_tmap(&to_hi, p.out_hi, 3, dims, strides, obox, 64, false, false));
    if (p.out_lo) SC_TRY(encode_tmap(&to_lo, p.out_lo, 3, dims, strides, obox, 64, false, false));
    else to_lo = to_hi;
  }
  // weight maps: (K, N) boxes of 32 columns x (N rows of one MMA) / CL rows: each CTA of a cluster fetches its share
  // of a stage.  The 1x1 conv of the split schedule (C = 192) is N2 = 96 wide.
  CUtensorMap tw[4];
  const GemmWeights* gw[2] = {&c7, &c1};
  for (int i = 0; i < 2; ++i) {
    const uint64_t wd[2] = {(uint64_t)gw[i]->kt * C, (uint64_t)C};
    const uint64_t ws[1] = {(uint64_t)gw[i]->kt * C * 2};
    const uint32_t wb[2] = {32u, (uint32_t)((i == 0 ? Cfg::N7 : Cfg::N2) / CL)};
    SC_TRY(encode_tmap(&tw[2 * i], gw[i]->hi_for(NTERMS), 2, wd, ws, wb, 64, true, false));
    SC_TRY(encode_tmap(&tw[2 * i + 1], NTERMS >= 2 ? gw[i]->lo_for(NTERMS) : gw[i]->w_hi, 2, wd, ws, wb, 64, true, false));
  }
  auto kern = resunit_fused_kernel<C, NTERMS, CL, PAIR>;
  static PerDevice cache;   // per instantiation and device: clusters that fit (0 = not initialised yet)
  int max_clusters = cache.here().load(std::memory_order_relaxed);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kRuThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  if (!max_clusters) {
    SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    if (CL > 1) {
      cfg.gridDim = dim3(num_sms / CL * CL);
      SC_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
      if (max_clusters < 1) { set_error("resunit_fused: no cluster of %d CTAs fits the device", CL); return SPARKCODEC_ECUDA; }
    } else {
      max_clusters = num_sms;
    }
    cache.here().store(max_clusters, std::memory_order_relaxed);
  }
  const int groups = (p.num_tiles + CL - 1) / CL;
  const int clusters = std::min(std::min(groups, max_clusters), num_sms / CL);
  cfg.gridDim = dim3(clusters * CL);
  static const bool trace = getenv("SPARKCODEC_RU_TRACE") != nullptr;   // schedule debugging only
  RuParams pp = p;
  static long long* dbg = nullptr;
  const size_t dbg_n = (size_t)kRuTraceTiles * kRuTraceEvents;
  if (trace) {
    if (!dbg) SC_CUDA(cudaMalloc(&dbg, dbg_n * sizeof(long long)));
    SC_CUDA(cudaMemsetAsync(dbg, 0, dbg_n * sizeof(long long), stream));
    pp.dbg = dbg;
  }
  cfg.numAttrs = add_pdl_attr(attr, CL > 1 ? 1 : 0);   // (the occupancy query above saw the cluster shape only)
  SC_CUDA(cudaLaunchKernelEx(&cfg, kern, ta_hi, ta_lo, tw[0], tw[1], tw[2], tw[3], t_res, to_hi, to_lo, pp));
  SC_LAUNCH_CHECK();
  if (trace) {
    std::vector<long long> hst(dbg_n);
    SC_CUDA(cudaStreamSynchronize(stream));
    SC_CUDA(cudaMemcpy(hst.data(), dbg, dbg_n * sizeof(long long), cudaMemcpyDeviceToHost));
    fprintf(stderr, "# ru trace C=%d terms=%d cl=%d dil=%d L=%d grid=%d (cycles relative to tile 2's k7 start; CTA 0)\n", C,
            NTERMS, CL, p.dil, L, clusters * CL);
    const long long t0 = hst[2 * kRuTraceEvents];
    for (int t = 2; t < 8; ++t) {
      fprintf(stderr, "tile %d:", t);
      for (int e = 0; e < kRuTraceEvents; ++e)
        if (hst[t * kRuTraceEvents + e]) fprintf(stderr, " e%d=%lld", e, hst[t * kRuTraceEvents + e] - (e >= 27 ? 0 : t0));
      fprintf(stderr, "\n");
    }
  }
  return 0;
}

}  // namespace

// k7 conv (single phase, 7 taps at multiples of one dilation <= 9) followed by a 1x1 conv, both C -> C with
// C in {96, 192, 384}: the shapes of the last three WaveGenerator stages (C = 768 would need 768 accumulator
// columns of tensor memory; that stage keeps its two kernels).
bool resunit_fusable(const GemmWeights& c7, const GemmWeights& c1, int* dil) {
  const int C = c7.c_in;
  static const int no_wide = [] { const char* e = getenv("SPARKCODEC_RU_WIDE"); return e && atoi(e) == 0; }();
  if (C != 96 && C != 192 && !(C == 384 && !no_wide)) return false;
  if (c7.n_total != C || c1.c_in != C || c1.n_total != C) return false;
  if (c7.taps.n_phase != 1 || c7.taps.ntaps[0] != 7 || c1.taps.n_phase != 1 || c1.taps.ntaps[0] != 1) return false;
  if (c1.taps.shift[0][0] != 0) return false;
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
  const bool f32 = is_split(precision);
  if (f32 && (!a.lo || (out.hi && !out.lo))) {
    set_error("resunit_fused: fp32 mode needs both operand planes");
    return SPARKCODEC_EINVAL;
  }
  if (a.fmt != op_fmt_for(precision) || (out.hi && out.fmt != a.fmt)) {
    set_error("resunit_fused: operand planes are not in the format of this precision mode");
    return SPARKCODEC_EINVAL;
  }
  const int terms = terms_for(precision);
  RuParams p;
  p.batch = batch; p.L = L; p.dil = dil;
  p.halo_rows = (kBlockM + 6 * dil + 7) / 8 * 8;
  p.m_tiles_per_utt = (L + kBlockM - 1) / kBlockM;
  p.num_tiles = batch * p.m_tiles_per_utt;
  p.bias7 = c7.bias; p.alpha_mid = alpha_mid; p.inv_mid = inv_mid;
  p.bias1 = c1.bias; p.alpha_out = alpha_out; p.inv_out = inv_out;
  p.x = x;
  p.dbg = nullptr;
  static const int l2pf = [] { const char* e = getenv("SPARKCODEC_RU_L2_PREFETCH"); return e ? atoi(e) : 0; }();
  p.l2_prefetch = l2pf;   // off: ncu showed 1.7x the algorithmic DRAM reads with it (evicted before use) and no speed-up
  p.out_hi = out.hi;
  p.out_lo = f32 ? out.lo : nullptr;
  // SPARKCODEC_CLUSTER: 1 = single CTAs, 2 = multicast clusters, 3 = CTA pairs; default: pairs in fp32 mode (tensor
  // pipe bound: 3 MMAs per MAC), multicast in bf16 mode at C <= 192 (epilogue bound) -- profiles/r1_pair_mode_ab.txt
  static const int forced = [] {
    const char* e = getenv("SPARKCODEC_CLUSTER");
    return e ? atoi(e) : 0;
  }();
  // (C = 384 in bf16 mode: pairs too -- 24 KB full-height weight stages leave a 3-4 deep ring that covers ~1.5 k cycles
  // of single-term MMAs, less than the refill round trip; A/B 2.70-2.81 ms against 2.94-2.98 ms per unit)
  const int mode = forced ? forced : ((f32 || c7.c_in == 384) ? 3 : 2);
#define RU_DISPATCH(CC, NT)                                                                              \
  return mode >= 3 ? launch_ru<CC, NT, 2, true>(c7, c1, a, batch, L, p, num_sms, stream)                 \
         : mode == 2 ? launch_ru<CC, NT, 2, false>(c7, c1, a, batch, L, p, num_sms, stream)              \
                     : launch_ru<CC, NT, 1, false>(c7, c1, a, batch, L, p, num_sms, stream)
  if (c7.c_in == 96) {
    if (terms == 3) { RU_DISPATCH(96, 3); } else if (terms == 2) { RU_DISPATCH(96, 2); } else { RU_DISPATCH(96, 1); }
  }
  if (c7.c_in == 384) {
    // fp32 modes: only the CTA pair fits (a full-height 48 KB weight stage would leave a 1-deep ring)
    if (terms == 3) return launch_ru<384, 3, 2, true>(c7, c1, a, batch, L, p, num_sms, stream);
    if (terms == 2) return launch_ru<384, 2, 2, true>(c7, c1, a, batch, L, p, num_sms, stream);
    RU_DISPATCH(384, 1);
  }
  if (terms == 3) { RU_DISPATCH(192, 3); } else if (terms == 2) { RU_DISPATCH(192, 2); } else { RU_DISPATCH(192, 1); }
#undef RU_DISPATCH
}

}  // namespace sparkcodec
