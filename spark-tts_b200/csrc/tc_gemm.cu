// tcgen05 implicit-GEMM convolution for sm_100a.
//
//   D[b, l, n] = sum_{tap j} sum_{c} A[b, l + shift_j, c] * W[n, j*C_in + c]        (+ fused epilogue)
//
// covers every dense contraction of the BiCodec detokenize path: k=7 (dilated) Conv1d, the 1x1 convs
// / Linear layers (one tap) and the ConvTranspose1d up-samplers (stride polyphase branches selected
// per N tile).  Activations are channels-last bf16 planes, so a tap shift is a row offset of the TMA
// box and TMA out-of-bounds zero fill IS the convolution's zero padding (3-D map C x L x batch:
// rows never bleed across utterances).
//
// Structure (persistent, warp specialised, one CTA per SM, 640 threads = 20 warps; optionally clusters of two CTAs that
// issue cta_group::2 M = 256 MMAs -- "pair mode"):
//   warp 0 / one lane  : TMA producer -> shared-memory rings (mbarrier full / empty).  Convs with several taps use the
//                        HALO mainloop: one activation tile with the conv halo per K chunk (all taps read it through
//                        row-offset descriptors) + a ring of per-(tap, K chunk) weight tiles; 1x1 convs / Linear layers
//                        use one ring of {A, W} stages
//   warp 1 / one lane  : tcgen05.mma issuer, accumulators in TMEM (2 stages of BLOCK_N columns)
//   warps 2..9, 10..17 : two epilogue teams on alternate 32-column chunks of an accumulator.  A thread owns 16 columns
//                        of its row: tcgen05.ld -> (+ residual slab, in place) + bias -> fp32 staging tile -> Snake /
//                        GELU -> operand-plane split -> plane staging tiles, all in the TMA box layouts; overlapped with
//                        the next tile's MMAs through the second TMEM stage
//   warps 18, 19 / one lane each : output threads, one per team: TMA stores of the staged tiles (fp32 and / or operand
//                        planes) and TMA loads of the fp32 residual slabs (instantiations with a residual input)
// Precision modes: NTERMS == 1 : A_hi*W_hi (bf16);  NTERMS == 3 : A_hi*W_hi + A_lo*W_hi + A_hi*W_lo in bf16
// (error-compensated split, ~2^-16 relative per product, fp32 accumulate);  NTERMS == 2 (the default "fp32" mode):
// A_hi*W_hi in fp16 (kind::f16) + both cross terms as e5m2 products (kind::f8f6f4, K = 32 per MMA) read from the packed
// second plane -- see OPFMT_F16F8 in common.cuh.  Same bytes per stage as NTERMS == 3, two thirds of its tensor time.
#include <algorithm>
#include <cstdlib>

#include "gemm_params.cuh"
#include "tc_ptx.cuh"

namespace sparkcodec {

int choose_bk_halo(int c_in, int block_n, int precision, int ring_bytes);

namespace {

constexpr int kResSlots = 4;   // residual slabs in flight (TMA, 16 KB each)
constexpr int kHaloRowsMax = 184;   // 128 + 6 * 9 (k=7, dilation 9) rounded up to 8
// Halo tiles in flight (runtime, 2 or 3).  One halo tile feeds all taps of a K chunk; the shared memory it does
// not need goes to the WEIGHT ring, which must cover the ~2k-cycle refill round trip (MMAs retire -> commit ->
// producer -> TMA from L2 -> full) or the tensor pipe starves.

constexpr int align_up(int v, int a) { return (v + a - 1) / a * a; }

// Two mainloop flavours share the kernel:
//  * plain  : one smem ring; a stage = {A_hi, A_lo, W_hi, W_lo} for one (tap, K-chunk); A is re-fetched per tap
//  * HALO   : (k>1 convs) ring 1 holds A tiles with the conv halo (128 + (k-1)*dilation rows) for one
//             K-chunk, fetched ONCE and used by all taps through row-offset smem descriptors; ring 2 holds
//             the per-(tap, K-chunk) W tiles.  Cuts the L2->smem operand traffic by 26-45 %.
template <int BLOCK_N, int BK, int NTERMS, bool RES, bool HALO, int CG>
struct TileCfg {
  static constexpr int kPlanes = (NTERMS >= 2) ? 2 : 1;
  static constexpr int kABytes = kBlockM * BK * 2;
  static constexpr int kWBytes = (BLOCK_N / CG) * BK * 2;   // CTA-pair mode: each CTA stages half of the N rows
  static constexpr int kSlabBytes = kBlockM * 128;   // 128 rows x 32 fp32, 128B-swizzled: residual slab / fp32 output staging
  static constexpr int kResBytes = RES ? kResSlots * kSlabBytes : 0;   // residual slabs
  // Shared memory: [operand ring(s)] [output staging of the two epilogue teams] [residual slabs] [barriers].  The
  // staging size depends on what the layer writes (fp32 and / or operand planes), so the ring budget and the ring
  // depths are RUNTIME values (ConvGemmParams::ring_bytes); kBudget is the largest ring any layer can get.
  static constexpr int kFixedBytes = kResBytes + 1024 /* barriers */ + 1024 /* alignment */;
  static constexpr int kBudget = 227 * 1024 - kFixedBytes - (RES ? 0 : 2 * kSlabBytes);
  // ring 1
  static constexpr int kAHaloBytes = align_up(kHaloRowsMax * BK * 2, 1024);            // per plane
  static constexpr int kStage1Bytes = HALO ? kPlanes * kAHaloBytes : kPlanes * (kABytes + kWBytes);
  // plain: ring depth fixed at compile time.  HALO: `halo_stages` (2 or 3, chosen per layer at launch time) halo
  // tiles, the rest of the budget goes to the weight ring -> depths are runtime values, barrier slots are laid
  // out for the maxima.
  static constexpr int kS1Raw = HALO ? 3 : kBudget / kStage1Bytes;
  static constexpr int kS1 = kS1Raw > 8 ? 8 : (kS1Raw < 1 ? 1 : kS1Raw);      // barrier slots (maximum depth)
  static constexpr int s1_for(int ring_bytes) {
    const int raw = ring_bytes / kStage1Bytes;
    return raw > kS1 ? kS1 : raw;
  }
  // ring 2 (HALO only)
  static constexpr int kStage2Bytes = kPlanes * kWBytes;
  static constexpr int kS2Max = 8;
  static constexpr int s2_for(int halo_stages, int ring_bytes) {
    const int raw = (ring_bytes - halo_stages * kStage1Bytes) / kStage2Bytes;
    return raw > kS2Max ? kS2Max : (raw < 0 ? 0 : raw);
  }
  static constexpr int kS2 = HALO ? kS2Max : 0;                               // barrier slots
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
  // barriers: ring full/empty, TMEM full/empty (2 + 2), residual full (kResSlots), staging full/empty per team (2 + 2),
  // weight ring full/empty
  static constexpr int kBarBytes = (2 * kS1 + 2 * kS2 + 4 + kResSlots + 4) * 8 + 16;
  static constexpr bool kValid = HALO ? (s2_for(2, kBudget) >= 3) : (kS1Raw >= 2);
  static_assert(kBarBytes <= 1024, "barrier block larger than budgeted");
  static_assert(2 * BLOCK_N <= 512, "two accumulator stages must fit TMEM");
  static_assert(kABytes % 1024 == 0 && kWBytes % 1024 == 0, "swizzled tiles must stay 1024 B aligned");
};

// warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 / 10..17 two epilogue teams (two warps per TMEM lane quarter
// and team, each draining half of a 32-column chunk), warp 18 residual TMA producer.
// The teams take alternate 32-column chunks of an accumulator, each with its own transpose slab, so two
// chunks are in flight: a chunk is a chain of long-latency steps (tcgen05.ld -> slab -> barrier -> slab ->
// SFU maths -> global stores) and ONE team left the short-reduction layers (pw1, the 1x1 convs, the narrow
// up-samplers) epilogue bound at ~1 element per cycle per SM.
constexpr int kEpiTeams = 2;
constexpr int kEpiWarps = 8;                    // per team
constexpr int kEpiThreads = kEpiWarps * 32;     // per team
constexpr int kResWarp = 2 + kEpiTeams * kEpiWarps;   // first of the two output warps (one per epilogue team)
constexpr int kNumThreads = (kResWarp + kEpiTeams) * 32;   // 20 warps: five per scheduler, 96 registers per thread

template <int BLOCK_N, int BK, int NTERMS, bool RES, bool HALO, int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                    const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                    const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_o_f32,
                    const __grid_constant__ CUtensorMap tm_o_hi, const __grid_constant__ CUtensorMap tm_o_lo,
                    const ConvGemmParams p) {
  using Cfg = TileCfg<BLOCK_N, BK, NTERMS, RES, HALO, CG>;
  constexpr int SB = Cfg::kS1, S2B = Cfg::kS2;                        // barrier slots (maxima)
  const int S = HALO ? p.halo_stages : Cfg::s1_for(p.ring_bytes);      // ring depths actually used
  const int S2 = HALO ? Cfg::s2_for(p.halo_stages, p.ring_bytes) : 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring2_base = smem_base + S * Cfg::kStage1Bytes;
  // output staging of the two epilogue teams: per team [fp32 tile (xs_bytes: 0 or 16 KB)] [operand-plane tiles (ps_bytes)]
  const uint32_t stage_base = smem_base + (uint32_t)p.ring_bytes;      // 1024 B aligned
  const uint32_t team_stage_bytes = (uint32_t)(p.xs_bytes + p.ps_bytes);
  const uint32_t res_base = stage_base + 2 * team_stage_bytes;        // kResSlots residual slabs (RES only)
  const uint32_t bar_base = res_base + Cfg::kResBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SB + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * SB + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * SB + 2 + a); };
  auto rfull_bar = [&](int r) { return bar_base + 8u * (2 * SB + 4 + r); };
  // a team's staging tiles: written by the team -> output thread (ofull) / read by the TMA unit -> team (oempty)
  auto ofull_bar = [&](int t) { return bar_base + 8u * (2 * SB + 4 + kResSlots + t); };
  auto oempty_bar = [&](int t) { return bar_base + 8u * (2 * SB + 4 + kResSlots + 2 + t); };
  auto wfull_bar = [&](int s) { return bar_base + 8u * (2 * SB + 4 + kResSlots + 4 + s); };
  auto wempty_bar = [&](int s) { return bar_base + 8u * (2 * SB + 4 + kResSlots + 4 + S2B + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SB + 4 + kResSlots + 4 + 2 * S2B);
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k_chunks = p.c_in / BK;
  // Tiles are dealt to clusters of CG CTAs: the CTAs of a cluster take CG consecutive M tiles of the SAME N tile
  // (CG == 2: one M = 256 pair MMA, each CTA stages half of the weight rows).  A trailing M tile without a
  // partner is a dummy (all rows out of range: TMA zero fill, no stores).
  const int rank = CG > 1 ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int cid = (int)blockIdx.x / CG, ncl = (int)gridDim.x / CG;
  const int num_super = ((p.num_m_tiles + CG - 1) / CG) * p.num_n_tiles;
  struct Tile { int n0, b, l0; };
  auto tile_of = [&](int st) {
    Tile t;
    t.n0 = (st % p.num_n_tiles) * BLOCK_N;
    const int m_tile = (st / p.num_n_tiles) * CG + rank;
    const bool valid = m_tile < p.num_m_tiles;
    t.b = valid ? m_tile / p.m_tiles_per_utt : p.batch;              // batch = out of range
    t.l0 = valid ? (m_tile % p.m_tiles_per_utt) * kBlockM : p.L;     // rows >= L are never stored
    return t;
  };
  // barrier that collects this stage's TMA bytes: the leader's (pair mode) or our own
  auto tx_bar = [&](uint32_t bar) { return CG > 1 ? mapa_shared(bar, 0) : bar; };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a_hi);
    prefetch_tmap(&tm_w_hi);
    if (NTERMS >= 2) {
      prefetch_tmap(&tm_a_lo);
      prefetch_tmap(&tm_w_lo);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < S2; ++s) {
      mbar_init(wfull_bar(s), 1);
      mbar_init(wempty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), CG * kEpiTeams);   // one arrive per epilogue team (pair mode: of both CTAs)
    }
    for (int r = 0; r < kResSlots; ++r) mbar_init(rfull_bar(r), 1);
    for (int t = 0; t < kEpiTeams; ++t) {
      mbar_init(ofull_bar(t), 1);
      mbar_init(oempty_bar(t), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG > 1) tmem_alloc_cg2<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if (CG > 1) cluster_sync_all();   // the peer's barriers exist before any remote arrive / TMA completion
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // PDL (common.cuh): everything above overlapped the previous kernel's tail; from here on its outputs are read
  griddep_wait();
  griddep_launch_dependents();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0, ws = 0, wphase = 0;
      const int wrow = rank * (BLOCK_N / CG);   // this CTA's share of the weight rows of an N tile
      auto load_a = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
        if (CG > 1) tma_load_3d_cg2(dst, map, tx_bar(bar), c0, c1, c2);
        else tma_load_3d(dst, map, bar, c0, c1, c2);
      };
      auto load_w = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
        if (CG > 1) tma_load_2d_cg2(dst, map, tx_bar(bar), c0, c1 + wrow);
        else tma_load_2d(dst, map, bar, c0, c1);
      };
      for (int st = cid; st < num_super; st += ncl) {
        const Tile t = tile_of(st);
        const int b = t.b, l0 = t.l0, n0 = t.n0;
        const int ph = n0 / p.taps.cols_per_phase;
        const int ntaps = p.taps.ntaps[ph];
        if (HALO) {
          const uint32_t a_tx = Cfg::kPlanes * p.halo_rows * BK * 2;
          for (int kc = 0; kc < k_chunks; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * Cfg::kStage1Bytes;
            if (leader) mbar_expect_tx(full_bar(stage), CG * a_tx);   // bytes of every CTA of the pair
            load_a(sa, &tm_a_hi, full_bar(stage), kc * BK, l0 + p.taps.shift[ph][0], b);
            if (NTERMS >= 2) load_a(sa + Cfg::kAHaloBytes, &tm_a_lo, full_bar(stage), kc * BK, l0 + p.taps.shift[ph][0], b);
            if (++stage == S) { stage = 0; phase ^= 1u; }
            for (int j = 0; j < ntaps; ++j) {
              mbar_wait(wempty_bar(ws), wphase ^ 1u);
              const uint32_t sw = ring2_base + ws * Cfg::kStage2Bytes;
              if (leader) mbar_expect_tx(wfull_bar(ws), CG * Cfg::kStage2Bytes);
              load_w(sw, &tm_w_hi, wfull_bar(ws), j * p.c_in + kc * BK, n0);
              if (NTERMS >= 2) load_w(sw + Cfg::kWBytes, &tm_w_lo, wfull_bar(ws), j * p.c_in + kc * BK, n0);
              if (++ws == S2) { ws = 0; wphase ^= 1u; }
            }
          }
        } else {
          for (int j = 0; j < ntaps; ++j) {
            const int row = l0 + p.taps.shift[ph][j];
            for (int kc = 0; kc < k_chunks; ++kc) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t sa = smem_base + stage * Cfg::kStage1Bytes;
              const uint32_t sw = sa + Cfg::kPlanes * Cfg::kABytes;
              if (leader) mbar_expect_tx(full_bar(stage), CG * Cfg::kStage1Bytes);
              load_a(sa, &tm_a_hi, full_bar(stage), kc * BK, row, b);
              load_w(sw, &tm_w_hi, full_bar(stage), j * p.c_in + kc * BK, n0);
              if (NTERMS >= 2) {
                load_a(sa + Cfg::kABytes, &tm_a_lo, full_bar(stage), kc * BK, row, b);
                load_w(sw + Cfg::kWBytes, &tm_w_lo, full_bar(stage), j * p.c_in + kc * BK, n0);
              }
              if (++stage == S) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // pair mode: the leader CTA's thread issues the M = 256 MMAs for both CTAs
    if (leader && elect_one()) {
      constexpr int HF = NTERMS == 2 ? 0 : 1;   // kind::f16 operand format: fp16 (two-term mode) or bf16
      constexpr uint32_t idesc = CG > 1 ? make_idesc_cg2<BLOCK_N, HF>() : make_idesc<BLOCK_N, HF>();
      constexpr uint32_t idesc8 = CG > 1 ? make_idesc_cg2<BLOCK_N, 1>() : make_idesc<BLOCK_N, 1>();   // e5m2 x e5m2
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t accumulate) {
        if (CG > 1) umma_bf16_cg2(d, a, b, idesc, accumulate);
        else umma_bf16(d, a, b, idesc, accumulate);
      };
      // cross terms of the two-term mode: 32 e5m2 values per K step = the same 32 B stride as a kind::f16 step; the
      // packed planes alternate [A_lo | A_hi] against [W_hi | W_lo] every 32 bytes
      auto mma8 = [&](uint32_t d, uint64_t a, uint64_t b) {
        if (CG > 1) umma_f8_cg2(d, a, b, idesc8, 1u);
        else umma_f8(d, a, b, idesc8, 1u);
      };
      auto commit = [&](uint32_t bar) {     // pair mode: arrives on the barrier at this offset in BOTH CTAs
        if (CG > 1) umma_commit_cg2(bar);
        else umma_commit(bar);
      };
      uint32_t stage = 0, phase = 0, iter = 0, ws = 0, wphase = 0;
      for (int st = cid; st < num_super; st += ncl, ++iter) {
        const int n0 = (st % p.num_n_tiles) * BLOCK_N;
        const int ph = n0 / p.taps.cols_per_phase;
        const int ntaps = p.taps.ntaps[ph];
        const uint32_t acc = iter & 1u, acc_phase = (iter >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        if (HALO) {
          for (int kc = 0; kc < k_chunks; ++kc) {
            mbar_wait(full_bar(stage), phase);
            const uint32_t sa = smem_base + stage * Cfg::kStage1Bytes;
            for (int j = 0; j < ntaps; ++j) {
              mbar_wait(wfull_bar(ws), wphase);
              tc_fence_after();
              const uint32_t sw = ring2_base + ws * Cfg::kStage2Bytes;
              // tap j reads the SAME halo tile, starting (shift_j - shift_0) rows further down
              const uint32_t a_off = (uint32_t)(p.taps.shift[ph][j] - p.taps.shift[ph][0]) * (BK * 2);
              const uint64_t a_hi = make_smem_desc_rows<BK>(sa + a_off, p.halo_bo_mode);
              const uint64_t w_hi = make_smem_desc<BK>(sw);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) mma(d_tmem, a_hi + 2 * k, w_hi + 2 * k, (kc | j | k) != 0);
              if (NTERMS >= 2) {
                const uint64_t a_lo = make_smem_desc_rows<BK>(sa + Cfg::kAHaloBytes + a_off, p.halo_bo_mode);
                const uint64_t w_lo = make_smem_desc<BK>(sw + Cfg::kWBytes);
                if (NTERMS == 2) {
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k) mma8(d_tmem, a_lo + 2 * k, w_lo + 2 * k);
                } else {
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k) mma(d_tmem, a_lo + 2 * k, w_hi + 2 * k, 1u);
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k) mma(d_tmem, a_hi + 2 * k, w_lo + 2 * k, 1u);
                }
              }
              commit(wempty_bar(ws));
              if (++ws == S2) { ws = 0; wphase ^= 1u; }
            }
            commit(empty_bar(stage));                        // halo tile free once all its taps retired
            if (kc == k_chunks - 1) commit(tfull_bar(acc));  // accumulator complete -> epilogue
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
        } else {
          const int n_k = ntaps * k_chunks;
          for (int ki = 0; ki < n_k; ++ki) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * Cfg::kStage1Bytes;
            const uint32_t sw = sa + Cfg::kPlanes * Cfg::kABytes;
            const uint64_t a_hi = make_smem_desc<BK>(sa), w_hi = make_smem_desc<BK>(sw);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)   // +32 B (=2 in >>4 units) per 16-element K step inside the swizzle row
              mma(d_tmem, a_hi + 2 * k, w_hi + 2 * k, (ki | k) != 0);
            if (NTERMS >= 2) {
              const uint64_t a_lo = make_smem_desc<BK>(sa + Cfg::kABytes);
              const uint64_t w_lo = make_smem_desc<BK>(sw + Cfg::kWBytes);
              if (NTERMS == 2) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) mma8(d_tmem, a_lo + 2 * k, w_lo + 2 * k);
              } else {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) mma(d_tmem, a_lo + 2 * k, w_hi + 2 * k, 1u);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) mma(d_tmem, a_hi + 2 * k, w_lo + 2 * k, 1u);
              }
            }
            commit(empty_bar(stage));                    // smem slot free once these MMAs retire
            if (ki == n_k - 1) commit(tfull_bar(acc));   // accumulator complete -> epilogue
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp >= kResWarp) {
    // ================================ output threads: residual slabs in, TMA stores out ================================
    // One per epilogue team (warp kResWarp + team).  The teams take alternate 32-column chunks of the CTA's tiles; with
    // q = running chunk number (tile iteration * kChunks + chunk) the team of chunk q is q & 1 in every case (an odd
    // chunk count alternates the first team from tile to tile), so this thread walks q = team, team + 2, ...  Per chunk:
    // wait until the team has written its staging tiles (and, with a residual, the new x into the residual slab), issue
    // the TMA stores, wait until the TMA unit has READ them, hand the staging back and refill the slab with the residual
    // of chunk q + kResSlots -- the thread that frees a slab (slots q % kResSlots of its own parity) refills it, so the
    // slab ring needs no "empty" barriers.  Rows beyond the utterance and dummy tiles are clipped by the TMA unit.
    // One thread for both teams was the bottleneck of the layers that write fp32 AND operand planes (32 KB per chunk).
    if (elect_one()) {
      constexpr int kChunks = BLOCK_N / 32;
      const int team = warp - kResWarp;
      const bool st_f32 = p.out_f32 != nullptr && p.dbg_skip_store == 0;
      const bool st_op = p.out_hi != nullptr && p.dbg_skip_store == 0;
      if (RES) prefetch_tmap(&tm_res);
      if (st_f32) prefetch_tmap(&tm_o_f32);
      if (st_op) prefetch_tmap(&tm_o_hi);
      const uint32_t my_tiles = cid < num_super ? (uint32_t)((num_super - cid + ncl - 1) / ncl) : 0u;
      const uint32_t total = my_tiles * kChunks;
      auto load_res = [&](uint32_t q) {
        const Tile t = tile_of(cid + (int)(q / kChunks) * ncl);
        const uint32_t rs = q % kResSlots;
        mbar_expect_tx(rfull_bar(rs), Cfg::kSlabBytes);
        tma_load_3d(res_base + rs * Cfg::kSlabBytes, &tm_res, rfull_bar(rs), t.n0 + (int)(q % kChunks) * 32, t.l0, t.b);
      };
      static_assert(kResSlots % kEpiTeams == 0, "each output thread owns the slabs of its parity");
      if (RES)
        for (uint32_t q = (uint32_t)team; q < (uint32_t)kResSlots && q < total; q += kEpiTeams) load_res(q);
      const uint32_t stg = stage_base + (uint32_t)team * team_stage_bytes;
      const uint32_t ps = stg + (uint32_t)p.xs_bytes;
      uint32_t used = 0;
      for (uint32_t q = (uint32_t)team; q < total; q += kEpiTeams, ++used) {
        const Tile t = tile_of(cid + (int)(q / kChunks) * ncl);
        const int ci = (int)(q % kChunks);
        mbar_wait(ofull_bar(team), used & 1u);
        const uint32_t xs = RES ? res_base + (q % kResSlots) * Cfg::kSlabBytes : stg;
        if (st_f32) tma_store_3d(&tm_o_f32, xs, t.n0 + ci * 32, t.l0, t.b);
        if (st_op) {
          tma_store_3d(&tm_o_hi, ps, t.n0 + ci * 32, t.l0, t.b);
          if (NTERMS >= 2) tma_store_3d(&tm_o_lo, ps + kBlockM * 64, t.n0 + ci * 32, t.l0, t.b);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the TMA unit has read the slab / the staging
        mbar_arrive(oempty_bar(team));
        if (RES && q + kResSlots < total) load_res(q + kResSlots);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output stores have landed
    }
  } else {
    // ================================ epilogue warps ================================
    // A thread owns 16 columns of ITS row (TMEM lane) of a 32-column chunk: tcgen05.ld -> + residual (slab row, in
    // place) + bias (+ per-utterance row bias) -> fp32 into the x staging tile -> Snake / GELU -> operand-plane split ->
    // plane staging tiles (all in the TMA box layouts), then the team arrives on its `ofull` barrier and the output
    // thread stores the tiles with TMA.  No transposed re-read, no per-thread global stores or address arithmetic.
    // (Before: tcgen05.ld -> transpose slab -> barrier -> transposed read-back -> per-thread global stores; the
    // short-reduction layers -- pw1, the narrow up-samplers -- were bound by exactly that.)
    const int group = warp & 3;                 // TMEM lane quarter this warp may read
    const int row_in_tile = group * 32 + lane;
    const int team = (warp - 2) / kEpiWarps;    // 0 / 1: alternate chunks of every accumulator
    const int ew = (warp - 2) % kEpiWarps;      // 0..7 within the team
    const int half = ew >> 2;                   // which 16 of a chunk's 32 columns this warp drains from TMEM
    constexpr int kChunks = BLOCK_N / 32;
    constexpr bool EXACT = NTERMS == 3;         // range-reduced sine only for the three-term split (gemm_params.cuh)
    // (Writing the fp32 output straight from the registers -- 64 contiguous bytes per thread, rows n_total apart --
    // was measured at + 30-90 % on the up-samplers: 32 row segments per warp store.  Everything goes out by TMA.)
    const bool st_f32 = p.out_f32 != nullptr, st_op = p.out_hi != nullptr;
    const uint32_t stg = stage_base + (uint32_t)team * team_stage_bytes;
    const int bar_a = 1 + 2 * team, bar_b = 2 + 2 * team;   // named barriers of this team
    uint8_t* const smem_gen = smem_raw - smem_u32(smem_raw);   // generic pointer of shared-window offset 0
    uint8_t* const xs_row = smem_gen + stg + (size_t)row_in_tile * 128;
    uint8_t* const hi_row = smem_gen + stg + p.xs_bytes + (size_t)row_in_tile * 64;
    uint8_t* const lo_row = hi_row + kBlockM * 64;
    const uint32_t swz128 = (uint32_t)row_in_tile & 7u, swz64 = (uint32_t)(row_in_tile >> 1) & 3u;
    uint32_t iter = 0, n_use = 0;
    for (int st = cid; st < num_super; st += ncl, ++iter) {
      const Tile t = tile_of(st);
      const int b = t.b, n0 = t.n0;
      const uint32_t acc = iter & 1u, acc_phase = (iter >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(group * 32) << 16) + acc * BLOCK_N;
      // chunks alternate between the teams; with an odd chunk count (BLOCK_N = 96: 3) the team that takes two
      // of them alternates from tile to tile, so both teams drain three chunks per two tiles
      const int first = (kChunks & 1) ? ((team + (int)iter) & 1) : team;
      const int my_last = ((kChunks - 1 - first) & ~1) + first;   // last chunk of this team (may be < 0: no chunk)
#pragma unroll 1
      for (int ci = first; ci < kChunks; ci += kEpiTeams, ++n_use) {
        const int c = ci * 32;
        const int nc = n0 + c + 16 * half;       // first of this thread's 16 output columns
        uint32_t r[16];
        tmem_ld_x16(t_row + c + 16 * half, r);
        // per-column parameters (same addresses in every lane: broadcast loads that hit L1), issued under the TMEM load
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(p.bias + nc) + j);
        tmem_ld_wait();
        tc_fence_before();
        // every thread of the team has drained its share of the accumulator: after the team's last chunk ONE thread
        // hands the TMEM stage back (the leader's MMA warp waits for both teams of both CTAs)
        if (ci == my_last) {
          asm volatile("bar.sync %0, %1;" ::"r"(bar_a), "n"(kEpiThreads) : "memory");
          if (ew == 1 && lane == 0) {
            if (CG > 1) mbar_arrive_cluster_relaxed(mapa_shared(tempty_bar(acc), 0));   // tensor-memory hand-off
            else mbar_arrive(tempty_bar(acc));
          }
        }
        uint8_t* x_row = xs_row;
        if (RES) {
          const uint32_t q = iter * kChunks + (uint32_t)ci;     // running chunk number -> residual slab / phase
          const uint32_t rs = q % kResSlots;
          mbar_wait(rfull_bar(rs), (q / kResSlots) & 1u);
          x_row = smem_gen + res_base + rs * Cfg::kSlabBytes + (size_t)row_in_tile * 128;
        }
        if (p.dbg_skip_store == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 a = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                   __uint_as_float(r[4 * j + 3]));
            if (RES) {   // same association as before: (acc + residual) + bias
              const float4 rr = *reinterpret_cast<const float4*>(x_row + ((((uint32_t)(4 * half + j)) ^ swz128) << 4));
              a.x += rr.x; a.y += rr.y; a.z += rr.z; a.w += rr.w;
            }
            v[j].x += a.x; v[j].y += a.y; v[j].z += a.z; v[j].w += a.w;
          }
          if (p.rowbias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 rb = __ldg(reinterpret_cast<const float4*>(p.rowbias + (size_t)b * p.n_total + nc) + j);
              v[j].x += rb.x; v[j].y += rb.y; v[j].z += rb.z; v[j].w += rb.w;
            }
          }
          if (RES && st_f32) {   // the slab is this chunk's own: the new x goes in right away
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(x_row + ((((uint32_t)(4 * half + j)) ^ swz128) << 4)) = v[j];
          }
        }
        // activation + operand-plane split into registers first: the staging tiles may still be read by the TMA unit
        uint4 hp[2], lp[2];
        if (st_op && p.dbg_skip_store == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float s0 = v[j].x, s1 = v[j].y, s2 = v[j].z, s3 = v[j].w;
            if (p.act == ACT_SNAKE) {
              const float4 al = __ldg(reinterpret_cast<const float4*>(p.alpha + nc) + j);
              const float4 iv = __ldg(reinterpret_cast<const float4*>(p.inv_alpha + nc) + j);
              s0 = snake_sel<EXACT>(s0, al.x, iv.x); s1 = snake_sel<EXACT>(s1, al.y, iv.y);
              s2 = snake_sel<EXACT>(s2, al.z, iv.z); s3 = snake_sel<EXACT>(s3, al.w, iv.w);
            } else if (p.act == ACT_GELU) {
              s0 = gelu_erf(s0); s1 = gelu_erf(s1); s2 = gelu_erf(s2); s3 = gelu_erf(s3);
            }
            uint32_t* hpp = reinterpret_cast<uint32_t*>(&hp[j >> 1]);
            if (NTERMS == 2) {   // lp[0] = the 16 lo8 bytes of this thread's columns, lp[1] = the 16 hi8 bytes
              const Split4F8 sp = split4_f16f8(s0, s1, s2, s3);
              hpp[2 * (j & 1)] = sp.hi[0];
              hpp[2 * (j & 1) + 1] = sp.hi[1];
              reinterpret_cast<uint32_t*>(&lp[0])[j] = sp.lo8;
              reinterpret_cast<uint32_t*>(&lp[1])[j] = sp.hi8;
            } else {
              const Split4Bf sp = split4_bf16(s0, s1, s2, s3, NTERMS == 3);
              hpp[2 * (j & 1)] = sp.hi[0];
              hpp[2 * (j & 1) + 1] = sp.hi[1];
              if (NTERMS == 3) {
                uint32_t* lpp = reinterpret_cast<uint32_t*>(&lp[j >> 1]);
                lpp[2 * (j & 1)] = sp.lo[0];
                lpp[2 * (j & 1) + 1] = sp.lo[1];
              }
            }
          }
        }
        // the TMA unit has finished reading this team's staging tiles of its previous chunk
        mbar_wait(oempty_bar(team), (n_use & 1u) ^ 1u);
        if (p.dbg_skip_store == 0) {
          if (!RES && st_f32) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(x_row + ((((uint32_t)(4 * half + j)) ^ swz128) << 4)) = v[j];
          }
          if (st_op) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const uint32_t off = (((uint32_t)(2 * half + jj)) ^ swz64) << 4;   // SWIZZLE_64B box layout
              *reinterpret_cast<uint4*>(hi_row + off) = hp[jj];
              if (NTERMS == 3) *reinterpret_cast<uint4*>(lo_row + off) = lp[jj];
              // two-term mode: the 64 B row of the packed plane is [lo8 x 32 | hi8 x 32]; jj = 0 lo8, jj = 1 hi8
              if (NTERMS == 2) *reinterpret_cast<uint4*>(lo_row + ((((uint32_t)(2 * jj + half)) ^ swz64) << 4)) = lp[jj];
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA unit
        asm volatile("bar.sync %0, %1;" ::"r"(bar_b), "n"(kEpiThreads) : "memory");
        if (ew == 0 && lane == 0) mbar_arrive(ofull_bar(team));
      }
      if (my_last < 0 && ew == 1 && lane == 0) {   // (BLOCK_N == 32 only) a team without a chunk still releases the accumulator
        if (CG > 1) mbar_arrive_cluster_relaxed(mapa_shared(tempty_bar(acc), 0));
        else mbar_arrive(tempty_bar(acc));
      }
    }
  }

  tc_fence_before();
  if (CG > 1) cluster_sync_all();   // no CTA leaves while its peer may still complete bytes / arrive on its barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG > 1) tmem_dealloc_cg2<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
PFN_tmapEncodeTiled g_encode = nullptr;
// 0: halo reuse off; 1: on, descriptor base-offset 0; 2: on, base-offset = start row phase.
// (env SPARKCODEC_HALO overrides; used to validate the descriptor convention on hardware)
int g_halo_mode = [] {
  const char* e = getenv("SPARKCODEC_HALO");
  return e ? atoi(e) : 1;
}();

}  // namespace

// programmatic dependent launch of the pass's kernels (common.cuh); SPARKCODEC_PDL=0 disables
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("SPARKCODEC_PDL");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}

// MMA terms of the fp32 precision mode (common.cuh): 2 = fp16 main term + two e5m2 cross terms (default), 3 = bf16 x 3
int fp32_terms() {
  static const int t = [] {
    const char* e = getenv("SPARKCODEC_FP32_TERMS");
    const int v = e ? atoi(e) : 2;
    return v == 3 ? 3 : 2;
  }();
  return t;
}

namespace {

int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, int bk, bool weights, bool fp32 = false) {
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bx[3], es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(map, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                        bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        weights ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return SPARKCODEC_ECUDA;
  }
  return 0;
}

template <int BLOCK_N, int BK, int NTERMS, bool RES, bool HALO, int CG>
int launch_inst(const GemmWeights& w, const OpBuf& a, int batch, int L, const ConvGemmParams& p, int num_sms,
                cudaStream_t stream) {
  using Cfg = TileCfg<BLOCK_N, BK, NTERMS, RES, HALO, CG>;
  if constexpr (!Cfg::kValid) {
    set_error("tile %dx%d (terms %d, residual %d, halo %d) does not fit shared memory", BLOCK_N, BK, NTERMS,
              (int)RES, (int)HALO);
    return SPARKCODEC_EINVAL;
  } else {
  CUtensorMap ta_hi, ta_lo;
  const uint64_t dims[3] = {(uint64_t)w.c_in, (uint64_t)L, (uint64_t)batch};
  const uint64_t strides[2] = {(uint64_t)w.c_in * 2, (uint64_t)L * w.c_in * 2};
  const uint32_t box[3] = {(uint32_t)BK, (uint32_t)(HALO ? p.halo_rows : kBlockM), 1u};
  SC_TRY(encode_map(&ta_hi, a.hi, 3, dims, strides, box, BK, false));
  if (NTERMS >= 2) SC_TRY(encode_map(&ta_lo, a.lo, 3, dims, strides, box, BK, false));
  else ta_lo = ta_hi;
  if (a.fmt != (NTERMS == 2 ? OPFMT_F16F8 : OPFMT_BF16)) {
    set_error("conv_gemm: operand planes are in format %d, the %d-term kernel needs the other one", a.fmt, NTERMS);
    return SPARKCODEC_EINVAL;
  }
  // residual in / outputs: (n_total x L x batch) tensors, boxes of 32 columns x 128 rows
  CUtensorMap t_res = ta_hi, to_f32 = ta_hi, to_hi = ta_hi, to_lo = ta_hi;
  const uint64_t rdims[3] = {(uint64_t)w.n_total, (uint64_t)L, (uint64_t)batch};
  const uint64_t rstr[2] = {(uint64_t)w.n_total * 4, (uint64_t)L * w.n_total * 4};
  const uint64_t ostr[2] = {(uint64_t)w.n_total * 2, (uint64_t)L * w.n_total * 2};
  const uint32_t rbox[3] = {32u, (uint32_t)kBlockM, 1u};
  if (RES) SC_TRY(encode_map(&t_res, p.residual, 3, rdims, rstr, rbox, 64, false, /*fp32=*/true));
  if (p.out_f32) SC_TRY(encode_map(&to_f32, p.out_f32, 3, rdims, rstr, rbox, 64, false, /*fp32=*/true));
  if (p.out_hi) {
    SC_TRY(encode_map(&to_hi, p.out_hi, 3, rdims, ostr, rbox, 32, false));
    if (NTERMS >= 2) SC_TRY(encode_map(&to_lo, p.out_lo, 3, rdims, ostr, rbox, 32, false));
  }
  // shared memory: what the output staging of the two epilogue teams leaves goes to the operand ring(s)
  ConvGemmParams pp = p;
  pp.xs_bytes = (p.out_f32 && !RES) ? Cfg::kSlabBytes : 0;
  pp.ps_bytes = p.out_hi ? Cfg::kPlanes * kBlockM * 64 : 0;
  pp.ring_bytes = (227 * 1024 - Cfg::kFixedBytes - 2 * (pp.xs_bytes + pp.ps_bytes)) / 1024 * 1024;
  if (HALO) {
    if (Cfg::s2_for(pp.halo_stages, pp.ring_bytes) < 3 && pp.halo_stages > 2) pp.halo_stages = 2;
    if (Cfg::s2_for(pp.halo_stages, pp.ring_bytes) < 3) {
      set_error("conv_gemm: weight ring too shallow (tile %dx%d, %d terms)", BLOCK_N, BK, NTERMS);
      return SPARKCODEC_EINVAL;
    }
  } else if (Cfg::s1_for(pp.ring_bytes) < 2) {
    set_error("conv_gemm: operand ring too shallow (tile %dx%d, %d terms)", BLOCK_N, BK, NTERMS);
    return SPARKCODEC_EINVAL;
  }
  const int smem_bytes = pp.ring_bytes + 2 * (pp.xs_bytes + pp.ps_bytes) + Cfg::kFixedBytes;
  // weight maps: the cached (BK x BLOCK_N) boxes, or (BK x BLOCK_N / 2) boxes for a CTA pair
  constexpr int mi = BK == 64 ? 0 : 1;
  CUtensorMap tw_hi = NTERMS == 2 ? w.tmap_h16[mi] : w.tmap_hi[mi];
  CUtensorMap tw_lo = NTERMS == 3 ? w.tmap_lo[mi] : (NTERMS == 2 ? w.tmap_p8[mi] : w.tmap_hi[mi]);
  if (CG > 1 || BLOCK_N != w.block_n) {   // (a narrowed N tile of a few-tile launch has no cached map either)
    const uint64_t wd[2] = {(uint64_t)w.kt * w.c_in, (uint64_t)w.n_total};
    const uint64_t ws[1] = {(uint64_t)w.kt * w.c_in * 2};
    const uint32_t wb[2] = {(uint32_t)BK, (uint32_t)(BLOCK_N / CG)};
    SC_TRY(encode_map(&tw_hi, w.hi_for(NTERMS), 2, wd, ws, wb, BK, true));
    if (NTERMS >= 2) SC_TRY(encode_map(&tw_lo, w.lo_for(NTERMS), 2, wd, ws, wb, BK, true));
    else tw_lo = tw_hi;
  }
  auto kern = conv_gemm_tc_kernel<BLOCK_N, BK, NTERMS, RES, HALO, CG>;
  static PerDevice cache;   // per instantiation and device: clusters that fit (0 = not initialised yet)
  int max_clusters = cache.here().load(std::memory_order_relaxed);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = CG > 1 ? 1 : 0;   // (the occupancy query below sees the cluster shape only)
  if (!max_clusters) {
    SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (CG > 1) {
      cfg.gridDim = dim3(num_sms / CG * CG);
      SC_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
      if (max_clusters < 1) { set_error("conv_gemm: no cluster of %d CTAs fits the device", CG); return SPARKCODEC_ECUDA; }
    } else {
      max_clusters = num_sms;
    }
    cache.here().store(max_clusters, std::memory_order_relaxed);
  }
  const int supers = ((p.num_m_tiles + CG - 1) / CG) * p.num_n_tiles;
  const int clusters = std::min(std::min(supers, max_clusters), num_sms / CG);
  cfg.gridDim = dim3(clusters * CG);
  if (CG == 1) cfg.numAttrs = 0;
  cfg.numAttrs = add_pdl_attr(attr, cfg.numAttrs);
  SC_CUDA(cudaLaunchKernelEx(&cfg, kern, ta_hi, ta_lo, tw_hi, tw_lo, t_res, to_f32, to_hi, to_lo, pp));
  SC_LAUNCH_CHECK();
  return 0;
  }
}

}  // namespace

int tma_init();
int encode_tmap(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, int swizzle_bytes, bool weights, bool fp32) {
  SC_TRY(tma_init());
  return encode_map(map, base, rank, dims, strides_bytes, box, swizzle_bytes == 128 ? 64 : 32, weights, fp32);
}

int tma_init() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return SPARKCODEC_ECUDA;
  }
  g_encode = reinterpret_cast<PFN_tmapEncodeTiled>(fn);
  return 0;
}

// N tile per layer: BLOCK_N must divide the per-phase width so an N tile never straddles two polyphase
// branches.
int choose_block_n(int cols_per_phase, int* block_n) {
  const int cand[] = {256, 192, 128, 96, 64};
  for (int c : cand)
    if (cols_per_phase % c == 0) { *block_n = c; return 0; }
  set_error("no tile width divides C_out=%d (need a multiple of 64 or 96)", cols_per_phase);
  return SPARKCODEC_EINVAL;
}

// K chunk per (layer, precision): 64 bf16 (128 B swizzle rows) when C_in allows it and the smem ring
// still gets >= 3 stages, else 32 (64 B swizzle rows).
int choose_bk(int c_in, int block_n, int precision, int ring_bytes) {
  if (c_in % 64 != 0) return 32;
  const int planes = is_split(precision) ? 2 : 1;
  const int stage64 = planes * (kBlockM * 64 * 2 + block_n * 64 * 2);
  return ring_bytes / stage64 >= 3 ? 64 : 32;
}

// K chunk for the halo mainloop: the W ring must keep >= 4 stages next to the halo tiles.
int choose_bk_halo(int c_in, int block_n, int precision, int ring_bytes) {
  if (c_in % 64 != 0) return 32;
  const int planes = is_split(precision) ? 2 : 1;
  const int a64 = planes * align_up(kHaloRowsMax * 64 * 2, 1024), w64 = planes * block_n * 64 * 2;
  return (ring_bytes - 2 * a64) / w64 >= 4 ? 64 : 32;
}

// N tile of one launch: the packed width, narrowed for few-tile launches (see launch_conv_gemm_tc)
int launch_block_n(int packed_bn, int n_total, int cols_per_phase, int num_m_tiles, int num_sms) {
  static const int small_n = [] { const char* e = getenv("SPARKCODEC_SMALL_N"); return e ? atoi(e) : 1; }();
  int bn = packed_bn;
  if (small_n) {
    const int cands[] = {128, 96, 64};
    for (int c : cands) {
      if (2 * num_m_tiles * (n_total / bn) > num_sms) break;
      if (c >= bn || cols_per_phase % c != 0) continue;
      bn = c;
    }
  }
  return bn;
}

int make_weight_tmaps(GemmWeights& w) {
  SC_TRY(tma_init());
  if (w.c_in % 32 != 0) { set_error("C_in=%d is not a multiple of 32", w.c_in); return SPARKCODEC_EINVAL; }
  SC_TRY(choose_block_n(w.taps.cols_per_phase, &w.block_n));
  w.has_bk64 = w.c_in % 64 == 0;
  const uint64_t dims[2] = {(uint64_t)w.kt * w.c_in, (uint64_t)w.n_total};
  const uint64_t strides[1] = {(uint64_t)w.kt * w.c_in * 2};
  for (int mi = 0; mi < 2; ++mi) {
    const int bk = mi == 0 ? 64 : 32;
    if (mi == 0 && !w.has_bk64) continue;
    const uint32_t box[2] = {(uint32_t)bk, (uint32_t)w.block_n};
    SC_TRY(encode_map(&w.tmap_hi[mi], w.w_hi, 2, dims, strides, box, bk, true));
    SC_TRY(encode_map(&w.tmap_lo[mi], w.w_lo, 2, dims, strides, box, bk, true));
    SC_TRY(encode_map(&w.tmap_h16[mi], w.w_h16, 2, dims, strides, box, bk, true));
    SC_TRY(encode_map(&w.tmap_p8[mi], w.w_p8, 2, dims, strides, box, bk, true));
  }
  return 0;
}

int fill_params(const GemmWeights& w, int batch, int L, const Epilogue& ep, int precision, ConvGemmParams* p) {
  if ((ep.out_op.hi != nullptr) && is_split(precision) && ep.out_op.lo == nullptr) {
    set_error("fp32 mode needs both operand planes");
    return SPARKCODEC_EINVAL;
  }
  p->batch = batch; p->L = L; p->n_total = w.n_total; p->c_in = w.c_in;
  p->taps = w.taps;
  p->m_tiles_per_utt = (L + kBlockM - 1) / kBlockM;
  p->num_m_tiles = batch * p->m_tiles_per_utt;
  p->num_n_tiles = w.n_total / w.block_n;
  p->halo_rows = 0; p->halo_bo_mode = 0; p->halo_stages = 3;
  p->ring_bytes = 0; p->xs_bytes = 0; p->ps_bytes = 0;
  static const int skip = [] { const char* e = getenv("SPARKCODEC_DEBUG_SKIP_STORE"); return e ? atoi(e) : 0; }();
  p->dbg_skip_store = skip;   // timing experiments only: drain the accumulators but do not finish / store them
  p->bias = w.bias; p->rowbias = ep.rowbias; p->residual = ep.residual;
  p->alpha = ep.alpha; p->inv_alpha = ep.inv_alpha; p->act = ep.act;
  p->out_f32 = ep.out_f32; p->out_hi = ep.out_op.hi;
  p->out_lo = is_split(precision) ? ep.out_op.lo : nullptr;
  p->out_fmt = ep.out_op.fmt;
  if (ep.out_op.hi != nullptr && ep.out_op.fmt != op_fmt_for(precision)) {
    set_error("output operand planes are not in the format of this precision mode");
    return SPARKCODEC_EINVAL;
  }
  return 0;
}

int launch_conv_gemm_tc(const GemmWeights& w, const OpBuf& a, int batch, int L, const Epilogue& ep, int precision,
                        int num_sms, cudaStream_t stream) {
  ConvGemmParams p;
  SC_TRY(fill_params(w, batch, L, ep, precision, &p));
  const bool f32 = is_split(precision);
  const int terms = terms_for(precision);
  const bool res = ep.residual != nullptr;
  // halo reuse: every conv with more than one tap (k=7 convs, conv-in, embed convs, polyphase up-samplers)
  int max_taps = 0, span = 0;
  bool ascending = true;
  for (int r = 0; r < w.taps.n_phase; ++r) {
    const int nt = w.taps.ntaps[r];
    max_taps = nt > max_taps ? nt : max_taps;
    if (nt > 0 && w.taps.shift[r][nt - 1] - w.taps.shift[r][0] > span) span = w.taps.shift[r][nt - 1] - w.taps.shift[r][0];
    for (int j = 1; j < nt; ++j) ascending &= w.taps.shift[r][j] > w.taps.shift[r][j - 1];
  }
  // CTA-pair mode (cta_group::2, M = 256): each CTA stages half of the weight rows, which halves the weight
  // traffic per SM and doubles the depth of the weight ring in tensor-pipe time.  SPARKCODEC_PAIR=0 disables.
  // Round 1 (transposing epilogue, three-term fp32 mode; profiles/r1_pair_mode_ab.txt): the pair won where a tile
  // carries a long reduction (k7 convs, the wide up-samplers, pw2) and lost on short tiles whose epilogue dominated
  // (pw1, the narrow 1x1 convs), so the threshold was K >= 768 (1536 in bf16 mode).
  // SPARKCODEC_PAIR = 0 never, 1 threshold on K (default), 2 always; SPARKCODEC_PAIR_MINK overrides the threshold.
  static const int pair_mode = [] { const char* e = getenv("SPARKCODEC_PAIR"); return e ? atoi(e) : 1; }();
  const int k_total = max_taps * w.c_in;
  static const int pair_mink = [] { const char* e = getenv("SPARKCODEC_PAIR_MINK"); return e ? atoi(e) : 0; }();
  // (round 2, TMA-store epilogue + two-term fp32 mode: with the epilogue off the teams' critical path the pair wins from
  // K = 384 on in both modes -- pw1 -12 %, the narrow up-samplers -3..-11 %; profiles/r2_pair_threshold_ab.txt)
  const int mink = pair_mink ? pair_mink : 384;
  const bool pair = p.num_m_tiles >= 2 && (pair_mode == 2 || (pair_mode == 1 && k_total >= mink));
  bool halo = g_halo_mode != 0 && max_taps > 1 && ascending && !res;
  // operand-ring budget of this layer: what the residual slabs and the output staging of the two epilogue teams leave
  // (launch_inst computes the same figure from the instantiation's constants)
  const int staging = 2 * (((ep.out_f32 && !res) ? kBlockM * 128 : 0) + (ep.out_op.hi ? (f32 ? 2 : 1) * kBlockM * 64 : 0));
  const int ring_bytes = 227 * 1024 - 2048 - (res ? kResSlots * kBlockM * 128 : 0) - staging;
  // Few-tile launches (single utterances, short chunks): with the packed tile width a layer's whole weight matrix
  // streams through a handful of SMs at ~64 B/clk each (conv-in at 1 x 500 frames: 24 tiles, 3.7 MB of weights per
  // CTA) while the rest of the chip idles.  A narrower N tile spreads the stream: narrow while at most half of the SMs
  // would be busy.  The K chunk stays the one the packed width gets (a narrower tile only deepens the ring), so every
  // output element sees the same sequence of MMAs -- in the two-term mode the fp16 and e5m2 products alternate per K
  // chunk -- and the result has the same bits; launches with more tiles than that (every layer of the batched
  // configs) keep the packed width.  SPARKCODEC_SMALL_N=0 disables.
  const int bn = launch_block_n(w.block_n, w.n_total, w.taps.cols_per_phase, p.num_m_tiles, num_sms);
  p.num_n_tiles = w.n_total / bn;
  int bk = choose_bk(w.c_in, w.block_n, precision, ring_bytes);
  if (halo) {
    p.halo_rows = (kBlockM + span + 7) / 8 * 8;
    // a halo tile is consumed over max_taps weight stages: with 7 taps two tiles in flight are plenty and the
    // weight ring gets the memory; the 2-3 tap polyphase branches turn their halo tiles over quickly
    static const int forced = [] { const char* e = getenv("SPARKCODEC_HALO_STAGES"); return e ? atoi(e) : 0; }();
    p.halo_stages = forced ? forced : (max_taps >= 5 ? 2 : 3);
    p.halo_bo_mode = 0;
    if (p.halo_rows > kHaloRowsMax) halo = false;
    else bk = choose_bk_halo(w.c_in, w.block_n, precision, ring_bytes);
  }
#define SC_INST3(BN, BKK, CGV, RESV, HALOV)                                                            \
    return terms == 3 ? launch_inst<BN, BKK, 3, RESV, HALOV, CGV>(w, a, batch, L, p, num_sms, stream)  \
         : terms == 2 ? launch_inst<BN, BKK, 2, RESV, HALOV, CGV>(w, a, batch, L, p, num_sms, stream)  \
                      : launch_inst<BN, BKK, 1, RESV, HALOV, CGV>(w, a, batch, L, p, num_sms, stream);
#define SC_INST2(BN, BKK, CGV)                       \
    if (halo) { SC_INST3(BN, BKK, CGV, false, true) } \
    if (res) { SC_INST3(BN, BKK, CGV, true, false) }  \
    SC_INST3(BN, BKK, CGV, false, false)
#define SC_INST(BN, BKK)                   \
  if (bn == BN && bk == BKK) {             \
    if (pair) { SC_INST2(BN, BKK, 2) }     \
    SC_INST2(BN, BKK, 1)                   \
  }
  SC_INST(256, 64) SC_INST(192, 64) SC_INST(128, 64) SC_INST(96, 64) SC_INST(64, 64)
  SC_INST(256, 32) SC_INST(192, 32) SC_INST(128, 32) SC_INST(96, 32) SC_INST(64, 32)
#undef SC_INST3
#undef SC_INST2
#undef SC_INST
  set_error("no tcgen05 instantiation for block_n=%d bk=%d", bn, bk);
  return SPARKCODEC_EINVAL;
}

}  // namespace sparkcodec
