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
// Structure (persistent, warp specialised, one CTA per SM, 608 threads; optionally clusters of two CTAs that issue
// cta_group::2 M = 256 MMAs -- "pair mode"):
//   warp 0 / one lane  : TMA producer -> shared-memory rings (mbarrier full / empty).  Convs with several taps use the
//                        HALO mainloop: one activation tile with the conv halo per K chunk (all taps read it through
//                        row-offset descriptors) + a ring of per-(tap, K chunk) weight tiles; 1x1 convs / Linear layers
//                        use one ring of {A, W} stages
//   warp 1 / one lane  : tcgen05.mma issuer, accumulators in TMEM (2 stages of BLOCK_N columns)
//   warps 2..9, 10..17 : two epilogue teams on alternate 32-column chunks of an accumulator: tcgen05.ld -> swizzled
//                        transpose slab -> bias / residual / Snake / GELU -> coalesced fp32 + bf16 hi/lo stores,
//                        overlapped with the next tile's MMAs through the second TMEM stage
//   warp 18 / one lane : residual TMA producer (instantiations with a residual input)
// Precision modes: NTERMS == 1 : A_hi*W_hi (bf16);  NTERMS == 3 : A_hi*W_hi + A_lo*W_hi + A_hi*W_lo in bf16
// (error-compensated split, ~2^-16 relative per product, fp32 accumulate);  NTERMS == 2 (the default "fp32" mode):
// A_hi*W_hi in fp16 (kind::f16) + both cross terms as e5m2 products (kind::f8f6f4, K = 32 per MMA) read from the packed
// second plane -- see OPFMT_F16F8 in common.cuh.  Same bytes per stage as NTERMS == 3, two thirds of its tensor time.
#include <algorithm>
#include <cstdlib>

#include "gemm_params.cuh"
#include "tc_ptx.cuh"

namespace sparkcodec {

int choose_bk_halo(int c_in, int block_n, int precision);

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
  static constexpr int kSlabBytes = kBlockM * 128;   // epilogue transpose slab: 128 rows x 32 fp32, 128B-swizzled
  static constexpr int kResBytes = RES ? kResSlots * kSlabBytes : 0;   // residual slabs (same swizzled format)
  // smem ring budget: 227 KB - 2 slabs - residual ring - barriers - alignment slack
  static constexpr int kBudget = 192 * 1024 - kResBytes;
  // ring 1
  static constexpr int kAHaloBytes = align_up(kHaloRowsMax * BK * 2, 1024);            // per plane
  static constexpr int kStage1Bytes = HALO ? kPlanes * kAHaloBytes : kPlanes * (kABytes + kWBytes);
  // plain: ring depth fixed at compile time.  HALO: `halo_stages` (2 or 3, chosen per layer at launch time) halo
  // tiles, the rest of the budget goes to the weight ring -> depths are runtime values, barrier slots are laid
  // out for the maxima.
  static constexpr int kS1Raw = HALO ? 3 : kBudget / kStage1Bytes;
  static constexpr int kS1 = kS1Raw > 8 ? 8 : (kS1Raw < 1 ? 1 : kS1Raw);      // (HALO: maximum)
  // ring 2 (HALO only)
  static constexpr int kStage2Bytes = kPlanes * kWBytes;
  static constexpr int kS2Max = 8;
  static constexpr int s2_for(int halo_stages) {
    const int raw = (kBudget - halo_stages * kStage1Bytes) / kStage2Bytes;
    return raw > kS2Max ? kS2Max : (raw < 0 ? 0 : raw);
  }
  static constexpr int kS2 = HALO ? kS2Max : 0;                               // barrier slots
  static constexpr int kRingBytes = HALO ? kBudget : kS1 * kStage1Bytes;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kBarBytes = (2 * kS1 + 2 * kS2 + 4 + 2 * kResSlots) * 8 + 16;
  static constexpr int kSmemBytes = kRingBytes + 2 * kSlabBytes + kResBytes + kBarBytes + 1024 /* alignment */;
  static constexpr bool kValid = HALO ? (s2_for(2) >= 3) : (kS1Raw >= 2);
  static_assert(!kValid || kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
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
constexpr int kResWarp = 2 + kEpiTeams * kEpiWarps;
constexpr int kNumThreads = (kResWarp + 1) * 32;

template <int BLOCK_N, int BK, int NTERMS, bool RES, bool HALO, int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                    const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                    const __grid_constant__ CUtensorMap tm_res, const ConvGemmParams p) {
  using Cfg = TileCfg<BLOCK_N, BK, NTERMS, RES, HALO, CG>;
  constexpr int SB = Cfg::kS1, S2B = Cfg::kS2;                        // barrier slots (maxima)
  const int S = HALO ? p.halo_stages : Cfg::kS1;                       // ring depths actually used
  const int S2 = HALO ? Cfg::s2_for(p.halo_stages) : 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring2_base = smem_base + S * Cfg::kStage1Bytes;
  const uint32_t slab_base = smem_base + Cfg::kRingBytes;          // 2 epilogue slabs, 1024 B aligned
  const uint32_t res_base = slab_base + 2 * Cfg::kSlabBytes;       // kResSlots residual slabs (RES only)
  const uint32_t bar_base = res_base + Cfg::kResBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SB + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * SB + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * SB + 2 + a); };
  auto rfull_bar = [&](int r) { return bar_base + 8u * (2 * SB + 4 + r); };
  auto rempty_bar = [&](int r) { return bar_base + 8u * (2 * SB + 4 + kResSlots + r); };
  auto wfull_bar = [&](int s) { return bar_base + 8u * (2 * SB + 4 + 2 * kResSlots + s); };
  auto wempty_bar = [&](int s) { return bar_base + 8u * (2 * SB + 4 + 2 * kResSlots + S2B + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SB + 4 + 2 * kResSlots + 2 * S2B);
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
    for (int r = 0; r < kResSlots; ++r) {
      mbar_init(rfull_bar(r), 1);
      mbar_init(rempty_bar(r), kEpiThreads);
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
  } else if (warp == kResWarp) {
    // ================================ residual TMA producer ================================
    // Streams the fp32 residual tile (B, L, N) as 128-row x 32-column boxes into a ring of slabs that
    // have exactly the epilogue's swizzled slab format (SWIZZLE_128B), kResSlots boxes in flight.
    if (RES && elect_one()) {
      prefetch_tmap(&tm_res);
      uint32_t rs = 0, rphase = 0;
      for (int st = cid; st < num_super; st += ncl) {
        const Tile t = tile_of(st);
        const int b = t.b, l0 = t.l0, n0 = t.n0;
        for (int c = 0; c < BLOCK_N; c += 32) {
          mbar_wait(rempty_bar(rs), rphase ^ 1u);
          mbar_expect_tx(rfull_bar(rs), Cfg::kSlabBytes);
          tma_load_3d(res_base + rs * Cfg::kSlabBytes, &tm_res, rfull_bar(rs), n0 + c, l0, b);
          if (++rs == kResSlots) { rs = 0; rphase ^= 1u; }
        }
      }
    }
  } else {
    // ================================ epilogue warps ================================
    // Two phases per 32-column chunk so that every global access is coalesced:
    //  (1) tcgen05.ld hands each thread 32 columns of ITS row (TMEM lane); the thread parks them in a
    //      128B-swizzled smem slab (row r, 16 B chunk j -> chunk j ^ (r & 7): conflict-free both ways);
    //  (2) after a 128-thread named barrier the slab is read back transposed: 8 lanes cover the 32
    //      columns of one row (float4 each), a warp covers 4 rows per access, so residual loads and
    //      fp32 / bf16 stores are full 128 B / 64 B row segments.  Per-column parameters (bias, Snake
    //      alpha) are fixed per lane and loaded once per chunk.
    const int group = warp & 3;                 // TMEM lane quarter this warp may read
    const int row_in_tile = group * 32 + lane;
    const int team = (warp - 2) / kEpiWarps;    // 0 / 1: even / odd chunks of every accumulator
    const int ew = (warp - 2) % kEpiWarps;      // 0..7 within the team
    const int half = ew >> 2;                   // which 16 of a chunk's 32 columns this warp drains from TMEM
    const int q4 = lane & 7, rsub = lane >> 3;  // phase-2 mapping: column quad, row within a 4-row group
    constexpr int kChunks = BLOCK_N / 32;
    const uint32_t slab = slab_base + (uint32_t)team * Cfg::kSlabBytes;
    const int bar_a = 1 + 2 * team, bar_b = 2 + 2 * team;   // named barriers of this team
    uint32_t iter = 0;
    // The residual (fp32, may alias out_f32: every element is read by TMA before the thread that owns it
    // stores the sum) arrives through the slab ring filled by the residual producer warp, one slot per chunk.
    for (int st = cid; st < num_super; st += ncl, ++iter) {
      const Tile t = tile_of(st);
      const int b = t.b, l0 = t.l0, n0 = t.n0;
      const size_t tile_off = ((size_t)b * p.L + l0) * (size_t)p.n_total;
      const uint32_t acc = iter & 1u, acc_phase = (iter >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(group * 32) << 16) + acc * BLOCK_N;
      // chunks alternate between the teams; with an odd chunk count (BLOCK_N = 96: 3) the team that takes two
      // of them alternates from tile to tile, so both teams drain three chunks per two tiles
      const int first = (kChunks & 1) ? ((team + (int)iter) & 1) : team;
      const int my_last = ((kChunks - 1 - first) & ~1) + first;   // last chunk of this team (may be < 0: no chunk)
#pragma unroll 1
      for (int ci = first; ci < kChunks; ci += kEpiTeams) {
        const int c = ci * 32;
        // per-column parameters of this lane's 4 columns: issued first so their latency hides behind the
        // TMEM load, the slab write and the barrier
        const int n = n0 + c + q4 * 4;
        float4 bias4, alpha4, inv4;
        load_col_params4(p, n, bias4, alpha4, inv4);
        uint32_t r[16];
        tmem_ld_x16(t_row + c + 16 * half, r);
        // the previous chunk of this team has been read back out of the slab by every warp of the team
        asm volatile("bar.sync %0, %1;" ::"r"(bar_a), "n"(kEpiThreads) : "memory");
        tmem_ld_wait();
        {   // plain C++ shared-memory accesses: the named barriers / mbarrier waits are the compiler barriers
          uint8_t* const row_ptr = smem_raw + (slab - smem_u32(smem_raw)) + (size_t)row_in_tile * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(row_ptr + (((4 * half + j) ^ (row_in_tile & 7)) << 4)) =
                make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
        tc_fence_before();
        asm volatile("bar.sync %0, %1;" ::"r"(bar_b), "n"(kEpiThreads) : "memory");
        // every thread of the team has drained its share of the accumulator (tcgen05.wait::ld precedes the barrier):
        // after the team's last chunk ONE thread hands the TMEM stage back (a cluster-scope arrive by every thread
        // costs thousands of cycles in pair mode; the leader's MMA warp waits for both teams of both CTAs)
        if (ci == my_last && ew == 0 && lane == 0) {
          if (CG > 1) mbar_arrive_cluster_relaxed(mapa_shared(tempty_bar(acc), 0));   // tensor-memory hand-off
          else mbar_arrive(tempty_bar(acc));
        }
        uint32_t rslab = 0, rs = 0;
        if (RES) {
          const uint32_t cg = iter * kChunks + (uint32_t)ci;     // running chunk number -> ring slot / phase
          rs = cg % kResSlots;
          rslab = res_base + rs * Cfg::kSlabBytes;
          mbar_wait(rfull_bar(rs), (cg / kResSlots) & 1u);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r_ = ew * 16 + i * 4 + rsub;
          const uint32_t off = r_ * 128 + ((q4 ^ (r_ & 7)) << 4);
          float4 v = *reinterpret_cast<const float4*>(smem_raw + (slab - smem_u32(smem_raw)) + off);
          if (RES) {
            const float4 rr = *reinterpret_cast<const float4*>(smem_raw + (rslab - smem_u32(smem_raw)) + off);
            v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
          }
          if (l0 + r_ < p.L && p.dbg_skip_store == 0)
            epilogue_store4(p, v, b, tile_off + (size_t)((uint32_t)r_ * (uint32_t)p.n_total), n, bias4, alpha4, inv4,
                            /*add_residual=*/false);
        }
        if (RES) mbar_arrive(rempty_bar(rs));
      }
      if (my_last < 0 && ew == 0 && lane == 0) {   // (BLOCK_N == 32 only) a team without a chunk still releases the accumulator
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
  CUtensorMap t_res = ta_hi;
  if (RES) {
    const uint64_t rdims[3] = {(uint64_t)w.n_total, (uint64_t)L, (uint64_t)batch};
    const uint64_t rstr[2] = {(uint64_t)w.n_total * 4, (uint64_t)L * w.n_total * 4};
    const uint32_t rbox[3] = {32u, (uint32_t)kBlockM, 1u};
    SC_TRY(encode_map(&t_res, p.residual, 3, rdims, rstr, rbox, 64, false, /*fp32=*/true));
  }
  // weight maps: the cached (BK x BLOCK_N) boxes, or (BK x BLOCK_N / 2) boxes for a CTA pair
  constexpr int mi = BK == 64 ? 0 : 1;
  CUtensorMap tw_hi = NTERMS == 2 ? w.tmap_h16[mi] : w.tmap_hi[mi];
  CUtensorMap tw_lo = NTERMS == 3 ? w.tmap_lo[mi] : (NTERMS == 2 ? w.tmap_p8[mi] : w.tmap_hi[mi]);
  if (CG > 1) {
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
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = CG > 1 ? 1 : 0;
  if (!max_clusters) {
    SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
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
  SC_CUDA(cudaLaunchKernelEx(&cfg, kern, ta_hi, ta_lo, tw_hi, tw_lo, t_res, p));
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
int choose_bk(int c_in, int block_n, int precision, bool residual) {
  if (c_in % 64 != 0) return 32;
  const int planes = is_split(precision) ? 2 : 1;
  const int stage64 = planes * (kBlockM * 64 * 2 + block_n * 64 * 2);
  const int budget = 192 * 1024 - (residual ? kResSlots * kBlockM * 128 : 0);
  return budget / stage64 >= 3 ? 64 : 32;
}

// K chunk for the halo mainloop: the W ring must keep >= 4 stages next to the halo tiles.
int choose_bk_halo(int c_in, int block_n, int precision) {
  if (c_in % 64 != 0) return 32;
  const int planes = is_split(precision) ? 2 : 1;
  const int a64 = planes * align_up(kHaloRowsMax * 64 * 2, 1024), w64 = planes * block_n * 64 * 2;
  return (192 * 1024 - 2 * a64) / w64 >= 4 ? 64 : 32;
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
  // Measured per layer (profiles/r1_pair_mode_ab.txt): the pair wins where a tile carries a long reduction
  // (k7 convs, the wide up-samplers, pw2), and loses where the tile is short and the epilogue / HBM dominates
  // (pw1, the narrow 1x1 convs): the two CTAs of a pair run in lock step, which costs overlap there.
  // SPARKCODEC_PAIR = 0 never, 1 heuristic (default), 2 always.
  static const int pair_mode = [] { const char* e = getenv("SPARKCODEC_PAIR"); return e ? atoi(e) : 1; }();
  const int k_total = max_taps * w.c_in;
  // (SPARKCODEC_PAIR_MINK: threshold experiments)
  static const int pair_mink = [] { const char* e = getenv("SPARKCODEC_PAIR_MINK"); return e ? atoi(e) : 0; }();
  const int mink = pair_mink ? pair_mink : (f32 ? 768 : 1536);
  const bool pair = p.num_m_tiles >= 2 && (pair_mode == 2 || (pair_mode == 1 && k_total >= mink));
  bool halo = g_halo_mode != 0 && max_taps > 1 && ascending && !res;
  int bk = choose_bk(w.c_in, w.block_n, precision, res);
  if (halo) {
    p.halo_rows = (kBlockM + span + 7) / 8 * 8;
    // a halo tile is consumed over max_taps weight stages: with 7 taps two tiles in flight are plenty and the
    // weight ring gets the memory; the 2-3 tap polyphase branches turn their halo tiles over quickly
    static const int forced = [] { const char* e = getenv("SPARKCODEC_HALO_STAGES"); return e ? atoi(e) : 0; }();
    p.halo_stages = forced ? forced : (max_taps >= 5 ? 2 : 3);
    p.halo_bo_mode = 0;
    if (p.halo_rows > kHaloRowsMax) halo = false;
    else bk = choose_bk_halo(w.c_in, w.block_n, precision);
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
  if (w.block_n == BN && bk == BKK) {      \
    if (pair) { SC_INST2(BN, BKK, 2) }     \
    SC_INST2(BN, BKK, 1)                   \
  }
  SC_INST(256, 64) SC_INST(192, 64) SC_INST(128, 64) SC_INST(96, 64) SC_INST(64, 64)
  SC_INST(256, 32) SC_INST(192, 32) SC_INST(128, 32) SC_INST(96, 32) SC_INST(64, 32)
#undef SC_INST3
#undef SC_INST2
#undef SC_INST
  set_error("no tcgen05 instantiation for block_n=%d bk=%d", w.block_n, bk);
  return SPARKCODEC_EINVAL;
}

}  // namespace sparkcodec
