// Main 3x3 reflect-pad convolution kernel: tcgen05 implicit GEMM fed by TMA (see conv_umma_impl.cuh).
#pragma once
#include "umma_common.cuh"

namespace ccst {
namespace {

// Shared-memory plan of the main kernel (tile = 8 x 16 output pixels x BN output channels).
//   A ring : kAStages slabs of {64 ch, 16 px, 10 rows} = 20 KiB.  One slab serves the three filter
//            rows r = 0,1,2 of one (channel chunk, filter column s): the operand of tap (r,s) is the
//            slab shifted by r tile rows = r * 2048 bytes, which keeps the 1024-byte swizzle phase.
//            (A bytes per tile: 3 slabs instead of 9 tiles per channel chunk -> 2.4x less L2->SM traffic.)
//   B      : BRES = 0: ring of kBStages weight tiles {64 k, BN}, one per tap and channel chunk;
//            BRES = number of 64-channel input chunks whose weights are RESIDENT: when
//            taps * Cin * BN * 2 bytes fit (Cin = 64 at N <= 128, the phase weights of the Cin = 128
//            upsample-fused layer) they are loaded once per CTA and never re-fetched.
//   store  : 2 x 16 KiB staging tiles (one per epilogue group) for the TMA store of the epilogue.
constexpr int kSlabRows = kTileH + 2;
constexpr int kASlabBytes = kSlabRows * kTileW * 128;  // 20480

template <int BN, int BRES, int CG, bool UPS = false>
struct UmmaCfg {
  // CG = 2 (CTA pair, tcgen05 cta_group::2): the pair computes M = 256 pixels x BN channels per MMA;
  // each CTA stages the A slab of its own 128-pixel tile and HALF of the weight tile (BN/2 rows),
  // so the per-SM shared-memory traffic of the B operand (TMA writes and tensor-core reads) halves.
  static constexpr int kBRows = BN / CG;
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kBStride = (kBBytes + 1023) / 1024 * 1024;
  static constexpr int kTaps = UPS ? 4 : 9;  // UPS: 2x2 phase convolution (see EPI_UPS)
  static constexpr int kAStages =
      CG == 2 ? (BRES ? (BN >= 128 && !UPS ? 5 : 6) : (BN >= 256 ? 4 : 5))
              : (BRES ? (BN >= 128 ? 3 : (UPS ? 6 : 5)) : (BN >= 256 ? 3 : 4));
  // BRES = 0: the weight tiles of the kTR filter rows of one (chunk, filter column) step travel
  // as ONE group -- one full/empty barrier pair, one wait per step in the producer and in the MMA
  // warp (a wait + elect + issue round per single tile costs ~300 cycles of serial scalar code in
  // each of those warps, more than the 256 cycles of math a tile feeds at N = 128).
  static constexpr int kBGroup = BRES ? 1 : (UPS ? 2 : 3);
  static constexpr int kBStagesRaw = BRES ? kTaps * BRES /* resident: all taps x BRES chunks */
                                     : CG == 2 ? (BN >= 256 ? 6 : 9)
                                               : (BN >= 256 ? 4 : (BN >= 128 ? 6 : 9));
  static constexpr int kBStages = kBStagesRaw / kBGroup * kBGroup;
  static constexpr int kStoreStageBytes = 2 * kBlockM * 128;
  static constexpr int kBiasBytes = 2048;  // up to 512 fp32 biases
  static constexpr int kAOff = 0;
  static constexpr int kBOff = kAStages * kASlabBytes;
  static constexpr int kStoreOff = kBOff + kBStages * kBStride;
  static constexpr int kBiasOff = kStoreOff + kStoreStageBytes;
  static constexpr int kBarOff = kBiasOff + kBiasBytes;
  static constexpr int kNumBars = 2 * kAStages + 2 * kBStages + 4 + 1;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kBarOff + 8 * kNumBars + 16 + 1024 /*align slack*/;
  static constexpr bool kFits = kSmemBytes <= 232448;  // 227 KiB; launch_cfg refuses plans that do not fit
  static_assert(BN == 64 || BN == 128 || BN == 256, "tile widths of the main kernel");
};

// work unit -> (N tile, pixel tile of CTA `rank` of the pair).  A pair takes two consecutive pixel
// tiles; when the number of pixel tiles is odd the last pair's second tile is a dummy at image
// index n = N: its TMA loads are out of bounds (zero fill) and its stores are clipped away.
// UPS: the four output phases (a, b) of one pixel tile are consecutive units, so the CTAs that run
// them concurrently share the tile's input slabs in L2.
// PSW (per-sample weights): pixel tiles are numbered per image with the per-image count padded to a
// multiple of CG, so the two tiles of a pair always belong to the same image `nw` (they share one
// weight tile); the padding tiles are dummies like above.
template <int CG, bool UPS = false, bool PSW = false, typename P>
__device__ __forceinline__ TileCoord decode_tile(const P& p, int unit, int rank) {
  TileCoord t;
  t.nt = unit % p.n_tiles;
  int u = unit / p.n_tiles;
  t.ph = 0;
  if (UPS) t.ph = u & 3, u >>= 2;
  int m = u * CG + rank;
  if (PSW) {
    t.nw = m / p.m_tiles_img;
    m -= t.nw * p.m_tiles_img;
    if (m >= p.tiles_x * p.tiles_y) {
      t.x0 = 0, t.y0 = 0, t.n = p.N;
      return t;
    }
    t.x0 = (m % p.tiles_x) * kTileW;
    t.y0 = (m / p.tiles_x) * kTileH;
    t.n = t.nw;
    return t;
  }
  t.nw = 0;
  if (CG == 2 && m >= p.m_tiles) {
    t.x0 = 0, t.y0 = 0, t.n = p.N;
    return t;
  }
  t.x0 = (m % p.tiles_x) * kTileW;
  m /= p.tiles_x;
  t.y0 = (m % p.tiles_y) * kTileH;
  t.n = m / p.tiles_y;
  return t;
}

template <typename T16, int BN, int EPI, int BRES, int CG, bool PSW = false>
__global__ void __launch_bounds__(kThreadsUmma, 1)
    conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a,
                     const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ OutMaps tmap_out, ConvParams<T16> p) {
  // UPS (EPI_UPS): `p.H x p.W` is the low-resolution input S (replicate halo); output phase (a, b)
  // holds pixels (2y + a, 2x + b) = sum over the 2x2 source window S[y + a - 1 + dy][x + b - 1 + dx]
  // with the 3x3 taps that fall on the same source pixel pre-summed (api.cu pack_layer).
  constexpr bool UPS = (EPI == EPI_UPS);
  constexpr int kTR = UPS ? 2 : 3, kTS = UPS ? 2 : 3;
  using Cfg = UmmaCfg<BN, BRES, CG, UPS>;
  static_assert(!PSW || (BRES == 0 && !UPS), "per-sample weights are streamed");
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte aligned bases (the dynamic shared window starts at the same
  // offset in both CTAs of a pair, so the carve-up below is identical in both)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic alias of smem_base
  const uint32_t store_base = smem_base + Cfg::kStoreOff;
  const uint32_t bar_base = smem_base + Cfg::kBarOff;
  float* s_bias = reinterpret_cast<float*>(smem_gen + Cfg::kBiasOff);
  auto a_smem = [&](int s) { return smem_base + Cfg::kAOff + s * kASlabBytes; };
  auto b_smem = [&](int s) { return smem_base + Cfg::kBOff + s * Cfg::kBStride; };
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (Cfg::kAStages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * Cfg::kAStages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * Cfg::kAStages + Cfg::kBStages + s); };
  constexpr int kBar2 = 2 * Cfg::kAStages + 2 * Cfg::kBStages;
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (kBar2 + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (kBar2 + 2 + s); };
  const uint32_t bres_bar = bar_base + 8u * (kBar2 + 4);  // resident weights landed
  const uint32_t tmem_slot = bar_base + 8u * Cfg::kNumBars;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = p.Cin / kBlockK;
  // CTA pair: "full" and "accumulator drained" barriers are the LEADER's (rank 0); the peer's TMA
  // bytes and epilogue arrivals are credited to them through their shared::cluster address.  The
  // "empty" / "accumulator ready" barriers exist in both CTAs and are signalled by multicast commits.
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  auto lead = [&](uint32_t bar) { return CG == 2 ? mapa_rank(bar, 0) : bar; };
  // Work units of this CTA (pair): unit_id, unit_id + unit_cnt, ... -- the CTAs that run concurrently work on
  // neighbouring tiles of the same image(s), which keeps the halo rows / columns they share, the weights and
  // (dec1) the per-image weights in L2.  (Contiguous ranges per CTA were measured slower: +15 % DRAM reads on
  // the N = 256 layers, 3x on dec1's per-image weights.)
  const int unit_id = CG == 2 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int unit_cnt = CG == 2 ? (int)ncluster_x() : (int)gridDim.x;
  const int unit_begin = unit_id, unit_end = p.total_tiles, unit_inc = unit_cnt;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_out.m[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kAStages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < Cfg::kBStages; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full_bar(s), 1);
      mbar_init(tmem_empty_bar(s), 4 * CG);  // one arrive per epilogue warp (of both CTAs)
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg<CG, Cfg::kTmemCols>(tmem_slot);
  if (!PSW)
    for (int i = threadIdx.x; i < p.CoutPad; i += kThreadsUmma) s_bias[i] = p.bias[i];
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base =
      *reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kBarOff + 8 * Cfg::kNumBars);
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();  // (the producer warp first issues the constant resident weights)

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, one lane issues) ==============
    const int b_row0 = (int)cta_rank * Cfg::kBRows;  // this CTA's half of the N tile
    if (BRES) {
      // all weight tiles of this (Cin == 64) layer, once.  UPS: the grid is a multiple of 4 (or has
      // one unit per CTA), so every unit of this CTA has the same phase and only its taps are kept.
      if (elect_one()) {
        if (leader) mbar_expect_tx(bres_bar, CG * Cfg::kBStages * Cfg::kBBytes);
        const uint32_t bar = lead(bres_bar);
        const int row_ph = UPS ? (unit_id & 3) * p.CoutPad : 0;
        for (int kc = 0; kc < BRES; ++kc)
          for (int tap = 0; tap < Cfg::kTaps; ++tap)
            tma_load_2d_cg<CG>(b_smem(kc * Cfg::kTaps + tap), &tmap_b, bar, tap * p.Cin + kc * kBlockK,
                               row_ph + b_row0);
      }
      __syncwarp();
    }
    pdl_wait();
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int unit = unit_begin; unit < unit_end; unit += unit_inc) {
      const TileCoord t = decode_tile<CG, UPS, PSW>(p, unit, (int)cta_rank);
      const int xs0 = t.x0 + (UPS ? (t.ph & 1) : 0);
      const int b_row = (UPS ? t.ph * p.CoutPad : 0) + t.nt * BN + b_row0 + (PSW ? t.nw * p.w_rows_per_n : 0);
      for (int kc = 0; kc < kchunks; ++kc) {
        for (int s = 0; s < kTS; ++s) {
          MBAR_WAIT_RELAXED(a_empty(as), aph ^ 1, 100 + as);
          if (CCST_ABLATE_BITS(p) & 4) {
            if (elect_one()) {
              if (leader) mbar_arrive(a_full(as));
            }
          } else if (elect_one()) {
            if (leader) mbar_expect_tx(a_full(as), CG * kASlabBytes);
            // interior pixel (y, x) is stored at (y+1, x+1): the slab for filter column s starts at
            // padded (y0, x0 + s) and spans the rows needed by r = 0..2 (UPS: source column
            // x + b - 1 + s, rows a + r of the slab).
            tma_load_4d_cg<CG>(a_smem(as), &tmap_a, lead(a_full(as)), kc * kBlockK, xs0 + s, t.y0, t.n);
          }
          __syncwarp();
          if (++as == Cfg::kAStages) as = 0, aph ^= 1;
          if (!BRES) {
            // the kTR weight tiles of this step: one barrier (that of the group's first slot)
            MBAR_WAIT_RELAXED(b_empty(bs), bph ^ 1, 150 + bs);
            if (elect_one()) {
              if (leader) mbar_expect_tx(b_full(bs), CG * kTR * Cfg::kBBytes);
              const uint32_t bar = lead(b_full(bs));
#pragma unroll
              for (int r = 0; r < kTR; ++r)
                tma_load_2d_cg<CG>(b_smem(bs + r), &tmap_b, bar, (r * kTS + s) * p.Cin + kc * kBlockK, b_row);
            }
            __syncwarp();
            if ((bs += kTR) == Cfg::kBStages) bs = 0, bph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; whole warp converged, one lane issues) =====
    if (leader) {
      constexpr uint32_t idesc = make_idesc<T16, BN, CG>();
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      if (BRES) {
        mbar_wait(bres_bar, 0, 250);
        tc_fence_after();
      }
      for (int unit = unit_begin; unit < unit_end; unit += unit_inc, ++it) {
        const int acs = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        if (CG == 2) mbar_wait_cluster(tmem_empty_bar(acs), aphase ^ 1, 200 + acs);
        else mbar_wait(tmem_empty_bar(acs), aphase ^ 1, 200 + acs);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acs * BN);
        // UPS: row phase a of this unit shifts the slab rows of the two taps to a + r
        const int row_shift = UPS ? (((unit / p.n_tiles) & 3) >> 1) : 0;
        for (int kc = 0; kc < kchunks; ++kc) {
#pragma unroll
          for (int s = 0; s < kTS; ++s) {
            // One elected-lane region per (chunk, filter column): all kTR filter rows x 4 K steps are
            // issued back to back.  (Electing per tap cost ~40 scalar/uniform instructions around
            // every 4 MMAs -- ~200 issue cycles against 128 cycles of math at N = 64 -- which made
            // the issuing warp, not the tensor pipe, the bound of the 64-channel layers.)
            mbar_wait(a_full(as), aph, 300 + as);
            const int bs0 = bs;
            if (!BRES) {
              mbar_wait(b_full(bs), bph, 350 + bs);
              if ((bs += kTR) == Cfg::kBStages) bs = 0, bph ^= 1;
            }
            tc_fence_after();
            if (elect_one()) {
              // tap (r, s): slab shifted by r rows (16 px * 128 B, swizzle-phase neutral)
              const uint64_t adesc0 = make_kmajor_sw128_desc(a_smem(as) + (uint32_t)(row_shift * kTileW) * 128u);
#pragma unroll
              for (int r = 0; r < kTR; ++r) {
                const uint64_t adesc = adesc0 + (uint64_t)(r * (kTileW * 128 >> 4));
                const uint64_t bdesc = make_kmajor_sw128_desc(BRES ? b_smem(kc * Cfg::kTaps + r * kTS + s) : b_smem(bs0 + r));
                if (!(CCST_ABLATE_BITS(p) & 2)) {
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k) {
                    // +16 elements (32 bytes) along K inside the swizzle atom = +2 in the start field
                    umma_f16_cg<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc,
                                    (kc | s | r | k) ? 1u : 0u);
                  }
                }
              }
              if (!BRES) umma_commit_cg<CG>(b_empty(bs0));  // frees the weight tiles when these MMAs retire
              umma_commit_cg<CG>(a_empty(as));              // ... and the slab
              if (kc == kchunks - 1 && s == kTS - 1) umma_commit_cg<CG>(tmem_full_bar(acs));
            }
            __syncwarp();
            if (++as == Cfg::kAStages) as = 0, aph ^= 1;
          }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: two groups of 4 warps, group g drains accumulator stage g
    // (tiles it = g, g+2, ...), so each group has two tile-times to finish one tile ==========
    // (Four groups -- one 64-column chunk of every tile per group at BN = 256, (tile parity, chunk) at BN = 128,
    // 640 threads, one staging tile per group and one slab stage less -- were measured on every layer of the step,
    // interleaved with this form on one box: slower everywhere, 5.11 -> 5.27 ms per step, dec6 0.19 -> 0.23 ms
    // (profiles/r03l_ab_main_kernel_4_epilogue_groups.txt).  Here the MMA side is the bound and the epilogue has
    // slack; on dec8, whose epilogue was the bound, the same change paid: conv_ups4.cuh.)
    const int grp = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;           // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;    // accumulator row = pixel inside the tile
    const int py = row / kTileW, px = row % kTileW;
    // warp 4 owns the bulk-store async groups: all its lanes execute the waits (a no-op for lanes
    // without groups), one elected lane -- always the same one -- issues and commits the stores
    const bool issuer_warp = (quad == 0);
    const uint32_t sbuf = store_base + grp * kStoreBytes;
    SatTracker<T16> sat;
    for (int it = grp;; it += 2) {
      const long long unit_ll = (long long)unit_begin + (long long)it * unit_inc;
      if (unit_ll >= unit_end) break;
      const TileCoord t = decode_tile<CG, UPS, PSW>(p, (int)unit_ll, (int)cta_rank);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int y = t.y0 + py, x = t.x0 + px;
      const bool valid = (y < p.H) && (x < p.W) && (CG == 1 || t.n < p.N);
      // per-sample bias (AdaIN folded into this conv): read straight from global, one broadcast
      // address per warp (a dummy tile reads image 0's)
      const float* bias_n = PSW ? p.bias + (size_t)(t.n < p.N ? t.n : 0) * p.bias_per_n : nullptr;
      MBAR_WAIT_RELAXED(tmem_full_bar(as), aphase, 400 + as);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
      if (CCST_ABLATE_BITS(p) & 1) {
        // measurement only: hand the accumulator back untouched
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(lead(tmem_empty_bar(as)));
          else mbar_arrive(tmem_empty_bar(as));
        }
        continue;
      }
#pragma unroll 1
      for (int ch = 0; ch < BN / 64; ++ch) {
        uint32_t r[64];
        {
          uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
          uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
          tmem_ld32(taddr + ch * 64, r0);
          tmem_ld32(taddr + ch * 64 + 32, r1);
        }
        tmem_ld_wait();
        if (ch == BN / 64 - 1) {
          // the accumulator stage is in registers: hand it back to the MMA warp before the
          // pack / stage / store work of this last chunk
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(lead(tmem_empty_bar(as)));
            else mbar_arrive(tmem_empty_bar(as));
          }
        }
        const int co = t.nt * BN + ch * 64;
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float b0, b1;
          if (PSW) {
            const float2 b2 = __ldg(reinterpret_cast<const float2*>(bias_n + co + 2 * j));
            b0 = b2.x, b1 = b2.y;
          } else {
            b0 = s_bias[co + 2 * j], b1 = s_bias[co + 2 * j + 1];
          }
          const float v0 = __uint_as_float(r[2 * j]) + b0;
          const float v1 = __uint_as_float(r[2 * j + 1]) + b1;
          pk[j] = p.relu ? pack16x2_relu<T16>(v0, v1) : pack16x2<T16>(v0, v1);
          // (pooling below) out-of-image pixels contribute 0, the identity for post-ReLU values
          if (EPI == EPI_ACT_POOL) pk[j] = valid ? pk[j] : 0u;
        }
        if (EPI == EPI_ACT_POOL) {
          // 2x2 window = lanes {l, l^1, l^16, l^17}, pooled on the packed pairs (rounding and ReLU are
          // monotonic, so max commutes with them: half the shuffles of fp32 pooling).  The shuffles are
          // issued in two blocks of 32: written word by word, each word's shuffle -> max -> shuffle -> max
          // chain would have to finish before the next word's first shuffle (convergent operations stay
          // in program order).
          uint32_t t[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = __shfl_xor_sync(0xffffffffu, pk[j], 1);
#pragma unroll
          for (int j = 0; j < 32; ++j) pk[j] = max16x2<T16>(pk[j], t[j]);
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = __shfl_xor_sync(0xffffffffu, pk[j], 16);
#pragma unroll
          for (int j = 0; j < 32; ++j) pk[j] = max16x2<T16>(pk[j], t[j]);
        }
        sat.track_block(pk, p.relu != 0);
        // the staging buffer about to be rewritten must have been read out by its TMA store (waited
        // for only now, so that the bias / ReLU / pack work above overlaps that read-out)
        if (issuer_warp) bulk_wait_read<0>();
        epi_barrier(grp);
        // stage the row (128 bytes = 8 chunks) with the 128-byte swizzle the TMA store expects
        int srow = row;
        bool writer = true;
        if (EPI == EPI_ACT_POOL) {
          writer = !(lane & 1) && lane < 16;  // anchor of a 2x2 window
          srow = (py >> 1) * (kTileW / 2) + (px >> 1);
        }
        if (writer) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t dst = sbuf + srow * 128 + ((j ^ (srow & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]),
                         "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                         : "memory");
          }
        }
        if (valid) {
          if (EPI == EPI_ACT || EPI == EPI_ACT_STATS) {
            store_aliases(p.out, t.n, y, x, co, pk, p.halo_edge);
          } else if (EPI == EPI_UPS) {
            store_aliases(p.out, t.n, 2 * y + (t.ph >> 1), 2 * x + (t.ph & 1), co, pk);
          } else if (EPI == EPI_ACT_UP2) {
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
              for (int b = 0; b < 2; ++b) store_aliases(p.out, t.n, 2 * y + a, 2 * x + b, co, pk);
          } else if (EPI == EPI_ACT_POOL) {
            if (writer) store_aliases(p.out, t.n, y >> 1, x >> 1, co, pk);
          }
        }
        fence_async_smem();
        epi_barrier(grp);
        if (issuer_warp && elect_one()) {
          // coordinates are interior pixels; TMA clips the box at the image border (and drops the
          // dummy tile of an odd pair entirely: n = N is out of bounds)
          if (EPI == EPI_ACT_POOL) {
            tma_store_4d(&tmap_out.m[0], sbuf, co, t.x0 >> 1, t.y0 >> 1, t.n);
          } else if (EPI == EPI_UPS) {
            tma_store_4d(&tmap_out.m[t.ph], sbuf, co, t.x0, t.y0, t.n);
          } else {
            tma_store_4d(&tmap_out.m[0], sbuf, co, t.x0, t.y0, t.n);
            if (EPI == EPI_ACT_UP2) {
              tma_store_4d(&tmap_out.m[1], sbuf, co, t.x0, t.y0, t.n);
              tma_store_4d(&tmap_out.m[2], sbuf, co, t.x0, t.y0, t.n);
              tma_store_4d(&tmap_out.m[3], sbuf, co, t.x0, t.y0, t.n);
            }
          }
          bulk_commit();
        }
        if (EPI == EPI_ACT_STATS && (CG == 1 || t.n < p.N)) {
          // statistics of the STORED (rounded) values -- the values the consumer of the feature map
          // reads -- taken from the staged tile: thread = (channel pair = lane, quarter of the tile =
          // two tile rows = warp); exact two-pass over the quarter's valid pixels, conflict-free (a
          // warp reads one 128-byte staged row at a time)
          const int wv = min(kTileW, p.W - t.x0);
          const int rows = max(0, min(2, p.H - (t.y0 + 2 * quad)));
          const int cnt = rows * wv;
          float s0 = 0.f, s1 = 0.f;
          for (int rr = 0; rr < rows; ++rr)
            for (int xx = 0; xx < wv; ++xx) {
              const int r2 = (2 * quad + rr) * kTileW + xx;
              const float2 f = unpack16x2<T16>(lds_u32(sbuf + r2 * 128 + ((((lane >> 2) ^ (r2 & 7))) << 4) + ((lane & 3) << 2)));
              s0 += f.x, s1 += f.y;
            }
          const float inv = cnt > 0 ? 1.f / (float)cnt : 0.f;
          const float m0 = s0 * inv, m1 = s1 * inv;
          float q0 = 0.f, q1 = 0.f;
          for (int rr = 0; rr < rows; ++rr)
            for (int xx = 0; xx < wv; ++xx) {
              const int r2 = (2 * quad + rr) * kTileW + xx;
              const float2 f = unpack16x2<T16>(lds_u32(sbuf + r2 * 128 + ((((lane >> 2) ^ (r2 & 7))) << 4) + ((lane & 3) << 2)));
              const float d0 = f.x - m0, d1 = f.y - m1;
              q0 = fmaf(d0, d0, q0), q1 = fmaf(d1, d1, q1);
            }
          const size_t tile = ((size_t)t.n * p.tiles_y + t.y0 / kTileH) * p.tiles_x + t.x0 / kTileW;
          float4* dst = reinterpret_cast<float4*>(p.tile_stats + (tile * 4 + quad) * p.Cout + co + 2 * lane);
          *dst = make_float4(m0, q0, m1, q1);
        }
      }
    }
    if (issuer_warp) bulk_wait_all();
    sat.flush(p.sat_count);
  }

  __syncwarp();
  tc_fence_before();
  // pair: the peer's shared memory / TMEM are operands of the leader's MMAs until the very end
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc_cg<CG, Cfg::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------ host side
template <typename T16, int BN, int EPI, int BRES, int CG, bool PSW = false>
int launch_cfg(const CUtensorMap& ma, const T16* wk, ConvParams<T16> p, cudaStream_t st) {
  constexpr bool UPS = (EPI == EPI_UPS);
  using Cfg = UmmaCfg<BN, BRES, CG, UPS>;
  if constexpr (!Cfg::kFits) {
    set_error("conv_umma: configuration BN=%d resident=%d pair=%d does not fit shared memory", BN, BRES, CG);
    return CCST_EINVAL;
  } else {
    int64_t m_tiles = p.m_tiles;
    if (PSW) {
      // tiles numbered per image, per-image count padded to a multiple of CG (see decode_tile)
      p.m_tiles_img = (p.tiles_x * p.tiles_y + CG - 1) / CG * CG;
      m_tiles = (int64_t)p.N * p.m_tiles_img;
    }
    CUtensorMap mb;
    const int w_rows = PSW ? p.N * p.w_rows_per_n : (UPS ? 4 : 1) * p.CoutPad;
    if (int e = make_weight_map(&mb, wk, Cfg::kTaps * p.Cin, w_rows, Cfg::kBRows)) return e;
    OutMaps mo;
    memset(&mo, 0, sizeof(mo));
    if (EPI == EPI_ACT || EPI == EPI_ACT_STATS) {
      if (int e = make_out_map(&mo.m[0], p.out, 0, 0, 1, 1, kTileW, kTileH)) return e;
    } else if (EPI == EPI_ACT_POOL) {
      if (int e = make_out_map(&mo.m[0], p.out, 0, 0, 1, 1, kTileW / 2, kTileH / 2)) return e;
    } else if (EPI == EPI_ACT_UP2 || EPI == EPI_UPS) {
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
          if (int e = make_out_map(&mo.m[a * 2 + b], p.out, a, b, 2, 2, kTileW, kTileH)) return e;
    }
    auto kernel = conv_umma_kernel<T16, BN, EPI, BRES, CG, PSW>;
    CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(kernel), Cfg::kSmemBytes));
    const int64_t units = (m_tiles + CG - 1) / CG * p.n_tiles * (UPS ? 4 : 1);
    CCST_CHECK_ARG(units < (1ll << 31), "conv_umma: too many tiles");
    p.total_tiles = (int)units;
    int slots = sm_count() / CG;  // persistent: one CTA (or CTA pair) per SM (pair)
    if (UPS && BRES) slots &= ~3;  // resident weights of ONE phase per CTA: unit stride % 4 == 0
    const int grid = (int)(units < slots ? units : slots) * CG;
    CCST_CUDA(launch_conv(kernel, grid, kThreadsUmma, Cfg::kSmemBytes, st, CG, ma, mb, mo, p));
    CCST_LAUNCHED();
    return CCST_OK;
  }
}

// Dispatch over the epilogues a tile width is used with (keeps the instantiation count down):
//   N = 256 (CTA pairs): plain, +pool, +tile statistics, fused-upsample input, upsample store,
//            per-sample weights (dec1 with AdaIN folded in);
//   N = 128 (CTA pairs): plain, +pool, upsample store, fused-upsample input (2 resident chunks),
//            resident Cin = 64 weights (conv2_1);
//   N = 64  (single CTA): only the fused-upsample input with Cin != 64 (Cin = 64 runs conv_ups4, every other
//            64-channel layer the s-merged kernel).
// CTA pairs (cta_group::2) for N >= 128: measured on B200, batch 32 @512^2: N=256 layers 1.50 -> 1.67
// PFLOP/s, N=128 layers 1.15 -> 1.27.
template <typename T16>
int launch_main(const CUtensorMap& ma, const T16* wk, const ConvParams<T16>& p, int BN, int epi, cudaStream_t st) {
  const bool psw = p.w_rows_per_n != 0;
  if (BN == 256) {
    if (psw) {
      if (epi == EPI_ACT) return launch_cfg<T16, 256, EPI_ACT, 0, 2, true>(ma, wk, p, st);
    } else {
      switch (epi) {
        case EPI_ACT: return launch_cfg<T16, 256, EPI_ACT, 0, 2>(ma, wk, p, st);
        case EPI_ACT_UP2: return launch_cfg<T16, 256, EPI_ACT_UP2, 0, 2>(ma, wk, p, st);
        case EPI_ACT_POOL: return launch_cfg<T16, 256, EPI_ACT_POOL, 0, 2>(ma, wk, p, st);
        case EPI_UPS: return launch_cfg<T16, 256, EPI_UPS, 0, 2>(ma, wk, p, st);
        case EPI_ACT_STATS: return launch_cfg<T16, 256, EPI_ACT_STATS, 0, 2>(ma, wk, p, st);
        default: break;
      }
    }
  } else if (BN == 128 && !psw) {
    switch (epi) {
      case EPI_ACT:
        // 64 -> 128 (conv2_1): the 9 weight tiles stay resident
        return p.Cin == kBlockK ? launch_cfg<T16, 128, EPI_ACT, 1, 2>(ma, wk, p, st)
                                : launch_cfg<T16, 128, EPI_ACT, 0, 2>(ma, wk, p, st);
      case EPI_ACT_UP2: return launch_cfg<T16, 128, EPI_ACT_UP2, 0, 2>(ma, wk, p, st);
      case EPI_ACT_POOL: return launch_cfg<T16, 128, EPI_ACT_POOL, 0, 2>(ma, wk, p, st);
      case EPI_UPS:
        // the 8 phase tiles (4 taps x 2 chunks) of the Cin = 128 layer stay resident
        return p.Cin == 2 * kBlockK ? launch_cfg<T16, 128, EPI_UPS, 2, 2>(ma, wk, p, st)
                                    : launch_cfg<T16, 128, EPI_UPS, 0, 2>(ma, wk, p, st);
      default: break;
    }
  } else if (BN == 64 && !psw) {
    if (epi == EPI_UPS) return launch_cfg<T16, 64, EPI_UPS, 0, 1>(ma, wk, p, st);
  }
  set_error("conv_umma: no kernel for N tile %d, epilogue %d%s", BN, epi, psw ? ", per-sample weights" : "");
  return CCST_EINVAL;
}

}  // namespace
}  // namespace ccst
