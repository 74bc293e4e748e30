// 64-output-channel layers (conv1_2, dec7; dec8 when its upsample is not fused): s-merged tcgen05 kernel.
#pragma once
#include "umma_common.cuh"

namespace ccst {
namespace {

// =====================================================================================
// 64-output-channel layers (conv1_2, dec7, dec8), "s-merged" variant.
// With N = 64 every tcgen05.mma reads 4 KiB of A and 2 KiB of B from shared memory for 32 cycles of
// math: the tap-by-tap kernel above is bound by shared-memory bandwidth at ~45 % tensor-pipe use.
// Here the three filter COLUMNS move into the N dimension:
//   P[(jy, jx), (s, co)] = sum_{r, c} X[(jy + r, jx), c] * W[co][c][r][s]        N = 192, K = 3 * Cin
//   out[(jy, ox), co]    = P[(jy, ox), (0, co)] + P[(jy, ox + 1), (1, co)] + P[(jy, ox + 2), (2, co)]
// so one slab {64 ch, 16 px, 10 rows} per channel chunk feeds three MMAs of N = 192 (A is read 3x per
// chunk instead of 9x) and the shifted sum over s is two warp shuffles per value in the epilogue
// (TMEM lane = slab pixel; jx neighbours are adjacent lanes).  A 16-pixel-wide slab yields 14 output
// columns, so tiles step by 14 pixels in x (12.5 % of the MMA rows are halo).
// =====================================================================================
constexpr int kSmOutW = kTileW - 2;                 // 14 output columns per tile
constexpr int kSmN = 192;
constexpr int kSmStoreBytes = kTileH * kSmOutW * 128;  // 14336 = 14 x 1024

// Two 4-warp epilogue groups (tile i is drained by group i % 2 from accumulator stage i % 2).  A third
// group was measured to buy nothing: the epilogue alone gets faster (0.50 -> 0.41 ms on conv1_2) but
// not the kernel (0.59 ms), see DESIGN.md.
template <bool BRES, int CG>
struct SmergeCfg {
  static constexpr int kASlabBytes = (kTileH + 2) * kTileW * 128;  // 20480
  static constexpr int kBRows = kSmN / CG;
  static constexpr int kBBytes = kBRows * kBlockK * 2;             // 24576 / CG
  // streamed weights: the 3 filter-row tiles of a chunk travel as one group (one barrier pair, one
  // wait per chunk in the producer and the MMA warp); two groups in flight
  static constexpr int kAStages = BRES ? 5 : (CG == 2 ? 4 : 2);
  static constexpr int kBStages = BRES ? 3 : 6;                    // resident: 3 filter rows x (Cin == 64)
  static constexpr int kAOff = 0;
  static constexpr int kBOff = kAStages * kASlabBytes;
  static constexpr int kStoreOff = kBOff + kBStages * kBBytes;
  static constexpr int kThreads = kThreadsUmma;                    // 4 control warps + 2 epilogue groups
  static constexpr int kFullBars = 2;
  static constexpr int kBiasOff = kStoreOff + 2 * kSmStoreBytes;
  static constexpr int kBarOff = kBiasOff + 256;
  static constexpr int kNumBars = 2 * kAStages + 2 * kBStages + kFullBars + 2 + 1;
  static constexpr int kTmemCols = 512;                            // 2 stages x 192 columns
  static constexpr int kSmemBytes = kBarOff + 8 * kNumBars + 16 + 1024;
  static_assert(kSmemBytes <= 232448, "shared memory plan exceeds 227 KiB");
  static_assert(kBBytes % 1024 == 0, "B tiles must keep the swizzle phase");
};

template <int CG, typename P>
__device__ __forceinline__ TileCoord decode_tile_sm(const P& p, int unit, int rank) {
  TileCoord t;
  t.nt = 0;
  int m = unit * CG + rank;
  if (CG == 2 && m >= p.m_tiles) {
    t.x0 = 0, t.y0 = 0, t.n = p.N;
    return t;
  }
  t.x0 = (m % p.tiles_x) * kSmOutW;
  m /= p.tiles_x;
  t.y0 = (m % p.tiles_y) * kTileH;
  t.n = m / p.tiles_y;
  return t;
}

template <typename T16, int EPI, bool BRES, int CG>
__global__ void __launch_bounds__(kThreadsUmma, 1)
    conv_smerge_kernel(const __grid_constant__ CUtensorMap tmap_a,
                       const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ OutMaps tmap_out, ConvParams<T16> p) {
  using Cfg = SmergeCfg<BRES, CG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t store_base = smem_base + Cfg::kStoreOff;
  const uint32_t bar_base = smem_base + Cfg::kBarOff;
  float* s_bias = reinterpret_cast<float*>(smem_gen + Cfg::kBiasOff);
  auto a_smem = [&](int s) { return smem_base + Cfg::kAOff + s * Cfg::kASlabBytes; };
  auto b_smem = [&](int s) { return smem_base + Cfg::kBOff + s * Cfg::kBBytes; };
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (Cfg::kAStages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * Cfg::kAStages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * Cfg::kAStages + Cfg::kBStages + s); };
  constexpr int kBar2 = 2 * Cfg::kAStages + 2 * Cfg::kBStages;
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (kBar2 + s); };  // s = tile counter % kFullBars
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (kBar2 + Cfg::kFullBars + s); };
  const uint32_t bres_bar = bar_base + 8u * (kBar2 + Cfg::kFullBars + 2);
  const uint32_t tmem_slot = bar_base + 8u * Cfg::kNumBars;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = p.Cin / kBlockK;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  auto lead = [&](uint32_t bar) { return CG == 2 ? mapa_rank(bar, 0) : bar; };
  const int unit0 = CG == 2 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int unit_step = CG == 2 ? (int)ncluster_x() : (int)gridDim.x;

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
    for (int s = 0; s < Cfg::kFullBars; ++s) mbar_init(tmem_full_bar(s), 1);
    for (int s = 0; s < 2; ++s) mbar_init(tmem_empty_bar(s), 4 * CG);
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg<CG, Cfg::kTmemCols>(tmem_slot);
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base =
      *reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kBarOff + 8 * Cfg::kNumBars);
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer
    const int b_row0 = (int)cta_rank * Cfg::kBRows;
    if (BRES) {
      if (elect_one()) {
        if (leader) mbar_expect_tx(bres_bar, CG * 3 * Cfg::kBBytes);
        const uint32_t bar = lead(bres_bar);
        for (int r = 0; r < 3; ++r) tma_load_2d_cg<CG>(b_smem(r), &tmap_b, bar, r * p.Cin, b_row0);
      }
      __syncwarp();
    }
    pdl_wait();
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int unit = unit0; unit < p.total_tiles; unit += unit_step) {
      const TileCoord t = decode_tile_sm<CG>(p, unit, (int)cta_rank);
      for (int kc = 0; kc < kchunks; ++kc) {
        MBAR_WAIT_RELAXED(a_empty(as), aph ^ 1, 500 + as);
        if (elect_one()) {
          if (leader) mbar_expect_tx(a_full(as), CG * Cfg::kASlabBytes);
          // slab column jx <-> interior x0 - 1 + jx <-> padded x0 + jx; rows y0 - 1 .. y0 + 8
          tma_load_4d_cg<CG>(a_smem(as), &tmap_a, lead(a_full(as)), kc * kBlockK, t.x0, t.y0, t.n);
        }
        __syncwarp();
        if (++as == Cfg::kAStages) as = 0, aph ^= 1;
        if (!BRES) {
          MBAR_WAIT_RELAXED(b_empty(bs), bph ^ 1, 550 + bs);
          if (elect_one()) {
            if (leader) mbar_expect_tx(b_full(bs), CG * 3 * Cfg::kBBytes);
            const uint32_t bar = lead(b_full(bs));
#pragma unroll
            for (int r = 0; r < 3; ++r)
              tma_load_2d_cg<CG>(b_smem(bs + r), &tmap_b, bar, r * p.Cin + kc * kBlockK, b_row0);
          }
          __syncwarp();
          if ((bs += 3) == Cfg::kBStages) bs = 0, bph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: per channel chunk 3 filter rows x 4 K steps, N = 192
    if (leader) {
      constexpr uint32_t idesc = make_idesc<T16, kSmN, CG>();
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      if (BRES) {
        mbar_wait(bres_bar, 0, 560);
        tc_fence_after();
      }
      for (int unit = unit0; unit < p.total_tiles; unit += unit_step, ++it) {
        const int acs = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        if (CG == 2) mbar_wait_cluster(tmem_empty_bar(acs), aphase ^ 1, 570 + acs);
        else mbar_wait(tmem_empty_bar(acs), aphase ^ 1, 570 + acs);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acs * kSmN);
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(a_full(as), aph, 580 + as);
          const int bs0 = bs;
          if (!BRES) {
            mbar_wait(b_full(bs), bph, 590 + bs);
            if ((bs += 3) == Cfg::kBStages) bs = 0, bph ^= 1;
          }
          tc_fence_after();
          if (elect_one()) {
            // one issue region per chunk: 3 filter rows x 4 K steps of N = 192
            const uint64_t adesc0 = make_kmajor_sw128_desc(a_smem(as));
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const uint64_t adesc = adesc0 + (uint64_t)(r * (kTileW * 128 >> 4));
              const uint64_t bdesc = make_kmajor_sw128_desc(BRES ? b_smem(r) : b_smem(bs0 + r));
              if (!(CCST_ABLATE_BITS(p) & 2)) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_f16_cg<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | r | k) ? 1u : 0u);
              }
            }
            if (!BRES) umma_commit_cg<CG>(b_empty(bs0));
            umma_commit_cg<CG>(a_empty(as));
            if (kc == kchunks - 1) umma_commit_cg<CG>(tmem_full_bar(it % Cfg::kFullBars));
          }
          __syncwarp();
          if (++as == Cfg::kAStages) as = 0, aph ^= 1;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: two groups of 4 warps, group g drains accumulator stage g.
    // (Letting all 8 warps share every tile -- half the channels each, to halve the time an
    // accumulator stage is held -- was measured SLOWER, 0.59 -> 0.73 ms on dec8, both with a joint
    // 256-thread barrier per tile and with two fully independent half-channel groups.)
    const int grp = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;   // TMEM lane = slab pixel (jy, jx)
    const int jy = row / kTileW, jx = row % kTileW;
    const int ox = jx - 1;              // output column inside the tile
    const bool col_ok = jx >= 1 && jx <= kSmOutW;
    const bool issuer_warp = (quad == 0);
    const uint32_t sbuf = store_base + grp * kSmStoreBytes;
    SatTracker<T16> sat;
    for (int it = grp;; it += 2) {
      const long long unit_ll = (long long)unit0 + (long long)it * unit_step;
      if (unit_ll >= p.total_tiles) break;
      const TileCoord t = decode_tile_sm<CG>(p, (int)unit_ll, (int)cta_rank);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int y = t.y0 + jy, x = t.x0 + ox;
      const bool valid = col_ok && (y < p.H) && (x < p.W) && (CG == 1 || t.n < p.N);
      MBAR_WAIT_RELAXED(tmem_full_bar(it % Cfg::kFullBars), (uint32_t)(it / Cfg::kFullBars) & 1u, 600 + as);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * kSmN);
      if (CCST_ABLATE_BITS(p) & 1) {  // measurement only: hand the accumulator back untouched
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(lead(tmem_empty_bar(as)));
          else mbar_arrive(tmem_empty_bar(as));
        }
        continue;
      }
      uint32_t pk[32];
#pragma unroll
      for (int cq = 0; cq < 4; ++cq) {
        uint32_t a[16], b[16], c[16];
        tmem_ld16(taddr + 0 * 64 + cq * 16, a);
        tmem_ld16(taddr + 1 * 64 + cq * 16, b);
        tmem_ld16(taddr + 2 * 64 + cq * 16, c);
        tmem_ld_wait();
        if (cq == 3) {
          // all TMEM reads of this accumulator stage are complete: hand it back before the math
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(lead(tmem_empty_bar(as)));
            else mbar_arrive(tmem_empty_bar(as));
          }
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float lft = __shfl_up_sync(0xffffffffu, __uint_as_float(a[j]), 1);    // P[jx-1][s=0]
          const float rgt = __shfl_down_sync(0xffffffffu, __uint_as_float(c[j]), 1);  // P[jx+1][s=2]
          v[j] = (lft + (__uint_as_float(b[j]) + s_bias[cq * 16 + j])) + rgt;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t w = p.relu ? pack16x2_relu<T16>(v[2 * j], v[2 * j + 1])
                              : pack16x2<T16>(v[2 * j], v[2 * j + 1]);
          if (EPI == EPI_ACT_POOL) {
            // 2x2 window: columns (jx odd, jx + 1), rows (jy even, jy + 1) = lanes l, l+1, l^16, ...,
            // pooled on the packed pairs (exact, see max16x2); invalid pixels contribute 0
            w = valid ? w : 0u;
            w = max16x2<T16>(w, __shfl_down_sync(0xffffffffu, w, 1));
            w = max16x2<T16>(w, __shfl_xor_sync(0xffffffffu, w, 16));
          }
          pk[cq * 8 + j] = w;
          sat.track(w);
        }
      }
      // the staging buffer about to be rewritten must have been read out by its TMA store
      if (issuer_warp) bulk_wait_read<0>();
      epi_barrier(grp);
      int srow = jy * kSmOutW + ox;
      bool writer = col_ok;
      if (EPI == EPI_ACT_POOL) {
        writer = col_ok && (jx & 1) && lane < 16;  // anchor of a 2x2 window
        srow = (jy >> 1) * (kSmOutW / 2) + (ox >> 1);
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
        if (EPI == EPI_ACT) {
          store_aliases(p.out, t.n, y, x, 0, pk, p.halo_edge);
        } else if (EPI == EPI_ACT_UP2) {
#pragma unroll
          for (int aa = 0; aa < 2; ++aa)
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) store_aliases(p.out, t.n, 2 * y + aa, 2 * x + bb, 0, pk);
        } else if (EPI == EPI_ACT_POOL) {
          if (writer) store_aliases(p.out, t.n, y >> 1, x >> 1, 0, pk);
        }
      }
      fence_async_smem();
      epi_barrier(grp);
      if (issuer_warp && elect_one()) {
        if (EPI == EPI_ACT_POOL) {
          tma_store_4d(&tmap_out.m[0], sbuf, 0, t.x0 >> 1, t.y0 >> 1, t.n);
        } else {
          tma_store_4d(&tmap_out.m[0], sbuf, 0, t.x0, t.y0, t.n);
          if (EPI == EPI_ACT_UP2) {
            tma_store_4d(&tmap_out.m[1], sbuf, 0, t.x0, t.y0, t.n);
            tma_store_4d(&tmap_out.m[2], sbuf, 0, t.x0, t.y0, t.n);
            tma_store_4d(&tmap_out.m[3], sbuf, 0, t.x0, t.y0, t.n);
          }
        }
        bulk_commit();
      }
    }
    if (issuer_warp) bulk_wait_all();
    sat.flush(p.sat_count);
  }

  __syncwarp();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc_cg<CG, Cfg::kTmemCols>(tmem_base);
}

template <typename T16, int EPI, bool BRES, int CG>
int launch_smerge_cfg(const CUtensorMap& ma, const T16* wk_sm, ConvParams<T16> p, cudaStream_t st) {
  using Cfg = SmergeCfg<BRES, CG>;
  CUtensorMap mb;
  if (int e = make_weight_map(&mb, wk_sm, 3 * p.Cin, kSmN, Cfg::kBRows)) return e;
  OutMaps mo;
  memset(&mo, 0, sizeof(mo));
  if (EPI == EPI_ACT) {
    if (int e = make_out_map(&mo.m[0], p.out, 0, 0, 1, 1, kSmOutW, kTileH)) return e;
  } else if (EPI == EPI_ACT_POOL) {
    if (int e = make_out_map(&mo.m[0], p.out, 0, 0, 1, 1, kSmOutW / 2, kTileH / 2)) return e;
  } else {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        if (int e = make_out_map(&mo.m[a * 2 + b], p.out, a, b, 2, 2, kSmOutW, kTileH)) return e;
  }
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_smerge_kernel<T16, EPI, BRES, CG>), Cfg::kSmemBytes));
  p.tiles_x = (p.W + kSmOutW - 1) / kSmOutW;
  p.n_tiles = 1;
  const int64_t m_tiles = (int64_t)p.N * p.tiles_x * p.tiles_y;
  CCST_CHECK_ARG(m_tiles < (1ll << 31), "conv_smerge: too many tiles");
  p.m_tiles = (int)m_tiles;
  const int64_t units = (m_tiles + CG - 1) / CG;
  p.total_tiles = (int)units;
  const int slots = sm_count() / CG;
  const int grid = (int)(units < slots ? units : slots) * CG;
  CCST_CUDA(launch_conv(conv_smerge_kernel<T16, EPI, BRES, CG>, grid, Cfg::kThreads, Cfg::kSmemBytes, st, CG, ma,
                        mb, mo, p));
  CCST_LAUNCHED();
  return CCST_OK;
}

// CTA pairs everywhere (dec7: 0.30 -> 0.26 ms, conv1_2 0.635 -> 0.58: each CTA stages half of every
// weight tile); Cin == 64 keeps its three weight tiles resident.
template <typename T16>
int launch_smerge(const CUtensorMap& ma, const T16* wk_sm, const ConvParams<T16>& p, int epi,
                  cudaStream_t st) {
  const bool res = p.Cin == kBlockK;
  switch (epi) {
    case EPI_ACT:
      return res ? launch_smerge_cfg<T16, EPI_ACT, true, 2>(ma, wk_sm, p, st)
                 : launch_smerge_cfg<T16, EPI_ACT, false, 2>(ma, wk_sm, p, st);
    case EPI_ACT_UP2:
      return res ? launch_smerge_cfg<T16, EPI_ACT_UP2, true, 2>(ma, wk_sm, p, st)
                 : launch_smerge_cfg<T16, EPI_ACT_UP2, false, 2>(ma, wk_sm, p, st);
    case EPI_ACT_POOL:
      return res ? launch_smerge_cfg<T16, EPI_ACT_POOL, true, 2>(ma, wk_sm, p, st)
                 : launch_smerge_cfg<T16, EPI_ACT_POOL, false, 2>(ma, wk_sm, p, st);
    default:
      set_error("conv_smerge: epilogue %d not available", epi);
      return CCST_EINVAL;
  }
}

}  // namespace
}  // namespace ccst
