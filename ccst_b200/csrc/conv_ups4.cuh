// dec8 (upsample-fused 64 -> 64 conv): all four output phases per tile over one linear slab.
#pragma once
#include "umma_common.cuh"

namespace ccst {
namespace {

// =====================================================================================
// Upsample-fused 64 -> 64 conv (dec8, net.py:30-32), all four output phases per tile.
// At N = 64 a tcgen05.mma is bound by the shared-memory fetch of its A operand (128 x 32 B at
// 64 B/clk = 64 cycles for 32 cycles of math), and the per-phase kernel issues 16 such MMAs per
// K step and source tile.  Here one linear slab {64 ch, 32 px, 6 rows} of the low-resolution map
// feeds all four phases: the operand view (R, S) = slab shifted by R rows and S pixels is shared
// by every (phase, tap) with a + dy = R, b + dx = S, so their weight tiles are stacked in N:
//   view (1,1): 4 phases, N = 256;  (0,1) (1,0) (1,2): 2 phases, N = 128;  (2,1): 2 x N = 64 (its two
//   phases are not adjacent in TMEM);  corners: N = 64       -> 10 MMAs per K step instead of 16,
// with accumulator columns ordered [phase 10 | 00 | 01 | 11].  The N = 256 view is issued first and
// initialises all four accumulators.  All 16 weight tiles (128 KiB) stay resident in shared memory.
// =====================================================================================
constexpr int kU4BoxW = 32, kU4Rows = 4, kU4OutW = kU4BoxW - 2;
constexpr int kU4SlabBytes = (kU4Rows + 2) * kU4BoxW * 128;  // 24576
// The kernel runs as CTA PAIRS (cta_group::2, two consecutive tiles per M = 256 MMA): every SM fetches only its
// own tile's A operand, each CTA keeps HALF of every stacked weight block (64 KiB instead of 128), and the freed
// shared memory holds a third slab stage and one staging tile per epilogue group.  Measured at batch 32 @512^2
// (same box, interleaved, profiles/r02z_ab_ups4_2_vs_4_groups.txt): pairs alone changed nothing (0.302 ms against
// 0.298-0.308 for the single-CTA form); with FOUR epilogue groups (below) 0.307 -> 0.278 ms.  The following
// last conv then takes 0.013 ms longer (0.222 -> 0.236): it starts while ~100 MB of this kernel's output are
// still being written back from L2 -- together the two kernels move 2.5 GB in 0.51 ms = 4.9 TB/s, and that HBM
// round trip of the 64-channel 512^2 map, not either kernel, is what is left.
constexpr int kU4CG = 2;
constexpr int kU4EpiGroups = 4;                              // one 4-warp epilogue group per output phase
constexpr int kU4Threads = 32 * (kEpiWarp0 + 4 * kU4EpiGroups);  // 640
constexpr int kU4AStages = 3;
constexpr int kU4OffB = kU4AStages * kU4SlabBytes;           // 16 half-tiles of 32 rows x 128 B per CTA
constexpr int kU4BBytes = 16 * 4096;
constexpr int kU4OffStore = kU4OffB + kU4BBytes;
constexpr int kU4StoreBytes = 16384;                          // 120 rows x 128 B, rounded
constexpr int kU4OffBias = kU4OffStore + kU4EpiGroups * kU4StoreBytes;   // one staging tile per epilogue group
constexpr int kU4OffBar = kU4OffBias + 256;
constexpr int kU4NumBars = 2 * kU4AStages + 4 + 1;
constexpr int kU4Smem = 1024 + kU4OffBar + 8 * kU4NumBars + 16;
static_assert(kU4Smem <= 232448, "ups4 shared memory plan exceeds 227 KiB");

struct U4Op {
  int R, S, first, ntiles, slot;
};
// issue order: the 4-phase view first (accumulate = 0), then the rest
__device__ constexpr U4Op kU4Ops[10] = {{1, 1, 6, 4, 0}, {0, 0, 0, 1, 1}, {0, 1, 1, 2, 1}, {0, 2, 3, 1, 2},
                                        {1, 0, 4, 2, 0}, {1, 2, 10, 2, 2}, {2, 0, 12, 1, 0}, {2, 1, 13, 1, 0},
                                        {2, 1, 14, 1, 3}, {2, 2, 15, 1, 3}};
// resident slot -> (phase = a*2+b, tap = dy*2+dx) of the packed phase weights
__device__ constexpr int kU4TilePh[16] = {0, 0, 1, 1, 2, 0, 2, 0, 1, 3, 1, 3, 2, 2, 3, 3};
__device__ constexpr int kU4TileTap[16] = {0, 1, 0, 1, 0, 2, 1, 3, 2, 0, 3, 1, 2, 3, 2, 3};
__device__ constexpr int kU4SlotPh[4] = {2, 0, 1, 3};  // accumulator column slot -> phase

template <typename T16>
__global__ void __launch_bounds__(kU4Threads, 1)
    conv_ups4_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ OutMaps tmap_out, ConvParams<T16> p) {
  constexpr int CG = kU4CG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kU4OffBar;
  float* s_bias = reinterpret_cast<float*>(gen + kU4OffBias);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (kU4AStages + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * kU4AStages + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * kU4AStages + 2 + s); };
  const uint32_t bres_bar = bar0 + 8u * (2 * kU4AStages + 4);
  const uint32_t tmem_slot = bar0 + 8u * kU4NumBars;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  auto lead = [&](uint32_t bar) { return mapa_rank(bar, 0); };
  const int unit_id = (int)cluster_id_x(), unit_cnt = (int)ncluster_x();
  const int units = (p.total_tiles + CG - 1) / CG;
  // tile of this CTA in work unit `unit`; the second tile of an odd last pair is a dummy at image n = N (TMA
  // loads out of bounds = zero fill, TMA stores clipped away, alias stores masked)
  auto tile_of = [&](int unit, int& n, int& y0, int& x0) {
    int tile = unit * CG + (int)cta_rank;
    if (tile >= p.total_tiles) {
      n = p.N, y0 = 0, x0 = 0;
      return;
    }
    x0 = (tile % p.tiles_x) * kU4OutW;
    tile /= p.tiles_x;
    y0 = (tile % p.tiles_y) * kU4Rows;
    n = tile / p.tiles_y;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_out.m[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kU4AStages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), 4 * kU4EpiGroups * CG);  // one arrive per epilogue warp of both CTAs
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg<CG, 512>(tmem_slot);
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + kU4OffBar + 8 * kU4NumBars);
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer: this CTA's half of every stacked weight block once (the operand view of
    // kU4Ops[o] multiplies ntiles x 64 stacked rows; rank r keeps rows r N/2 .. (r+1) N/2 - 1 of them, as 32-row
    // boxes in issue order), then one slab per tile
    if (elect_one()) {
      if (leader) mbar_expect_tx(bres_bar, CG * kU4BBytes);
      const uint32_t bar = lead(bres_bar);
      int blk = 0;
#pragma unroll
      for (int o = 0; o < 10; ++o) {
#pragma unroll
        for (int j = 0; j < kU4Ops[o].ntiles; ++j, ++blk) {
          const int row = (int)cta_rank * 32 * kU4Ops[o].ntiles + 32 * j;  // stacked row of the view's weight block
          const int t = kU4Ops[o].first + (row >> 6);
          tma_load_2d_cg<CG>(base + kU4OffB + blk * 4096, &tmap_b, bar, kU4TileTap[t] * kBlockK,
                             kU4TilePh[t] * 64 + (row & 32));
        }
      }
    }
    __syncwarp();
    pdl_wait();
    int s = 0;
    uint32_t ph = 0;
    for (int unit = unit_id; unit < units; unit += unit_cnt) {
      int n, y0, x0;
      tile_of(unit, n, y0, x0);
      MBAR_WAIT_RELAXED(a_empty(s), ph ^ 1, 900 + s);
      if (elect_one()) {
        if (leader) mbar_expect_tx(a_full(s), CG * kU4SlabBytes);
        // slab position (jy, jx) = padded pixel (y0 + jy, x0 + jx) = source (y0 - 1 + jy, x0 - 1 + jx)
        tma_load_4d_cg<CG>(base + s * kU4SlabBytes, &tmap_a, lead(a_full(s)), 0, x0, y0, n);
      }
      __syncwarp();
      if (++s == kU4AStages) s = 0, ph ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader): 10 operand views x 4 K steps per PAIR of tiles
    if (leader) {
      mbar_wait(bres_bar, 0, 905);
      tc_fence_after();
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int unit = unit_id; unit < units; unit += unit_cnt, ++it) {
        const int acs = it & 1;
        mbar_wait_cluster(t_empty(acs), ((it >> 1) & 1) ^ 1, 910 + acs);
        mbar_wait(a_full(s), ph, 920 + s);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc0 = make_kmajor_sw128_desc(base + s * kU4SlabBytes);
          const uint64_t bdesc0 = make_kmajor_sw128_desc(base + kU4OffB);
          const uint32_t d0 = tmem_base + (uint32_t)(acs * 256);
          int blk = 0;
#pragma unroll
          for (int o = 0; o < 10; ++o) {
            const uint64_t adesc = adesc0 + (uint64_t)((kU4Ops[o].R * kU4BoxW + kU4Ops[o].S) * 128 >> 4);
            const uint64_t bdesc = bdesc0 + (uint64_t)(blk * (4096 >> 4));
            const uint32_t d = d0 + (uint32_t)(kU4Ops[o].slot * 64);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint32_t acc = (o | k) ? 1u : 0u;
              if (CCST_ABLATE_BITS(p) & 2) continue;  // (CCST_DEV builds: measurement only)
              if (kU4Ops[o].ntiles == 4) umma_f16_cg<CG>(d, adesc + 2 * k, bdesc + 2 * k, make_idesc<T16, 256, CG>(), acc);
              else if (kU4Ops[o].ntiles == 2) umma_f16_cg<CG>(d, adesc + 2 * k, bdesc + 2 * k, make_idesc<T16, 128, CG>(), acc);
              else umma_f16_cg<CG>(d, adesc + 2 * k, bdesc + 2 * k, make_idesc<T16, 64, CG>(), acc);
            }
            blk += kU4Ops[o].ntiles;
          }
          umma_commit_cg<CG>(a_empty(s));
          umma_commit_cg<CG>(t_full(acs));
        }
        __syncwarp();
        if (++s == kU4AStages) s = 0, ph ^= 1;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: FOUR groups of 4 warps, group g drains output phase slot g (64 accumulator
    // columns) of EVERY tile; warp <-> tile row, lane <-> column.  Stage ablation (CCST_DEV build,
    // tools/ablate_ups4.sh, batch 32 @512^2): MMAs + loads alone 0.18 ms, the two-group epilogue alone 0.26 ms
    // (tcgen05.ld + math 0.13, staging 0.06, TMA store 0.07 -- one serial chain per chunk, four chunks per tile,
    // two warps per scheduler at ~0.2 instructions per clock each): the epilogue was the bound and it was a
    // LATENCY bound, so it gets twice the warps and a quarter of the chain per tile.  (Prefetching the next
    // 32-column half into a second register buffer with two groups was measured first: slower, 0.32 ms.)
    const int grp = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;
    const bool issuer_warp = (quad == 0);
    const uint32_t sbuf = base + kU4OffStore + grp * kU4StoreBytes;
    const int srow = quad * kU4OutW + lane;
    const int phs = kU4SlotPh[grp], a = phs >> 1, b = phs & 1;
    SatTracker<T16> sat;
    int it = 0;
    for (int unit = unit_id; unit < units; unit += unit_cnt, ++it) {
      int n, y0, x0;
      tile_of(unit, n, y0, x0);
      const int acs = it & 1;
      const int y = y0 + quad, x = x0 + lane;
      const bool col_ok = lane < kU4OutW;
      const bool valid = col_ok && y < p.H && x < p.W && n < p.N;
      MBAR_WAIT_RELAXED(t_full(acs), (it >> 1) & 1, 930 + acs);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acs * 256 + grp * 64);
      if (CCST_ABLATE_BITS(p) & 1) {  // (CCST_DEV builds: measurement only) hand the accumulator back untouched
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead(t_empty(acs)));
        continue;
      }
      uint32_t pk[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld32(taddr + half * 32, r);
        tmem_ld_wait();
        if (half == 1) {
          // this group's columns are in registers: hand them back (the stage is free once all 16 warps have)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(lead(t_empty(acs)));
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 v = add2_f32(make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])),
                                    *reinterpret_cast<const float2*>(&s_bias[half * 32 + 2 * j]));
          pk[half * 16 + j] = p.relu ? pack16x2_relu<T16>(v.x, v.y) : pack16x2<T16>(v.x, v.y);
        }
      }
      // (The per-word `p.relu ? :` select compiles to BOTH conversions under complementary predicates; making ReLU a
      // compile-time property of this loop -- two instances, one uniform branch -- was measured: 0.247 -> 0.293 ms,
      // profiles/r03s_ab_relu_compile_time_dispatch.txt.  Kept as it is.)
      sat.track_block(pk, p.relu != 0);
      // the group's staging tile has been read out by the TMA store of its previous tile
      if (issuer_warp) bulk_wait_read<0>();
      epi_barrier(grp);
      if (col_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t dst = sbuf + srow * 128 + ((j ^ (srow & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]),
                       "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                       : "memory");
        }
      }
      if (valid) store_aliases(p.out, n, 2 * y + a, 2 * x + b, 0, pk);
      fence_async_smem();
      epi_barrier(grp);
      if (issuer_warp && elect_one() && !(CCST_ABLATE_BITS(p) & 16)) {
        tma_store_4d(&tmap_out.m[phs], sbuf, 0, x0, y0, n);
        bulk_commit();
      }
    }
    if (issuer_warp) bulk_wait_all();
    sat.flush(p.sat_count);
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_cg<CG, 512>(tmem_base);
}

template <typename T16>
int launch_ups4(ActView<T16> in, const T16* wk_up, ConvParams<T16> p, cudaStream_t st) {
  constexpr int CG = kU4CG;
  CUtensorMap m4, mb;
  if (int e = make_act_map(&m4, in, kU4BoxW, kU4Rows + 2)) return e;
  if (int e = make_weight_map(&mb, wk_up, 4 * in.C, 4 * p.Cout, 32)) return e;  // half-tile boxes of 32 rows
  OutMaps mo;
  memset(&mo, 0, sizeof(mo));
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      if (int e = make_out_map(&mo.m[a * 2 + b], p.out, a, b, 2, 2, kU4OutW, kU4Rows)) return e;
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_ups4_kernel<T16>), kU4Smem));
  p.tiles_x = (in.W + kU4OutW - 1) / kU4OutW;
  p.tiles_y = (in.H + kU4Rows - 1) / kU4Rows;
  const int64_t tiles = (int64_t)in.N * p.tiles_x * p.tiles_y;
  CCST_CHECK_ARG(tiles < (1ll << 31) - 2, "conv_ups4: too many tiles");
  p.m_tiles = p.total_tiles = (int)tiles;
  const int64_t units = (tiles + CG - 1) / CG;
  const int slots = sm_count() / CG;
  const int grid = (int)(units < slots ? units : slots) * CG;
  CCST_CUDA(launch_conv(conv_ups4_kernel<T16>, grid, kU4Threads, kU4Smem, st, CG, m4, mb, mo, p));
  CCST_LAUNCHED();
  return CCST_OK;
}

}  // namespace
}  // namespace ccst
