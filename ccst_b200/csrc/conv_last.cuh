// Last decoder conv (64 -> 3 channels, HBM-bound) on tcgen05.
#pragma once
#include "umma_common.cuh"

namespace ccst {
namespace {

// =====================================================================================
// Last decoder conv (net.py:34-35, 64 -> 3 channels, no ReLU, fp32 NCHW or uint8 NHWC result):
// filter ROWS by operand shifts, filter COLUMNS in N.  HBM-bound (reads 128 B, writes 12 B or 3 B per
// pixel).  The tile is 4 rows x 32 columns of a linear slab {64 ch, 32 px, 6 rows} (one TMA load):
//   P[(jy, jx), (s, co)] = sum_{r, c} X[(jy + r, jx), c] * W[co][c][r][s]     3 MMAs chains (r), N = 16
//   out[(y, x), co]      = bias[co] + P[(y, x), (0, co)] + P[(y, x+1), (1, co)] + P[(y, x+2), (2, co)]
// The row shift r is a start-address offset of r * 32 * 128 B into the slab; the column shift s
// is two warp shuffles in the epilogue (one tile row = one warp = one TMEM lane quadrant), so a
// thread reads 16 accumulator columns and writes its pixel: no shared-memory staging at all.
// 30 of the 32 columns are outputs (the last two would need the next tile's pixels).
// =====================================================================================
constexpr int kRBoxW = 32, kRRows = 4, kROutW = kRBoxW - 2;
constexpr int kRSlabBytes = (kRRows + 2) * kRBoxW * 128;  // 24576
constexpr int kRStages = 6;
constexpr int kROffB = kRStages * kRSlabBytes;            // 3 weight tiles of 16 rows x 128 B
constexpr int kROffBar = kROffB + 3 * 2048 + 1024;        // (+ slack: the last MMA rows read 256 B past a slab)
constexpr int kRNumBars = 2 * kRStages + 4;
constexpr int kRSmem = 1024 + kROffBar + 8 * kRNumBars + 16;

// CG = 2: a CTA pair runs two consecutive tiles as ONE M = 256 MMA (cta_group::2).  At N = 16 a single-CTA
// tcgen05.mma is bound by the fetch of its A operand (64 cycles per MMA for 8 cycles of math), and the 12 MMAs of
// a tile took 0.20 ms of the kernel's 0.26 at batch 32 @512^2 -- above its HBM time; in the pair each SM fetches
// only its own tile's slab, so the MMA time per tile halves.  B: each CTA holds 8 of the 16 (s, co) rows.
template <typename T16, int CG>
__global__ void __launch_bounds__(kThreadsUmma, 1)
    conv_last_rows_kernel(const __grid_constant__ CUtensorMap tmap_a, const T16* __restrict__ wk,
                          ConvParams<T16> p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kROffBar;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (kRStages + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * kRStages + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * kRStages + 2 + s); };
  const uint32_t tmem_slot = bar0 + 8u * kRNumBars;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  auto lead = [&](uint32_t bar) { return CG == 2 ? mapa_rank(bar, 0) : bar; };
  const int unit_id = CG == 2 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int unit_cnt = CG == 2 ? (int)ncluster_x() : (int)gridDim.x;
  const int units = (p.total_tiles + CG - 1) / CG;
  // tile of this CTA in work unit `unit`; the second tile of an odd last pair is a dummy at image n = N (its TMA
  // box is out of bounds = zero fill, its stores are masked)
  auto tile_of = [&](int unit, int& n, int& y0, int& x0) {
    int tile = unit * CG + (int)cta_rank;
    if (tile >= p.total_tiles) {
      n = p.N, y0 = 0, x0 = 0;
      return;
    }
    // (consuming the tiles in reverse order, so that the producer's last ~100 MB are found in L2, was measured:
    // no gain)
    x0 = (tile % p.tiles_x) * kROutW;
    tile /= p.tiles_x;
    y0 = (tile % p.tiles_y) * kRRows;
    n = tile / p.tiles_y;
  };

  // B_r[n = s*4 + co][k = c] = W[co][c][r][s] from the packed weights wk[co][(r*3+s)*64 + c]; K-major
  // rows of 128 B with the 128-byte swizzle; unused rows are zero.  A pair: this CTA keeps rows 8 rank .. 8 rank + 7.
  constexpr int kRowsB = 16 / CG;
  for (int i = threadIdx.x; i < 3 * kRowsB * 8; i += kThreadsUmma) {
    const int r = i / (kRowsB * 8), nl = (i >> 3) % kRowsB, j = i & 7;
    const int n = nl + (int)cta_rank * kRowsB;
    const int sc = n >> 2, co = n & 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (sc < 3 && co < p.Cout && co < 3)
      v = *reinterpret_cast<const uint4*>(wk + (size_t)co * (9 * kBlockK) + (r * 3 + sc) * kBlockK + j * 8);
    *reinterpret_cast<uint4*>(gen + kROffB + r * 2048 + nl * 128 + ((j ^ (nl & 7)) << 4)) = v;
  }
  if (warp == 0 && lane == 0) prefetch_tmap(&tmap_a);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRStages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), 4 * CG);  // one arrive per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg<CG, 32>(tmem_slot);
  fence_async_smem();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + kROffBar + 8 * kRNumBars);
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer: one slab per tile; the leader's "full" barrier counts both CTAs' bytes
    int s = 0;
    uint32_t ph = 0;
    for (int unit = unit_id; unit < units; unit += unit_cnt) {
      int n, y0, x0;
      tile_of(unit, n, y0, x0);
      MBAR_WAIT_RELAXED(a_empty(s), ph ^ 1, 800 + s);
      if (elect_one()) {
        if (leader) mbar_expect_tx(a_full(s), CG * kRSlabBytes);
        tma_load_4d_cg<CG>(base + s * kRSlabBytes, &tmap_a, lead(a_full(s)), 0, x0, y0, n);
      }
      __syncwarp();
      if (++s == kRStages) s = 0, ph ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader): 3 filter rows x 4 K steps, M = 128 CG, N = 16
    if (leader) {
      constexpr uint32_t idesc = make_idesc<T16, 16, CG>();
      const uint64_t bdesc0 = make_kmajor_sw128_desc(base + kROffB);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int unit = unit_id; unit < units; unit += unit_cnt, ++it) {
        const int acs = it & 1;
        if (CG == 2) mbar_wait_cluster(t_empty(acs), ((it >> 1) & 1) ^ 1, 810 + acs);
        else mbar_wait(t_empty(acs), ((it >> 1) & 1) ^ 1, 810 + acs);
        mbar_wait(a_full(s), ph, 820 + s);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc0 = make_kmajor_sw128_desc(base + s * kRSlabBytes);
          const uint32_t d = tmem_base + (uint32_t)(acs * 16);
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16_cg<CG>(d, adesc0 + (uint64_t)(r * (kRBoxW * 128 >> 4) + 2 * k),
                              bdesc0 + (uint64_t)(r * (2048 >> 4) + 2 * k), idesc, (r | k) ? 1u : 0u);
          umma_commit_cg<CG>(a_empty(s));
          umma_commit_cg<CG>(t_full(acs));
        }
        __syncwarp();
        if (++s == kRStages) s = 0, ph ^= 1;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: group g takes tiles g, g+2, ...; warp <-> tile row, lane <-> column
    const int grp = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;
    float bias[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bias[c] = c < p.Cout ? p.bias[c] : 0.f;
    for (int it = grp;; it += 2) {
      const long long unit_ll = (long long)unit_id + (long long)it * unit_cnt;
      if (unit_ll >= units) break;
      int n, y0, x0;
      tile_of((int)unit_ll, n, y0, x0);
      const int acs = it & 1;
      MBAR_WAIT_RELAXED(t_full(acs), (it >> 1) & 1, 830 + acs);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acs * 16), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(lead(t_empty(acs)));
        else mbar_arrive(t_empty(acs));
      }
      const int y = y0 + quad, x = x0 + lane;
      const bool ok = lane < kROutW && y < p.H && x < p.W && n < p.N;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float p1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[4 + c]), 1);
        const float p2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[8 + c]), 2);
        float o = ((bias[c] + __uint_as_float(v[c])) + p1) + p2;
        if (p.relu) o = fmaxf(o, 0.f);
        if (ok && c < p.Cout) {
          if (p.out_u8) p.out_u8[(((size_t)n * p.H + y) * p.W + x) * p.Cout + c] = quantize_u8(o);
          else p.out_nchw[(((size_t)n * p.Cout + c) * p.H + y) * p.W + x] = o;
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc_cg<CG, 32>(tmem_base);
}

template <typename T16>
int launch_last_rows(ActView<T16> in, const T16* wk, ConvParams<T16> p, cudaStream_t st) {
  constexpr int CG = 2;
  CUtensorMap mr;
  if (int e = make_act_map(&mr, in, kRBoxW, kRRows + 2)) return e;
  auto kernel = conv_last_rows_kernel<T16, CG>;
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(kernel), kRSmem));
  p.tiles_x = (in.W + kROutW - 1) / kROutW;
  p.tiles_y = (in.H + kRRows - 1) / kRRows;
  const int64_t tiles = (int64_t)in.N * p.tiles_x * p.tiles_y;
  CCST_CHECK_ARG(tiles < (1ll << 31) - 2, "conv_last_rows: too many tiles");
  p.m_tiles = p.total_tiles = (int)tiles;
  const int64_t units = (tiles + CG - 1) / CG;
  const int slots = sm_count() / CG;
  const int grid = (int)(units < slots ? units : slots) * CG;
  CCST_CUDA(launch_conv(kernel, grid, kThreadsUmma, kRSmem, st, CG, mr, wk, p));
  CCST_LAUNCHED();
  return CCST_OK;
}

// =====================================================================================
// The same last conv for the x3 engines (split operands, conv_x3.cuh): the input is a [hi | lo] map of 64
// logical channels, so a tile takes TWO slabs (channels 0..63 = a_hi, 64..127 = a_lo), and the weight tile of
// filter row r has N = 32 rows [w_hi (s, co) | w_lo (s, co)] that both slabs meet:
//   P_h[(jy, jx), (half, s, co)] = sum_{r, c} X_h[(jy + r, jx), c] * W_half[co][c][r][s]       h = hi, lo slab
// Each slab's 12 MMAs accumulate into their own 32 TMEM columns (the promoted-partial granularity of
// conv_x3_kernel: tcgen05 accumulates with truncation), the epilogue adds the four terms in fp32, applies
// 2^-e and the bias, and shifts / sums the filter columns by two shuffles.  HBM-bound (reads 256 B per pixel).
// wk_x3: [hi | lo][64 co rows, <= 3 used][9 * 64] as packed by pack_layer.
// =====================================================================================
constexpr int kR3Stages = 3;                                // pairs of slabs
constexpr int kR3OffB = kR3Stages * 2 * kRSlabBytes;        // 3 weight tiles of 32 rows x 128 B
constexpr int kR3OffBar = kR3OffB + 3 * 4096 + 1024;
constexpr int kR3NumBars = 2 * kR3Stages + 4;
constexpr int kR3Smem = 1024 + kR3OffBar + 8 * kR3NumBars + 16;
static_assert(kR3Smem <= 232448, "conv_last_rows_x3: shared memory plan exceeds 227 KiB");

template <typename T16, int CG>
__global__ void __launch_bounds__(kThreadsUmma, 1)
    conv_last_rows_x3_kernel(const __grid_constant__ CUtensorMap tmap_a, const T16* __restrict__ wk_x3,
                             ConvParams<T16> p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kR3OffBar;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (kR3Stages + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * kR3Stages + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * kR3Stages + 2 + s); };
  const uint32_t tmem_slot = bar0 + 8u * kR3NumBars;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA pairs as in conv_last_rows_kernel (24 MMAs of N = 32 per tile are A-fetch-bound at twice its cost)
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  auto lead = [&](uint32_t bar) { return CG == 2 ? mapa_rank(bar, 0) : bar; };
  const int unit_id = CG == 2 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int unit_cnt = CG == 2 ? (int)ncluster_x() : (int)gridDim.x;
  const int units = (p.total_tiles + CG - 1) / CG;
  auto tile_of = [&](int unit, int& n, int& y0, int& x0) {
    int tile = unit * CG + (int)cta_rank;
    if (tile >= p.total_tiles) {  // dummy second tile of an odd last pair
      n = p.N, y0 = 0, x0 = 0;
      return;
    }
    x0 = (tile % p.tiles_x) * kROutW;
    tile /= p.tiles_x;
    y0 = (tile % p.tiles_y) * kRRows;
    n = tile / p.tiles_y;
  };

  // B_r[n = half*16 + s*4 + co][k = c] = W_half[co][c][r][s]; K-major rows of 128 B, 128-byte swizzle, unused rows
  // zero.  A pair: rank 0 keeps the w_hi rows (n < 16), rank 1 the w_lo rows.
  constexpr int kRowsB = 32 / CG;
  for (int i = threadIdx.x; i < 3 * kRowsB * 8; i += kThreadsUmma) {
    const int r = i / (kRowsB * 8), nl = (i >> 3) % kRowsB, j = i & 7;
    const int n = nl + (int)cta_rank * kRowsB;
    const int half = n >> 4, sc = (n >> 2) & 3, co = n & 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (sc < 3 && co < p.Cout && co < 3)
      v = *reinterpret_cast<const uint4*>(wk_x3 + (size_t)(half * 64 + co) * (9 * kBlockK) + (r * 3 + sc) * kBlockK + j * 8);
    *reinterpret_cast<uint4*>(gen + kR3OffB + r * 4096 + nl * 128 + ((j ^ (nl & 7)) << 4)) = v;
  }
  if (warp == 0 && lane == 0) prefetch_tmap(&tmap_a);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kR3Stages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), 4 * CG);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg<CG, 128>(tmem_slot);
  fence_async_smem();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + kR3OffBar + 8 * kR3NumBars);

  if (warp == 0) {
    // ===================== TMA producer: the hi and the lo slab of a tile under one barrier
    int s = 0;
    uint32_t ph = 0;
    for (int unit = unit_id; unit < units; unit += unit_cnt) {
      int n, y0, x0;
      tile_of(unit, n, y0, x0);
      MBAR_WAIT_RELAXED(a_empty(s), ph ^ 1, 840 + s);
      if (elect_one()) {
        if (leader) mbar_expect_tx(a_full(s), CG * 2 * kRSlabBytes);
        tma_load_4d_cg<CG>(base + (2 * s) * kRSlabBytes, &tmap_a, lead(a_full(s)), 0, x0, y0, n);
        tma_load_4d_cg<CG>(base + (2 * s + 1) * kRSlabBytes, &tmap_a, lead(a_full(s)), kBlockK, x0, y0, n);
      }
      __syncwarp();
      if (++s == kR3Stages) s = 0, ph ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader): per slab 3 filter rows x 4 K steps, M = 128 CG, N = 32
    if (leader) {
      constexpr uint32_t idesc = make_idesc<T16, 32, CG>();
      const uint64_t bdesc0 = make_kmajor_sw128_desc(base + kR3OffB);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int unit = unit_id; unit < units; unit += unit_cnt, ++it) {
        const int acs = it & 1;
        if (CG == 2) mbar_wait_cluster(t_empty(acs), ((it >> 1) & 1) ^ 1, 850 + acs);
        else mbar_wait(t_empty(acs), ((it >> 1) & 1) ^ 1, 850 + acs);
        mbar_wait(a_full(s), ph, 860 + s);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const uint64_t adesc0 = make_kmajor_sw128_desc(base + (2 * s + h2) * kRSlabBytes);
            const uint32_t d = tmem_base + (uint32_t)(acs * 64 + h2 * 32);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_f16_cg<CG>(d, adesc0 + (uint64_t)(r * (kRBoxW * 128 >> 4) + 2 * k),
                                bdesc0 + (uint64_t)(r * (4096 >> 4) + 2 * k), idesc, (r | k) ? 1u : 0u);
          }
          umma_commit_cg<CG>(a_empty(s));
          umma_commit_cg<CG>(t_full(acs));
        }
        __syncwarp();
        if (++s == kR3Stages) s = 0, ph ^= 1;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: group g takes tiles g, g+2, ...; warp <-> tile row, lane <-> column
    const int grp = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;
    float bias[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bias[c] = c < p.Cout ? p.bias[c] : 0.f;
    for (int it = grp;; it += 2) {
      const long long unit_ll = (long long)unit_id + (long long)it * unit_cnt;
      if (unit_ll >= units) break;
      int n, y0, x0;
      tile_of((int)unit_ll, n, y0, x0);
      const int acs = it & 1;
      MBAR_WAIT_RELAXED(t_full(acs), (it >> 1) & 1, 870 + acs);
      tc_fence_after();
      uint32_t vh[32], vl[32];  // partial sums of the a_hi and the a_lo slab
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acs * 64);
      tmem_ld32(taddr, vh);
      tmem_ld32(taddr + 32, vl);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(lead(t_empty(acs)));
        else mbar_arrive(t_empty(acs));
      }
      const int y = y0 + quad, x = x0 + lane;
      const bool ok = lane < kROutW && y < p.H && x < p.W && n < p.N;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float q[3];
#pragma unroll
        for (int sc = 0; sc < 3; ++sc) {
          const int i = sc * 4 + c;  // columns i: * w_hi, 16 + i: * w_lo
          q[sc] = ((__uint_as_float(vh[16 + i]) + __uint_as_float(vl[i])) + __uint_as_float(vl[16 + i])) +
                  __uint_as_float(vh[i]);
        }
        const float p1 = __shfl_down_sync(0xffffffffu, q[1], 1);
        const float p2 = __shfl_down_sync(0xffffffffu, q[2], 2);
        float o = fmaf((q[0] + p1) + p2, p.out_scale, bias[c]);
        if (p.relu) o = fmaxf(o, 0.f);
        if (ok && c < p.Cout) {
          if (p.out_u8) p.out_u8[(((size_t)n * p.H + y) * p.W + x) * p.Cout + c] = quantize_u8(o);
          else p.out_nchw[(((size_t)n * p.Cout + c) * p.H + y) * p.W + x] = o;
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc_cg<CG, 128>(tmem_base);
}

// in: [hi | lo] map (in.C == 128); p carries bias, out_scale, out_nchw / out_u8, relu, Cout <= 3
template <typename T16>
int launch_last_rows_x3(ActView<T16> in, const T16* wk_x3, ConvParams<T16> p, cudaStream_t st) {
  constexpr int CG = 2;
  CUtensorMap mr;
  if (int e = make_act_map(&mr, in, kRBoxW, kRRows + 2)) return e;
  auto kernel = conv_last_rows_x3_kernel<T16, CG>;
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(kernel), kR3Smem));
  p.tiles_x = (in.W + kROutW - 1) / kROutW;
  p.tiles_y = (in.H + kRRows - 1) / kRRows;
  const int64_t tiles = (int64_t)in.N * p.tiles_x * p.tiles_y;
  CCST_CHECK_ARG(tiles < (1ll << 31) - 2, "conv_last_rows_x3: too many tiles");
  p.m_tiles = p.total_tiles = (int)tiles;
  const int64_t units = (tiles + CG - 1) / CG;
  const int slots = sm_count() / CG;
  const int grid = (int)(units < slots ? units : slots) * CG;
  CCST_CUDA(launch_conv(kernel, grid, kThreadsUmma, kR3Smem, st, CG, mr, wk_x3, p));
  CCST_LAUNCHED();
  return CCST_OK;
}

}  // namespace
}  // namespace ccst
