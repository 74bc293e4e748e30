// x3 engines ("fp16x3", "bf16x3"): 3x3 reflect-pad convolution with SPLIT 16-bit operands on the tensor pipe.
//
// The reference computes relu4_1 in fp32 (net.py:38-69 under torch defaults) and BASELINE.json asks the style
// statistics to match it to 1e-5 relative; one f16 rounding per layer is 5e-4.  Here every activation v is
// stored as hi = f16(v) and lo = f16(v - hi) in the channel ranges [0, C) and [C, 2C) of its map, every weight
// w (times an exact power of two that keeps its low part out of the f16 subnormals) as w_hi + w_lo, and
//   a * w = a_hi * w_hi + a_hi * w_lo + a_lo * w_hi + a_lo * w_lo
// runs as ONE tcgen05 implicit GEMM per 64-channel output tile: K = 9 x 2*Cin over the [a_hi | a_lo] channels,
// N = 128 = [w_hi rows | w_lo rows] of the tile (N = 64 would be capped at half the tensor rate by the
// A-operand fetch; the fourth, 2^-22-sized term comes for free).  Operand precision: 22 significand bits.
//
// Accumulation.  tcgen05.mma adds into its fp32 TMEM accumulator with truncation, not round-to-nearest
// (measured: accumulating a whole K = 9 * 3 * Cin reduction in TMEM leaves every positive output 1e-5 low per
// layer, 7e-5 on relu4_1 -- the effect Ootomo & Yokota describe for earlier tensor cores).  So the
// accumulator is PROMOTED: the 12 MMAs of one (channel chunk, filter column) step write a fresh partial sum
// into one of four 128-column TMEM buffers, and the epilogue warps add the partials in registers with
// round-to-nearest fp32 adds while the next steps' MMAs run (buffers 2g, 2g+1 belong to epilogue group g,
// which owns the tiles of parity g).  Truncation then only acts inside a 12-MMA partial.  (Two steps = 24 MMAs per
// partial were measured: 1.6 % faster, relu4_1 statistics 2.2e-6 -> 4.3e-6 relative, encoder max-abs 1.2e-5 ->
// 2.2e-5 -- not taken; profiles/r03n_ab_x3_two_steps_per_partial.txt.  ncu: tensor pipe 74-80 % active on the
// large layers, profiles/r03m_ncu_x3_summary.txt.)
//
// Epilogue: v = acc * 2^-k + bias, ReLU, optional 2x2 ceil-mode max-pool (all fp32), split into hi / lo and
// stored as two TMA tiles at channels co and Cout + co (EPI_ACT, EPI_ACT_POOL; EPI_ACT_UP2 stores every pixel
// at its four nearest-x2 replicas), or written as fp32 NCHW / save_image-quantised uint8 NHWC (EPI_NCHW_F32:
// the last decoder conv, whose <= 3 output channels ride in a zero-padded 64-channel tile).
//
// With T16 = __nv_bfloat16 the same kernel gives 16 significand bits at the fp32 exponent range ("bf16x3"):
// the tensor-core path for weights whose activations leave the f16 range, meeting the image bar that single
// bf16 operands miss.
#pragma once
#include "conv_last.cuh"
#include "conv_main.cuh"

namespace ccst {
namespace {

template <int CG>
struct X3Cfg {
  static constexpr int kN = 128;
  static constexpr int kBRows = kN / CG;
  static constexpr int kBBytes = kBRows * kBlockK * 2;  // 16 KiB / CG
  static constexpr int kAStages = CG == 2 ? 5 : 4;
  static constexpr int kBStages = 6;  // two groups of three filter-row tiles
  static constexpr int kAOff = 0;
  static constexpr int kBOff = kAStages * kASlabBytes;
  static constexpr int kStoreOff = kBOff + kBStages * kBBytes;
  static constexpr int kBiasOff = kStoreOff + 2 * kStoreBytes;
  static constexpr int kBiasBytes = 2048;  // up to 512 fp32 biases
  static constexpr int kBarOff = kBiasOff + kBiasBytes;
  static constexpr int kNumBars = 2 * kAStages + 2 * kBStages + 8;
  static constexpr int kTmemCols = 512;  // 4 partial buffers x 128 columns
  static constexpr int kSmemBytes = kBarOff + 8 * kNumBars + 16 + 1024;
  static_assert(kSmemBytes <= 232448, "shared memory plan exceeds 227 KiB");
};

template <typename T16, int EPI, int CG>
__global__ void __launch_bounds__(kThreadsUmma, 1)
    conv_x3_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ OutMaps tmap_out, ConvParams<T16> p) {
  static_assert(EPI == EPI_ACT || EPI == EPI_ACT_POOL || EPI == EPI_ACT_UP2 || EPI == EPI_NCHW_F32 || EPI == EPI_UPS,
                "x3 epilogues");
  // EPI_UPS: nearest-x2 upsample folded into this conv as four 2x2 phase convolutions over the low-resolution
  // input (conv_main.cuh): unit = (pixel tile, phase), 2 x 2 taps, phase weights split like the rest
  constexpr bool UPS = (EPI == EPI_UPS);
  constexpr int kTR = UPS ? 2 : 3, kTS = UPS ? 2 : 3;
  using Cfg = X3Cfg<CG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t store_base = smem_base + Cfg::kStoreOff;
  const uint32_t bar_base = smem_base + Cfg::kBarOff;
  float* s_bias = reinterpret_cast<float*>(smem_gen + Cfg::kBiasOff);
  auto a_smem = [&](int s) { return smem_base + Cfg::kAOff + s * kASlabBytes; };
  auto b_smem = [&](int s) { return smem_base + Cfg::kBOff + s * Cfg::kBBytes; };
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (Cfg::kAStages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * Cfg::kAStages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * Cfg::kAStages + Cfg::kBStages + s); };
  constexpr int kBar2 = 2 * Cfg::kAStages + 2 * Cfg::kBStages;
  auto part_full = [&](int b) { return bar_base + 8u * (kBar2 + b); };       // partial buffer b written
  auto part_empty = [&](int b) { return bar_base + 8u * (kBar2 + 4 + b); };  // ... and read out again
  const uint32_t tmem_slot = bar_base + 8u * Cfg::kNumBars;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int klog = p.Cin / kBlockK;  // 64-channel chunks of the logical input
  const int kchunks = 2 * klog;      // [hi | lo]
  const int steps = kchunks * kTS;   // (chunk, filter column) steps per tile: even
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  auto lead = [&](uint32_t bar) { return CG == 2 ? mapa_rank(bar, 0) : bar; };
  const int unit_id = CG == 2 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int unit_cnt = CG == 2 ? (int)ncluster_x() : (int)gridDim.x;

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
    for (int b = 0; b < 4; ++b) {
      mbar_init(part_full(b), 1);
      mbar_init(part_empty(b), 4 * CG);  // one arrive per warp of the owning epilogue group (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg<CG, Cfg::kTmemCols>(tmem_slot);
  for (int i = threadIdx.x; i < p.n_tiles * 64; i += kThreadsUmma) s_bias[i] = i < p.Cout ? p.bias[i] : 0.f;
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kBarOff + 8 * Cfg::kNumBars);

  if (warp == 0) {
    // ===================== TMA producer
    const int b_row0 = (int)cta_rank * Cfg::kBRows;
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int unit = unit_id; unit < p.total_tiles; unit += unit_cnt) {
      const TileCoord t = decode_tile<CG, UPS>(p, unit, (int)cta_rank);
      const int xs0 = t.x0 + (UPS ? (t.ph & 1) : 0);
      const int b_row = (UPS ? t.ph * p.n_tiles * Cfg::kN : 0) + t.nt * Cfg::kN + b_row0;
      for (int kc = 0; kc < kchunks; ++kc) {
        const int kb = kc >= klog ? kc - klog : kc;  // the lo chunks meet the same weights as the hi chunks
        for (int s = 0; s < kTS; ++s) {
          MBAR_WAIT_RELAXED(a_empty(as), aph ^ 1, 700 + as);
          if (elect_one()) {
            if (leader) mbar_expect_tx(a_full(as), CG * kASlabBytes);
            tma_load_4d_cg<CG>(a_smem(as), &tmap_a, lead(a_full(as)), kc * kBlockK, xs0 + s, t.y0, t.n);
          }
          __syncwarp();
          if (++as == Cfg::kAStages) as = 0, aph ^= 1;
          MBAR_WAIT_RELAXED(b_empty(bs), bph ^ 1, 720 + bs);
          if (elect_one()) {
            if (leader) mbar_expect_tx(b_full(bs), CG * kTR * Cfg::kBBytes);
            const uint32_t bar = lead(b_full(bs));
#pragma unroll
            for (int r = 0; r < kTR; ++r)
              tma_load_2d_cg<CG>(b_smem(bs + r), &tmap_b, bar, (r * kTS + s) * p.Cin + kb * kBlockK, b_row);
          }
          __syncwarp();
          if ((bs += kTR) == Cfg::kBStages) bs = 0, bph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: every step starts a fresh partial sum
    if (leader) {
      constexpr uint32_t idesc = make_idesc<T16, Cfg::kN, CG>();
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (int unit = unit_id; unit < p.total_tiles; unit += unit_cnt, ++it) {
        const uint32_t use0 = (uint32_t)(it >> 1) * (uint32_t)(steps >> 1);  // uses of this tile parity's buffers so far
        // UPS: row phase a of this unit shifts the slab rows of the two taps to a + r
        const int row_shift = UPS ? (((unit / p.n_tiles) & 3) >> 1) : 0;
        int j = 0;
        for (int kc = 0; kc < kchunks; ++kc) {
#pragma unroll
          for (int s = 0; s < kTS; ++s, ++j) {
            const int pb = 2 * (it & 1) + (j & 1);
            const uint32_t use = use0 + (uint32_t)(j >> 1);
            mbar_wait(a_full(as), aph, 740 + as);
            mbar_wait(b_full(bs), bph, 750 + bs);
            if (CG == 2) mbar_wait_cluster(part_empty(pb), (use & 1u) ^ 1u, 760 + pb);
            else mbar_wait(part_empty(pb), (use & 1u) ^ 1u, 760 + pb);
            tc_fence_after();
            const int bs0 = bs;
            if ((bs += kTR) == Cfg::kBStages) bs = 0, bph ^= 1;
            if (elect_one()) {
              const uint32_t tmem_d = tmem_base + (uint32_t)(pb * Cfg::kN);
              const uint64_t adesc0 = make_kmajor_sw128_desc(a_smem(as) + (uint32_t)(row_shift * kTileW) * 128u);
#pragma unroll
              for (int r = 0; r < kTR; ++r) {
                const uint64_t adesc = adesc0 + (uint64_t)(r * (kTileW * 128 >> 4));
                const uint64_t bdesc = make_kmajor_sw128_desc(b_smem(bs0 + r));
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_f16_cg<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (r | k) ? 1u : 0u);
              }
              umma_commit_cg<CG>(b_empty(bs0));
              umma_commit_cg<CG>(a_empty(as));
              umma_commit_cg<CG>(part_full(pb));
            }
            __syncwarp();
            if (++as == Cfg::kAStages) as = 0, aph ^= 1;
          }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue group g: tiles of parity g, partial buffers 2g and 2g + 1
    const int grp = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int py = row / kTileW, px = row % kTileW;
    const bool issuer_warp = (quad == 0);
    const uint32_t sbuf = store_base + grp * kStoreBytes;
    SatTracker<T16> sat;
    for (int it = grp;; it += 2) {
      const long long unit_ll = (long long)unit_id + (long long)it * unit_cnt;
      if (unit_ll >= p.total_tiles) break;
      const TileCoord t = decode_tile<CG, UPS>(p, (int)unit_ll, (int)cta_rank);
      const int y = t.y0 + py, x = t.x0 + px;
      const bool valid = (y < p.H) && (x < p.W) && (CG == 1 || t.n < p.N);
      const uint32_t use0 = (uint32_t)(it >> 1) * (uint32_t)(steps >> 1);
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
#pragma unroll 1
      for (int j = 0; j < steps; ++j) {
        const int pb = 2 * grp + (j & 1);
        const uint32_t use = use0 + (uint32_t)(j >> 1);
        mbar_wait(part_full(pb), use & 1u, 780 + pb);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(pb * Cfg::kN);
        {
          uint32_t u[32], v[32];
          tmem_ld32(taddr, u);       // a * w_hi, channels 0..31
          tmem_ld32(taddr + 64, v);  // a * w_lo
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = (acc[i] + __uint_as_float(u[i])) + __uint_as_float(v[i]);
        }
        {
          uint32_t u[32], v[32];
          tmem_ld32(taddr + 32, u);
          tmem_ld32(taddr + 96, v);
          tmem_ld_wait();
          // the buffer is in registers: hand it back to the MMA warp before the adds
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(lead(part_empty(pb)));
            else mbar_arrive(part_empty(pb));
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[32 + i] = (acc[32 + i] + __uint_as_float(u[i])) + __uint_as_float(v[i]);
        }
      }
      // ---- v = acc * 2^-k + bias, ReLU, pooling: all in fp32
      const int co = t.nt * 64;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float v = fmaf(acc[i], p.out_scale, s_bias[co + i]);
        if (p.relu) v = fmaxf(v, 0.f);
        if (EPI == EPI_ACT_POOL) v = valid ? v : 0.f;  // outside the image: 0, the identity for post-ReLU values
        acc[i] = v;
      }
      if (EPI == EPI_ACT_POOL) {
        // 2x2 window = lanes {l, l^1, l^16, l^17}
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          float tt[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) tt[i] = __shfl_xor_sync(0xffffffffu, acc[h2 * 32 + i], 1);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[h2 * 32 + i] = fmaxf(acc[h2 * 32 + i], tt[i]);
#pragma unroll
          for (int i = 0; i < 32; ++i) tt[i] = __shfl_xor_sync(0xffffffffu, acc[h2 * 32 + i], 16);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[h2 * 32 + i] = fmaxf(acc[h2 * 32 + i], tt[i]);
        }
      }
      if (EPI == EPI_NCHW_F32) {
        // last conv: the first p.Cout (<= 3) channels of the tile go to the caller's tensor, fp32 NCHW or
        // quantised like save_image as uint8 NHWC; a tile row = 16 consecutive pixels of one image row
        if (valid && t.nt == 0) {
          if (p.out_u8 != nullptr) {
            uint8_t* dst = p.out_u8 + (((size_t)t.n * p.H + y) * p.W + x) * p.Cout;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c < p.Cout) dst[c] = quantize_u8(acc[c]);
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c < p.Cout) p.out_nchw[(((size_t)t.n * p.Cout + c) * p.H + y) * p.W + x] = acc[c];
          }
        }
        continue;
      }
      int srow = row;
      bool writer = true;
      if (EPI == EPI_ACT_POOL) {
        writer = !(lane & 1) && lane < 16;  // anchor of a 2x2 window
        srow = (py >> 1) * (kTileW / 2) + (px >> 1);
      }
      // ---- the two halves, one after the other through the group's staging tile
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v0 = acc[2 * i], v1 = acc[2 * i + 1];
          const uint32_t hi = pack16x2<T16>(v0, v1);
          if (part == 0) {
            pk[i] = hi;
          } else {
            const float2 hf = unpack16x2<T16>(hi);
            pk[i] = pack16x2<T16>(v0 - hf.x, v1 - hf.y);
          }
        }
        if (part == 0) sat.track_block(pk, p.relu != 0);
        const int cs = co + part * p.Cout;  // channel of this half in the [hi | lo] map
        if (issuer_warp) bulk_wait_read<0>();
        epi_barrier(grp);
        if (writer) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t dst = sbuf + srow * 128 + ((q ^ (srow & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                         "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                         : "memory");
          }
        }
        if (valid) {
          if (EPI == EPI_ACT) {
            store_aliases(p.out, t.n, y, x, cs, pk, p.halo_edge);
          } else if (EPI == EPI_UPS) {
            store_aliases(p.out, t.n, 2 * y + (t.ph >> 1), 2 * x + (t.ph & 1), cs, pk, 1);
          } else if (EPI == EPI_ACT_UP2) {
#pragma unroll
            for (int aa = 0; aa < 2; ++aa)
#pragma unroll
              for (int bb = 0; bb < 2; ++bb) store_aliases(p.out, t.n, 2 * y + aa, 2 * x + bb, cs, pk, 1);
          } else if (writer) {
            store_aliases(p.out, t.n, y >> 1, x >> 1, cs, pk, 1);
          }
        }
        fence_async_smem();
        epi_barrier(grp);
        if (issuer_warp && elect_one()) {
          if (EPI == EPI_ACT_POOL) {
            tma_store_4d(&tmap_out.m[0], sbuf, cs, t.x0 >> 1, t.y0 >> 1, t.n);
          } else if (EPI == EPI_UPS) {
            tma_store_4d(&tmap_out.m[t.ph], sbuf, cs, t.x0, t.y0, t.n);
          } else {
            tma_store_4d(&tmap_out.m[0], sbuf, cs, t.x0, t.y0, t.n);
            if (EPI == EPI_ACT_UP2) {
              tma_store_4d(&tmap_out.m[1], sbuf, cs, t.x0, t.y0, t.n);
              tma_store_4d(&tmap_out.m[2], sbuf, cs, t.x0, t.y0, t.n);
              tma_store_4d(&tmap_out.m[3], sbuf, cs, t.x0, t.y0, t.n);
            }
          }
          bulk_commit();
        }
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

template <typename T16, int EPI, int CG>
int launch_x3_cfg(const CUtensorMap& ma, const T16* wk_x3, ConvParams<T16> p, cudaStream_t st) {
  using Cfg = X3Cfg<CG>;
  constexpr bool UPS = (EPI == EPI_UPS);
  CUtensorMap mb;
  // [n_tiles * 128 rows = (tile, hi | lo, co)][9 * Cin]; UPS: [4 phases x n_tiles * 128 rows][4 * Cin]
  if (int e = make_weight_map(&mb, wk_x3, (UPS ? 4 : 9) * p.Cin, (UPS ? 4 : 1) * p.n_tiles * Cfg::kN, Cfg::kBRows)) return e;
  OutMaps mo;
  memset(&mo, 0, sizeof(mo));
  if (EPI == EPI_ACT) {
    if (int e = make_out_map(&mo.m[0], p.out, 0, 0, 1, 1, kTileW, kTileH)) return e;
  } else if (EPI == EPI_ACT_POOL) {
    if (int e = make_out_map(&mo.m[0], p.out, 0, 0, 1, 1, kTileW / 2, kTileH / 2)) return e;
  } else if (EPI == EPI_ACT_UP2 || EPI == EPI_UPS) {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        if (int e = make_out_map(&mo.m[a * 2 + b], p.out, a, b, 2, 2, kTileW, kTileH)) return e;
  }
  auto kernel = conv_x3_kernel<T16, EPI, CG>;
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(kernel), Cfg::kSmemBytes));
  const int64_t units = ((int64_t)p.m_tiles + CG - 1) / CG * p.n_tiles * (UPS ? 4 : 1);
  CCST_CHECK_ARG(units < (1ll << 31), "conv_x3: too many tiles");
  p.total_tiles = (int)units;
  const int slots = sm_count() / CG;
  const int grid = (int)(units < slots ? units : slots) * CG;
  CCST_CUDA(launch_conv(kernel, grid, kThreadsUmma, Cfg::kSmemBytes, st, CG, ma, mb, mo, p));
  CCST_LAUNCHED();
  return CCST_OK;
}

// in: [hi | lo] map of in.C / 2 logical channels; out: [hi | lo] map of Cout logical channels (or the caller's
// NCHW fp32 / NHWC uint8 tensor for EPI_NCHW_F32)
template <typename T16>
int launch_x3(const UmmaConvArgs<T16>& a, ConvParams<T16> p, cudaStream_t st) {
  const ActView<T16>& in = a.in;
  const bool last = a.epi == EPI_NCHW_F32, ups = a.epi == EPI_UPS;
  CCST_CHECK_ARG((a.epi == EPI_ACT || a.epi == EPI_ACT_POOL || a.epi == EPI_ACT_UP2 || last || ups) &&
                     (a.halo_edge == 1 || a.epi == EPI_ACT) && !a.per_sample && a.wk_x3 != nullptr,
                 "conv_x3: epilogue %d is not available on the x3 engines", a.epi);
  CCST_CHECK_ARG(!ups || (a.wk_x3_up != nullptr && a.out.H == 2 * in.H && a.out.W == 2 * in.W),
                 "conv_x3: EPI_UPS needs the split phase weights and an output twice the input's size");
  CCST_CHECK_ARG(in.C % (2 * kBlockK) == 0, "conv_x3: Cin=%d/2 must be a multiple of 64", in.C);
  if (last) {
    CCST_CHECK_ARG(a.Cout >= 1 && a.Cout <= 4 && (a.out_nchw != nullptr || a.out_u8 != nullptr),
                   "conv_x3: the NCHW epilogue is the (<= 4)-channel last conv");
  } else {
    CCST_CHECK_ARG(a.Cout % 64 == 0 && a.Cout <= 512 && a.out.C == 2 * a.Cout,
                   "conv_x3: Cout=%d must be a multiple of 64 (<= 512)", a.Cout);
  }
  p.Cin = in.C / 2;
  p.out_scale = a.out_scale;
  // the decoder's 64 -> 3 last conv: filter columns in N, two slabs per tile (conv_last_rows_x3_kernel)
  if (last && a.Cout <= 3 && in.C == 2 * kBlockK) return launch_last_rows_x3<T16>(in, a.wk_x3, p, st);
  p.n_tiles = (a.Cout + 63) / 64;
  const int64_t mt = (int64_t)in.N * p.tiles_x * p.tiles_y;
  CCST_CHECK_ARG(mt * p.n_tiles < (1ll << 30), "conv_x3: too many tiles");
  p.m_tiles = (int)mt;
  CUtensorMap ma;
  if (int e = make_act_map(&ma, in)) return e;
  switch (a.epi) {
    case EPI_ACT: return launch_x3_cfg<T16, EPI_ACT, 2>(ma, a.wk_x3, p, st);
    case EPI_ACT_POOL: return launch_x3_cfg<T16, EPI_ACT_POOL, 2>(ma, a.wk_x3, p, st);
    case EPI_ACT_UP2: return launch_x3_cfg<T16, EPI_ACT_UP2, 2>(ma, a.wk_x3, p, st);
    case EPI_UPS: return launch_x3_cfg<T16, EPI_UPS, 2>(ma, a.wk_x3_up, p, st);
    default: return launch_x3_cfg<T16, EPI_NCHW_F32, 2>(ma, a.wk_x3, p, st);
  }
}

}  // namespace
}  // namespace ccst
