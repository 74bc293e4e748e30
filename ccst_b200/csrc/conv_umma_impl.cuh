// 3x3 reflect-pad convolution as a tcgen05 / TMEM implicit GEMM fed by TMA (sm_100a).
//
// Replaces the 17 `ReflectionPad2d(1) + Conv2d(3x3) (+ReLU)` stages of net.py:6-36 / :40-69 (all
// convolutions except conv1_1, whose K = 27 is handled by conv_first), with the nearest x2
// `Upsample` (net.py:10,23,30) and the ceil-mode `MaxPool2d` (net.py:46,53,66) fused into the store.
//
// GEMM view   D[M = pixels, N = Cout] = A[M, K = 9*Cin] * B[K, N]
//   A  is never materialised: activations live in HBM as NHWC bf16 with a one-pixel reflection halo,
//      so the A tile of filter tap (r,s) and channel chunk c0 for the output tile (n, y0..y0+7,
//      x0..x0+15) is the plain TMA box {64 ch, 16 px, 8 rows, 1 img} at (c0, x0+s, y0+r, n).  The box
//      lands in shared memory as 128 rows x 128 bytes with the 128-byte swizzle, which is exactly the
//      K-major SWIZZLE_128B operand layout of tcgen05.mma.
//   B  = weights packed [CoutPad][9*Cin] bf16 (K contiguous, k = tap*Cin + c), TMA box {64, BN}.
//   D  accumulates in TMEM (fp32), two accumulator stages of BN columns so the epilogue of tile i
//      overlaps the MMAs of tile i+1.
//
// Warp roles (256 threads, one CTA per SM, persistent over tiles):
//   warp 0 lane 0 : TMA producer          warp 1 lane 0 : tcgen05.mma issuer
//   warp 2        : TMEM alloc / dealloc  warps 4..7    : epilogue (TMEM -> regs -> bias/ReLU -> HBM)
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "layers.h"

namespace ccst {

namespace {

template <typename T16>
struct Fmt16;  // operand format code of the kind::f16 instruction descriptor + TMA data type
template <>
struct Fmt16<__nv_bfloat16> {
  static constexpr uint32_t kIdescFmt = 1;  // BF16
  static constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
};
template <>
struct Fmt16<__half> {
  static constexpr uint32_t kIdescFmt = 0;  // F16
  static constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
};

constexpr int kTileH = 8, kTileW = 16, kBlockM = kTileH * kTileW, kBlockK = 64;
constexpr int kThreadsUmma = 384;  // 4 control warps + 2 epilogue groups of 4 warps
constexpr int kEpiWarp0 = 4;

// Shared-memory plan of the main kernel.
//   A ring : kAStages slabs of {64 ch, 16 px, 10 rows} = 20 KiB.  One slab serves the three filter
//            rows r = 0,1,2 of one (channel chunk, filter column s): the operand of tap (r,s) is the
//            slab shifted by r tile rows = r * 2048 bytes, which keeps the 1024-byte swizzle phase.
//            (A bytes per tile: 3 slabs instead of 9 tiles per channel chunk -> 2.4x less L2->SM traffic.)
//   B      : BRES = false: ring of kBStages weight tiles {64 k, BN}, one per tap and channel chunk;
//            BRES = number of 64-channel input chunks whose weights are RESIDENT (0 = stream): when
//            taps * Cin * BN * 2 bytes fit (Cin = 64 at N <= 128, the phase weights of the Cin = 128
//            upsample-fused layer) they are loaded once per CTA and never re-fetched.
//   store  : kStoreBufs x 16 KiB staging tiles for the TMA store of the epilogue.
// Tile geometry of the main kernel.
//   GEO 0: 8 x 16 output pixels; one slab {64 ch, 16 px, 10 rows} PER FILTER COLUMN s (the operand of
//          tap (r, s) is slab s shifted by r rows = r * 2048 B).
//   GEO 1/2 ("linear slab"): ONE slab per channel chunk serves all taps.  The slab is read as a linear
//          run of pixels with row pitch kBoxW: accumulator row i <-> slab position i, and the operand of
//          tap (r, s) is the same slab starting (r * kBoxW + s) * 128 B later (the 128-byte swizzle is a
//          function of the shared-memory address, so any 128-byte-multiple start keeps the pattern).
//          Positions whose column is >= kBoxW - 2 wrap into the next row and are discarded: 14 of 16
//          (GEO 1, 8 rows) or 30 of 32 (GEO 2, 4 rows) accumulator rows are outputs.  3x less L2 -> SM
//          and TMA -> shared-memory traffic for the activations, paid with 12.5 % / 6.25 % idle MMA rows.
template <int GEO>
struct Geo {
  static constexpr bool kLin = GEO != 0;
  static constexpr int kBoxW = GEO == 2 ? 32 : 16;
  static constexpr int kRows = kBlockM / kBoxW;             // output rows per tile
  static constexpr int kOutW = kLin ? kBoxW - 2 : kBoxW;    // output columns per tile
  static constexpr int kSlabRows = kRows + 2;
  static constexpr int kSlabBytes = kSlabRows * kBoxW * 128;  // 20480 / 24576
};

template <int BN, int BRES, int CG, bool UPS = false, int GEO = 0>
struct UmmaCfg {
  static constexpr int kSlabRows = Geo<GEO>::kSlabRows;
  static constexpr int kASlabBytes = Geo<GEO>::kSlabBytes;
  // CG = 2 (CTA pair, tcgen05 cta_group::2): the pair computes M = 256 pixels x BN channels per MMA;
  // each CTA stages the A slab of its own 128-pixel tile and HALF of the weight tile (BN/2 rows),
  // so the per-SM shared-memory traffic of the B operand (TMA writes and tensor-core reads) halves.
  static constexpr int kBRows = BN / CG;
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kBStride = (kBBytes + 1023) / 1024 * 1024;
  static constexpr int kTaps = UPS ? 4 : 9;  // UPS: 2x2 phase convolution (see EPI_UPS)
  // (a linear slab carries a whole channel chunk -- all taps -- so fewer stages cover the same work)
  static constexpr int kAStages =
      GEO != 0 ? (BN >= 128 ? 3 : 4)
      : CG == 2 ? (BRES ? (BN >= 128 && !UPS ? 5 : 6) : (BN >= 256 ? 4 : 5))
                : (BRES ? (BN >= 128 ? 3 : (BN >= 64 ? (UPS ? 6 : 5) : 6)) : (BN >= 256 ? 3 : 4));
  // BRES = false: the weight tiles of the kTR filter rows of one (chunk, filter column) step travel
  // as ONE group -- one full/empty barrier pair, one wait per step in the producer and in the MMA
  // warp (a wait + elect + issue round per single tile costs ~300 cycles of serial scalar code in
  // each of those warps, more than the 256 cycles of math a tile feeds at N = 128).
  static constexpr int kBGroup = BRES ? 1 : (UPS ? 2 : 3);
  static constexpr int kBStagesRaw = BRES ? kTaps * BRES /* resident: all taps x BRES chunks */
                                     : CG == 2 ? (BN >= 256 ? 6 : 9)
                                               : (BN >= 256 ? 4 : (BN >= 128 ? 6 : 9));
  static constexpr int kBStages = kBStagesRaw / kBGroup * kBGroup;
  static constexpr int kStoreBufs = 2;  // one staging tile per epilogue group
  static constexpr int kStoreStageBytes = (BN >= 64) ? kStoreBufs * kBlockM * 128 : 0;
  static constexpr int kBiasBytes = 2048;  // up to 512 fp32 biases
  static constexpr int kAOff = 0;
  static constexpr int kBOff = kAStages * kASlabBytes;
  static constexpr int kStoreOff = kBOff + kBStages * kBStride;
  static constexpr int kBiasOff = kStoreOff + kStoreStageBytes;
  static constexpr int kBarOff = kBiasOff + kBiasBytes;
  static constexpr int kNumBars = 2 * kAStages + 2 * kBStages + 4 + 1;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kBarOff + 8 * kNumBars + 16 + 1024 /*align slack*/;
  static constexpr bool kFits = kSmemBytes <= 232448;  // 227 KiB; launch_cfg refuses plans that do not fit
  static_assert(CG == 1 || (BN >= 32 && BN % 32 == 0), "cta_group::2 needs N % 32 == 0");
};

template <typename T16>
struct ConvParams {
  int N, H, W, Cin;
  int Cout, CoutPad;
  int tiles_x, tiles_y, n_tiles, m_tiles;  // m_tiles = pixel tiles (N * tiles_y * tiles_x)
  int total_tiles;                        // work units: (pixel tile | pair of pixel tiles) x n_tiles
  const T16* in_ptr;  // host side only (tensor maps of the non-default tile geometries)
  int relu;
  int desc_mode;  // linear slabs: see make_kmajor_sw128_desc_off
  int ablate;     // measurement only (CCST_ABLATE): 1 skip the epilogue's work, 2 skip the MMAs, 4 skip the A loads
  int halo_edge;  // halo written around `out`: 1 reflection, 0 replicate (for_each_halo_alias)
  const float* bias;
  ActView<T16> out;
  float* out_nchw;
  uint8_t* out_u8;  // last conv only: NHWC uint8 store quantised like torchvision's save_image
  float2* tile_stats;  // EPI_ACT_STATS: [(pixel tile * 4 + row quarter) * Cout + channel] {mean, M2}
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
#ifndef CCST_RELAXED_WAITS
#define CCST_RELAXED_WAITS 1
#endif
#if CCST_RELAXED_WAITS
#define MBAR_WAIT_RELAXED mbar_wait_relaxed
#else
#define MBAR_WAIT_RELAXED mbar_wait
#endif
// The same for the waits that are NOT on the tensor pipe's critical path (producer waiting for a free
// stage, epilogue waiting for an accumulator): back off between polls instead of spinning, so the
// pollers leave the issue slots (and the power budget -- long runs are power-capped) to the MMA warp.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (clock64() - t0 > 4000000000LL) {
      printf("ccst conv_umma: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag,
             (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  }
}
// Bounded wait: a pipeline bug must trap (reported as a CUDA error), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("ccst conv_umma: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag,
             (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
template <typename T16>
__device__ __forceinline__ float2 unpack16x2(uint32_t w);
template <>
__device__ __forceinline__ float2 unpack16x2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <>
__device__ __forceinline__ float2 unpack16x2<__nv_bfloat16>(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

// Programmatic dependent launch: the conv kernels of a step are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the CTAs of layer i+1 may be scheduled on an
// SM as soon as layer i's CTA there has exited and run their prologue (barrier init, TMEM allocation,
// tensor-map prefetch, resident-weight loads) under layer i's tail.  pdl_wait() returns once the
// preceding kernel has completed and its memory is visible; nothing produced or still read by that
// kernel (activations in, activations out) is touched before it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants.  CG = 1 forwards to the single-CTA forms above.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t ncluster_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// (default .release.cta semantics: these arrivals hand over TMEM / shared-memory stages whose accesses
// are ordered by tcgen05 fences and wait::ld, no global data is published through them -- the
// .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR + an L1 invalidate per arrival, which cost
// the pair kernels ~15 % of their epilogue time)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait on a barrier that CTAs of the whole cluster arrive on
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("ccst conv_umma: cluster mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag,
             (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  }
}
// TMA loads whose completion bytes are credited to `bar`, a shared::cluster address that may belong
// to the peer CTA of the pair (the leader's "full" barrier counts the bytes of both CTAs)
template <int CG>
__device__ __forceinline__ void tma_load_4d_cg(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                               int c0, int c1, int c2, int c3) {
  if (CG == 1) {
    tma_load_4d(dst, map, bar, c0, c1, c2, c3);
  } else {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tma_load_2d_cg(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                               int c0, int c1) {
  if (CG == 1) {
    tma_load_2d(dst, map, bar, c0, c1);
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  }
}
template <int CG, int COLS>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t smem_dst) {
  if (CG == 1) {
    tmem_alloc<COLS>(smem_dst);
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG, int COLS>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr) {
  if (CG == 1) {
    tmem_dealloc<COLS>(taddr);
  } else {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
                 : "memory");
  }
}
// CG = 2: ONE thread of the leader CTA issues the MMA for the pair: D[256 x N] lives in the TMEM of
// both CTAs (128 lanes each), A = each CTA's own slab, B = the two N-halves held by the two CTAs.
template <int CG>
__device__ __forceinline__ void umma_f16_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  if (CG == 1) {
    umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// CG = 2: the arrive is multicast to the barrier at the same offset in both CTAs of the pair
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint32_t bar) {
  if (CG == 1) {
    umma_commit(bar);
  } else {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
            "r"(bar),
        "h"((uint16_t)3)
        : "memory");
  }
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major) |
//   [32,46) SBO >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// Same for a start address that is a 128-byte multiple but not 1024-byte aligned (linear slabs).
// mode 0: address only; mode 1: also the descriptor's base-offset field [49,52) = (addr >> 7) & 7.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc_off(uint32_t smem_addr, int mode) {
  uint64_t d = make_kmajor_sw128_desc(smem_addr);
  if (mode == 1) d |= (uint64_t)((smem_addr >> 7) & 7u) << 49;
  return d;
}
// kind::f16 instruction descriptor: D = fp32, A = B = bf16 or f16, both K-major, M = 128, N = BN
template <typename T16, int BN, int CG = 1>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) /*D fp32*/ | (Fmt16<T16>::kIdescFmt << 7) /*A*/ | (Fmt16<T16>::kIdescFmt << 10) /*B*/ |
         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((CG * kBlockM) >> 4) << 24);
}

struct TileCoord {
  int n, y0, x0, nt, ph;
};
// work unit -> (N tile, pixel tile of CTA `rank` of the pair).  A pair takes two consecutive pixel
// tiles; when the number of pixel tiles is odd the last pair's second tile is a dummy at image
// index n = N: its TMA loads are out of bounds (zero fill) and its stores are clipped away.
// UPS: the four output phases (a, b) of one pixel tile are consecutive units, so the CTAs that run
// them concurrently share the tile's input slabs in L2.
template <int CG, bool UPS = false, int GEO = 0, typename P>
__device__ __forceinline__ TileCoord decode_tile(const P& p, int unit, int rank) {
  TileCoord t;
  t.nt = unit % p.n_tiles;
  int u = unit / p.n_tiles;
  t.ph = 0;
  if (UPS) t.ph = u & 3, u >>= 2;
  int m = u * CG + rank;
  if (CG == 2 && m >= p.m_tiles) {
    t.x0 = 0, t.y0 = 0, t.n = p.N;
    return t;
  }
  t.x0 = (m % p.tiles_x) * Geo<GEO>::kOutW;
  m /= p.tiles_x;
  t.y0 = (m % p.tiles_y) * Geo<GEO>::kRows;
  t.n = m / p.tiles_y;
  return t;
}

// ------------------------------------------------------------------ main kernel
// One elected lane of a converged warp (the compiler keeps descriptors / barrier addresses in
// uniform registers; a `lane == 0` branch instead makes it wrap every tcgen05/TMA instruction in a
// per-lane waterfall loop that costs ~600 issue cycles per K block).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// barrier among the 128 threads of one epilogue group (ids 1, 2; id 0 is __syncthreads)
__device__ __forceinline__ void epi_barrier(int group) {
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

constexpr int kStoreBytes = kBlockM * 128;  // one 64-channel chunk of a 128-pixel tile, 16 KiB

// Output tensor maps: [0] main store; [1..3] the other three 2x2 replicas of the fused upsample.
struct OutMaps {
  CUtensorMap m[4];
};

// direct (register) stores of the reflection-halo aliases of pixel (y, x); the pixel itself goes
// out through the TMA store of the staged tile
template <typename T16>
__device__ __forceinline__ void store_aliases(const ActView<T16>& out, int n, int y, int x, int co,
                                              const uint32_t (&pk)[32], int edge = 1) {
  const bool ya = (y == edge) || (y == out.H - 1 - edge), xa = (x == edge) || (x == out.W - 1 - edge);
  if (!(ya || xa)) return;
  for_each_halo_alias(y, x, out.H, out.W, [&](int yy, int xx) {
    if (yy == y && xx == x) return;
    uint4* dst = reinterpret_cast<uint4*>(out.px(n, yy, xx) + co);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }, edge);
}

template <typename T16, int BN, int EPI, int BRES, int CG, int GEO = 0>
__global__ void __launch_bounds__(kThreadsUmma, 1)
    conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a,
                     const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ OutMaps tmap_out, ConvParams<T16> p) {
  // UPS (EPI_UPS): `p.H x p.W` is the low-resolution input S (replicate halo); output phase (a, b)
  // holds pixels (2y + a, 2x + b) = sum over the 2x2 source window S[y + a - 1 + dy][x + b - 1 + dx]
  // with the 3x3 taps that fall on the same source pixel pre-summed (api.cu pack_layer).
  constexpr bool UPS = (EPI == EPI_UPS);
  constexpr int kTR = UPS ? 2 : 3, kTS = UPS ? 2 : 3;
  using G = Geo<GEO>;
  constexpr bool LIN = G::kLin;
  using Cfg = UmmaCfg<BN, BRES, CG, UPS, GEO>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte aligned bases (the dynamic shared window starts at the same
  // offset in both CTAs of a pair, so the carve-up below is identical in both)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic alias of smem_base
  const uint32_t store_base = smem_base + Cfg::kStoreOff;
  const uint32_t bar_base = smem_base + Cfg::kBarOff;
  float* s_bias = reinterpret_cast<float*>(smem_gen + Cfg::kBiasOff);
  auto a_smem = [&](int s) { return smem_base + Cfg::kAOff + s * Cfg::kASlabBytes; };
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
  const int unit0 = CG == 2 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int unit_step = CG == 2 ? (int)ncluster_x() : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (EPI != EPI_NCHW_F32) prefetch_tmap(&tmap_out.m[0]);
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
        const int row_ph = UPS ? (unit0 & 3) * p.CoutPad : 0;
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
    for (int unit = unit0; unit < p.total_tiles; unit += unit_step) {
      const TileCoord t = decode_tile<CG, UPS, GEO>(p, unit, (int)cta_rank);
      const int xs0 = t.x0 + ((UPS && !LIN) ? (t.ph & 1) : 0);
      const int b_row = (UPS ? t.ph * p.CoutPad : 0) + t.nt * BN + b_row0;
      for (int kc = 0; kc < kchunks; ++kc) {
        for (int s = 0; s < kTS; ++s) {
          if (!LIN || s == 0) {
            MBAR_WAIT_RELAXED(a_empty(as), aph ^ 1, 100 + as);
            if (p.ablate & 4) {
              if (elect_one()) {
                if (leader) mbar_arrive(a_full(as));
              }
            } else if (elect_one()) {
              if (leader) mbar_expect_tx(a_full(as), CG * Cfg::kASlabBytes);
              // interior pixel (y, x) is stored at (y+1, x+1): the slab for filter column s starts at
              // padded (y0, x0 + s) and spans the rows needed by r = 0..2 (UPS: source column
              // x + b - 1 + s, rows a + r of the slab).  LIN: one slab at (y0, x0) for every tap.
              tma_load_4d_cg<CG>(a_smem(as), &tmap_a, lead(a_full(as)), kc * kBlockK, xs0 + s, t.y0,
                                 t.n);
            }
            __syncwarp();
            if (++as == Cfg::kAStages) as = 0, aph ^= 1;
          }
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
      for (int unit = unit0; unit < p.total_tiles; unit += unit_step, ++it) {
        const int acs = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        if (CG == 2) mbar_wait_cluster(tmem_empty_bar(acs), aphase ^ 1, 200 + acs);
        else mbar_wait(tmem_empty_bar(acs), aphase ^ 1, 200 + acs);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acs * BN);
        // UPS: row phase a of this unit shifts the slab rows of the two taps to a + r
        const int row_shift = UPS ? (((unit / p.n_tiles) & 3) >> 1) : 0;
        const int col_shift = (UPS && LIN) ? ((unit / p.n_tiles) & 1) : 0;
        if (LIN && BRES == 1) {
          // linear slab + resident weights: the whole tile (all taps x 4 K steps) in ONE region
          mbar_wait(a_full(as), aph, 300 + as);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc0 =
                make_kmajor_sw128_desc(a_smem(as) + (uint32_t)(row_shift * G::kBoxW + col_shift) * 128u);
#pragma unroll
            for (int s = 0; s < kTS; ++s)
#pragma unroll
              for (int r = 0; r < kTR; ++r) {
                const uint64_t adesc = adesc0 + (uint64_t)((r * G::kBoxW + s) * 128 >> 4);
                const uint64_t bdesc = make_kmajor_sw128_desc(b_smem(r * kTS + s));
                if (!(p.ablate & 2)) {
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k)
                    umma_f16_cg<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (s | r | k) ? 1u : 0u);
                }
              }
            umma_commit_cg<CG>(a_empty(as));
            umma_commit_cg<CG>(tmem_full_bar(acs));
          }
          __syncwarp();
          if (++as == Cfg::kAStages) as = 0, aph ^= 1;
          continue;
        }
        for (int kc = 0; kc < kchunks; ++kc) {
#pragma unroll
          for (int s = 0; s < kTS; ++s) {
            // One elected-lane region per (chunk, filter column): all kTR filter rows x 4 K steps are
            // issued back to back.  (Electing per tap cost ~40 scalar/uniform instructions around
            // every 4 MMAs -- ~200 issue cycles against 128 cycles of math at N = 64 -- which made
            // the issuing warp, not the tensor pipe, the bound of the 64-channel layers.)
            if (!LIN || s == 0) {
              mbar_wait(a_full(as), aph, 300 + as);
            }
            const int bs0 = bs;
            if (!BRES) {
              mbar_wait(b_full(bs), bph, 350 + bs);
              if ((bs += kTR) == Cfg::kBStages) bs = 0, bph ^= 1;
            }
            tc_fence_after();
            if (elect_one()) {
              // tap (r, s): slab shifted by r rows (kBoxW px * 128 B, swizzle-phase neutral for the
              // 16-px box); LIN: and by s pixels inside the same slab
              const uint32_t a0 =
                  a_smem(as) + (uint32_t)(row_shift * G::kBoxW + (LIN ? s + col_shift : 0)) * 128u;
              const uint64_t adesc0 = make_kmajor_sw128_desc(a0);
#pragma unroll
              for (int r = 0; r < kTR; ++r) {
                const uint64_t adesc = adesc0 + (uint64_t)(r * (G::kBoxW * 128 >> 4));
                const uint64_t bdesc = make_kmajor_sw128_desc(BRES ? b_smem(kc * Cfg::kTaps + r * kTS + s) : b_smem(bs0 + r));
                if (!(p.ablate & 2)) {
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k) {
                    // +16 elements (32 bytes) along K inside the swizzle atom = +2 in the start field
                    umma_f16_cg<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc,
                                    (kc | s | r | k) ? 1u : 0u);
                  }
                }
              }
              if (!BRES) umma_commit_cg<CG>(b_empty(bs0));  // frees the weight tiles when these MMAs retire
              if (!LIN || s == kTS - 1) umma_commit_cg<CG>(a_empty(as));  // ... and the slab after its last tap
              if (kc == kchunks - 1 && s == kTS - 1) umma_commit_cg<CG>(tmem_full_bar(acs));
            }
            __syncwarp();
            if (!LIN || s == kTS - 1) {
              if (++as == Cfg::kAStages) as = 0, aph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: two groups of 4 warps, group g drains accumulator stage g
    // (tiles it = g, g+2, ...), so each group has two tile-times to finish one tile ==========
    const int grp = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;           // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;    // accumulator row = pixel inside the tile
    const int py = row / G::kBoxW, px = row % G::kBoxW;
    const bool col_ok = !LIN || px < G::kOutW;  // LIN: the last two columns wrap into the next row
    // warp 4 owns the bulk-store async groups: all its lanes execute the waits (a no-op for lanes
    // without groups), one elected lane -- always the same one -- issues and commits the stores
    const bool issuer_warp = (quad == 0);
    const uint32_t sbuf = store_base + grp * kStoreBytes;
    for (int it = grp;; it += 2) {
      const long long unit_ll = (long long)unit0 + (long long)it * unit_step;
      if (unit_ll >= p.total_tiles) break;
      const TileCoord t = decode_tile<CG, UPS, GEO>(p, (int)unit_ll, (int)cta_rank);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int y = t.y0 + py, x = t.x0 + px;
      const bool valid = col_ok && (y < p.H) && (x < p.W) && (CG == 1 || t.n < p.N);
      MBAR_WAIT_RELAXED(tmem_full_bar(as), aphase, 400 + as);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
      if (p.ablate & 1) {
        // measurement only: hand the accumulator back untouched
      } else if (EPI == EPI_NCHW_F32) {
        uint32_t r[16];
        tmem_ld16(taddr, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            if (c < p.Cout) {
              float v = __uint_as_float(r[c]) + s_bias[c];
              if (p.relu) v = fmaxf(v, 0.f);
              if (p.out_u8) p.out_u8[(((size_t)t.n * p.H + y) * p.W + x) * p.Cout + c] = quantize_u8(v);
              else p.out_nchw[(((size_t)t.n * p.Cout + c) * p.H + y) * p.W + x] = v;
            }
          }
        }
      } else {
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
            const float v0 = __uint_as_float(r[2 * j]) + s_bias[co + 2 * j];
            const float v1 = __uint_as_float(r[2 * j + 1]) + s_bias[co + 2 * j + 1];
            pk[j] = p.relu ? pack16x2_relu<T16>(v0, v1) : pack16x2<T16>(v0, v1);
            if (EPI == EPI_ACT_POOL) {
              // 2x2 window = lanes {l, l^1, l^16, l^17}, pooled on the packed pairs (rounding and
              // ReLU are monotonic, so max commutes with them: half the shuffles of fp32 pooling);
              // out-of-image pixels contribute 0, the identity for post-ReLU values
              uint32_t w = valid ? pk[j] : 0u;
              w = max16x2<T16>(w, __shfl_xor_sync(0xffffffffu, w, 1));
              w = max16x2<T16>(w, __shfl_xor_sync(0xffffffffu, w, 16));
              pk[j] = w;
            }
          }
          // the staging buffer about to be rewritten must have been read out by its TMA store (waited
          // for only now, so that the bias / ReLU / pack work above overlaps that read-out)
          if (issuer_warp) bulk_wait_read<0>();
          epi_barrier(grp);
          // stage the row (128 bytes = 8 chunks) with the 128-byte swizzle the TMA store expects
          int srow = LIN ? py * G::kOutW + px : row;
          bool writer = col_ok;
          if (EPI == EPI_ACT_POOL) {
            static_assert(EPI != EPI_ACT_POOL || G::kBoxW == 16, "fused pooling needs 2 tile rows per warp");
            writer = col_ok && !(lane & 1) && lane < 16;  // anchor of a 2x2 window
            srow = (py >> 1) * (G::kOutW / 2) + (px >> 1);
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
            // statistics of the STORED (rounded) values, read back from the staged tile: thread =
            // (channel pair = lane, quarter of the tile = two tile rows = warp); exact two-pass over
            // the quarter's valid pixels, conflict-free (a warp reads one 128-byte staged row at a time)
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
      if (EPI == EPI_NCHW_F32 || (p.ablate & 1)) {
        // all TMEM reads of this accumulator stage are complete (wait::ld above)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(lead(tmem_empty_bar(as)));
          else mbar_arrive(tmem_empty_bar(as));
        }
      }
    }
    if (issuer_warp) bulk_wait_all();
  }

  __syncwarp();
  tc_fence_before();
  // pair: the peer's shared memory / TMEM are operands of the leader's MMAs until the very end
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc_cg<CG, Cfg::kTmemCols>(tmem_base);
}

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

// NG = number of 4-warp epilogue groups (tile i is drained by group i % NG from accumulator stage
// i % 2): the s-merged epilogue (192 accumulator columns, shuffles, pooling) is latency-bound with two
// epilogue warps per scheduler, a third group adds issue capacity.
template <bool BRES, int CG, int NG = 2>
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
  static constexpr int kThreads = 128 + NG * 128;                  // 4 control warps + NG epilogue groups
  // "accumulator ready" barriers: one per residue of the tile counter mod lcm(2 stages, NG groups), so
  // that each barrier is waited on by ONE group and that group sees every one of its phases (a parity
  // wait cannot tell a phase from the one two earlier)
  static constexpr int kFullBars = NG == 2 ? 2 : 2 * NG;
  static constexpr int kBiasOff = kStoreOff + NG * kSmStoreBytes;
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

template <typename T16, int EPI, bool BRES, int CG, int NG = 2>
__global__ void __launch_bounds__(128 + NG * 128, 1)
    conv_smerge_kernel(const __grid_constant__ CUtensorMap tmap_a,
                       const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ OutMaps tmap_out, ConvParams<T16> p) {
  using Cfg = SmergeCfg<BRES, CG, NG>;
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
              if (!(p.ablate & 2)) {
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
    for (int it = grp;; it += NG) {
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
      if (p.ablate & 1) {  // measurement only: hand the accumulator back untouched
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
  }

  __syncwarp();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc_cg<CG, Cfg::kTmemCols>(tmem_base);
}

// =====================================================================================
// conv1_1 (+ folded 1x1 colour conv, net.py:39-42) on the tensor cores.
// K = 27 is too thin for TMA-fed tiles, so the 128 threads of a CTA build the im2col rows
// themselves: CTA tile = 128 consecutive pixels of one image row; thread p gathers the 27 taps of
// pixel p from a staged fp32 window of the NCHW image (reflection applied while staging), converts
// to T16 and writes one 64-byte K-major row (K padded to 32) into shared memory with the 128-byte
// swizzle applied by hand (16-byte chunk j of row r lives at chunk j ^ (r & 7)).  One thread then
// issues two tcgen05.mma (M=128, N=64, K=16), the accumulator comes back through tcgen05.ld and is
// stored as NHWC (the tile is one contiguous 16 KiB span of the activation).
// =====================================================================================
constexpr int kFirstPx = 128;

template <typename T16>
struct FirstParams {
  const float* img;  // [N,3,H,W]
  int N, H, W;
  const T16* wk;     // [64][32] K-major (k = (r*3+s)*3 + ci, 27..31 zero)
  const float* bias; // [64]
  ActView<T16> out;
  int total_tiles, tiles_x;
};

constexpr int kFirstWin = 9 * (kFirstPx + 2);         // 3 channels x 3 rows x 130 columns
constexpr int kFirstWinBytes = (kFirstWin * 4 + 127) / 128 * 128;
constexpr int kFirstSmem = 1024 /*align*/ + kFirstPx * 128 * 2 + 64 * 128 + 2 * kFirstWinBytes + 256 + 64;

template <typename T16>
__global__ void __launch_bounds__(kFirstPx)
    conv_first_umma_kernel(const __grid_constant__ CUtensorMap tmap_out, FirstParams<T16> p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base;                              // im2col rows, 128 x 128 B (swizzled)
  const uint32_t sOut = base + kFirstPx * 128;           // staged output tile for the TMA store
  uint8_t* sB_gen = gen + 2 * kFirstPx * 128;            // weights, 64 x 128 B (swizzled)
  const uint32_t sB = base + 2 * kFirstPx * 128;
  const uint32_t win_off = 2 * kFirstPx * 128 + 64 * 128;
  float* sbias = reinterpret_cast<float*>(gen + win_off + 2 * kFirstWinBytes);
  const uint32_t bar = base + win_off + 2 * kFirstWinBytes + 256;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(gen + win_off + 2 * kFirstWinBytes + 256 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;

  // one-time: weights -> swizzled K-major B tile, barrier, TMEM
  for (int i = tid; i < 64 * 4; i += kFirstPx) {
    const int o = i >> 2, j = i & 3;
    const uint4 v = reinterpret_cast<const uint4*>(p.wk)[o * 4 + j];
    *reinterpret_cast<uint4*>(sB_gen + o * 128 + ((j ^ (o & 7)) << 4)) = v;
  }
  if (tid < 64) sbias[tid] = p.bias[tid];
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    prefetch_tmap(&tmap_out);
  }
  if (warp == 0) tmem_alloc<64>(base + win_off + 2 * kFirstWinBytes + 256 + 16);

  // Input window of one tile: 3 ch x 3 rows x 130 cols of the NCHW fp32 image, reflection resolved
  // per element, fetched with 4-byte cp.async one tile AHEAD of its use (the loads are the only
  // DRAM-latency-bound part of this kernel).
  auto stage_window = [&](int tile, int buf) {
    int b = tile;
    const int x0 = (b % p.tiles_x) * kFirstPx;
    b /= p.tiles_x;
    const int y = b % p.H;
    const int n = b / p.H;
    const uint32_t dst0 = base + win_off + buf * kFirstWinBytes;
    for (int i = tid; i < kFirstWin; i += kFirstPx) {
      const int col = i % (kFirstPx + 2);
      const int rc = i / (kFirstPx + 2);  // ci*3 + row
      const int row = rc % 3, ci = rc / 3;
      int yy = y + row - 1;
      yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);
      int xx = x0 + col - 1;
      if (xx <= p.W) {
        xx = xx < 0 ? -xx : (xx >= p.W ? 2 * p.W - 2 - xx : xx);
        const float* src = p.img + (((size_t)n * 3 + ci) * p.H + yy) * p.W + xx;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + 4 * i), "l"(src)
                     : "memory");
      } else {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst0 + 4 * i), "r"(0u) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if ((int)blockIdx.x < p.total_tiles) stage_window(blockIdx.x, 0);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const uint64_t adesc = make_kmajor_sw128_desc(sA);
  const uint64_t bdesc = make_kmajor_sw128_desc(sB);
  constexpr uint32_t idesc = make_idesc<T16, 64>();
  uint32_t phase = 0;
  int buf = 0;

  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, buf ^= 1) {
    int b = tile;
    const int x0 = (b % p.tiles_x) * kFirstPx;
    b /= p.tiles_x;
    const int y = b % p.H;
    const int n = b / p.H;
    // (1) prefetch the next tile's window, then wait for this tile's
    const int next = tile + gridDim.x;
    if (next < p.total_tiles) {
      stage_window(next, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* win = reinterpret_cast<const float*>(gen + win_off + buf * kFirstWinBytes);
    // (2) im2col row of pixel tid -> swizzled K-major A tile
    {
      uint32_t pk[16];
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 2 * k2 + e;
          if (k < 27) {
            const int tap = k / 3, ci = k - 3 * tap;
            const int r = tap / 3, s = tap - 3 * r;
            v[e] = win[(ci * 3 + r) * (kFirstPx + 2) + tid + s];
          } else {
            v[e] = 0.f;
          }
        }
        pk[k2] = pack16x2<T16>(v[0], v[1]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t dst = sA + tid * 128 + ((j ^ (tid & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]),
                     "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                     : "memory");
      }
    }
    fence_async_smem();  // generic smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    // (3) two K=16 steps, issued by one elected lane of the converged warp 0
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        umma_bf16(tmem_base, adesc, bdesc, idesc, 0u);
        umma_bf16(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);
        umma_commit(bar);
      }
      __syncwarp();
    }
    // (4) accumulator ready
    mbar_wait(bar, phase, 900);
    phase ^= 1;
    tc_fence_after();
    // (5) epilogue: row tid of the accumulator = pixel x0 + tid
    const int x = x0 + tid;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t r0[32], r1[32];
    tmem_ld32(taddr, r0);
    tmem_ld32(taddr + 32, r1);
    tmem_ld_wait();
    uint32_t pk[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      pk[j] = pack16x2_relu<T16>(__uint_as_float(r0[2 * j]) + sbias[2 * j],
                                 __uint_as_float(r0[2 * j + 1]) + sbias[2 * j + 1]);
      pk[16 + j] = pack16x2_relu<T16>(__uint_as_float(r1[2 * j]) + sbias[32 + 2 * j],
                                      __uint_as_float(r1[2 * j + 1]) + sbias[33 + 2 * j]);
    }
    // TMEM reads are complete (wait::ld); every thread passes two more block barriers before warp 0
    // overwrites the accumulator with the next tile
    tc_fence_before();
    // each warp stages and stores its own 32-pixel quarter of the row segment, so only its own
    // previous TMA store has to have drained (no block-wide barrier on the store path)
    bulk_wait_read<0>();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t dst = sOut + tid * 128 + ((j ^ (tid & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]),
                   "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                   : "memory");
    }
    if (x < p.W) store_aliases(p.out, n, y, x, 0, pk);
    fence_async_smem();
    __syncwarp();
    if (elect_one()) {
      tma_store_4d(&tmap_out, sOut + warp * (32 * 128), 0, x0 + warp * 32, y, n);  // clipped at W
      bulk_commit();
    }
    __syncwarp();
  }
  bulk_wait_all();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem_base);
}

// =====================================================================================
// conv1_1, warp-specialised persistent variant (used when W % 4 == 0, i.e. the image rows are
// 16-byte aligned and TMA can fetch the input window).  Same math as conv_first_umma_kernel; the
// per-tile chain  window -> im2col rows -> MMA -> epilogue -> store  is cut into four roles that
// run on different tiles at the same time, so the kernel is bound by its HBM writes (128 B per
// pixel) instead of by the latency of the chain:
//   warp 0      TMA producer: {136 col, 3 row, 3 ch} fp32 window of the NCHW image per tile, ring
//               of kF2WinStages (out-of-image rows / columns arrive as zeros and are never read:
//               reflection is an index remap in the builders)
//   warps 1-4   builders: thread p gathers the 27 taps of pixel p, converts to T16 and writes the
//               swizzled K-major row p of the A tile (ring of 2)
//   warp 5      MMA issuer: 2 x tcgen05.mma (M=128, N=64, K=16) per tile into one of 2 TMEM stages
//   warps 6-9   epilogue: tcgen05.ld -> bias + ReLU -> T16 -> per-warp staging -> TMA store
// =====================================================================================
constexpr int kF2Threads = 320;
constexpr int kF2WinCols = 136;  // columns x0-4 .. x0+131: a non-swizzled TMA box must start 16-byte aligned
constexpr int kF2WinX0 = 4;     // window column of pixel x0
constexpr int kF2WinElems = 9 * kF2WinCols;
constexpr int kF2WinTx = kF2WinElems * 4;                      // 4896 bytes per TMA box
constexpr int kF2WinBytes = (kF2WinTx + 127) / 128 * 128;      // 4992
constexpr int kF2WinStages = 4;
constexpr int kF2ABytes = kFirstPx * 128;                      // 16 KiB
constexpr int kF2OffA = 0;                                     // 2 A tiles
constexpr int kF2OffOut = 2 * kF2ABytes;                       // 2 staging tiles
constexpr int kF2OffB = 4 * kF2ABytes;                         // weights 64 x 128 B
constexpr int kF2OffWin = kF2OffB + 64 * 128;
constexpr int kF2OffBias = kF2OffWin + kF2WinStages * kF2WinBytes;
constexpr int kF2OffBar = kF2OffBias + 256;
constexpr int kF2NumBars = 2 * kF2WinStages + 8;
constexpr int kF2Smem = 1024 + kF2OffBar + 8 * kF2NumBars + 16;

template <typename T16>
__global__ void __launch_bounds__(kF2Threads, 2)
    conv_first_umma_ws_kernel(const __grid_constant__ CUtensorMap tmap_img,
                              const __grid_constant__ CUtensorMap tmap_out, FirstParams<T16> p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kF2OffBar;
  auto win_full = [&](int s) { return bar0 + 8u * s; };
  auto win_empty = [&](int s) { return bar0 + 8u * (kF2WinStages + s); };
  auto a_full = [&](int s) { return bar0 + 8u * (2 * kF2WinStages + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 * kF2WinStages + 2 + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * kF2WinStages + 4 + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * kF2WinStages + 6 + s); };
  const uint32_t tmem_slot = bar0 + 8u * kF2NumBars;
  float* sbias = reinterpret_cast<float*>(gen + kF2OffBias);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // one-time: weights -> swizzled K-major B tile, bias, barriers, TMEM (2 stages x 64 columns)
  for (int i = tid; i < 64 * 4; i += kF2Threads) {
    const int o = i >> 2, j = i & 3;
    const uint4 v = reinterpret_cast<const uint4*>(p.wk)[o * 4 + j];
    *reinterpret_cast<uint4*>(gen + kF2OffB + o * 128 + ((j ^ (o & 7)) << 4)) = v;
  }
  if (tid < 64) sbias[tid] = p.bias[tid];
  if (tid == 0) {
    for (int s = 0; s < kF2WinStages; ++s) {
      mbar_init(win_full(s), 1);
      mbar_init(win_empty(s), 128);  // every builder thread arrives for itself
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), 128);
      mbar_init(a_empty(s), 1);
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), 4);
    }
    fence_barrier_init();
    prefetch_tmap(&tmap_img);
    prefetch_tmap(&tmap_out);
  }
  if (warp == 5) tmem_alloc<128>(tmem_slot);
  fence_async_smem();  // the weight tile is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + kF2OffBar + 8 * kF2NumBars);
  pdl_launch_dependents();
  pdl_wait();

  auto tile_coord = [&](int tile, int& n, int& y, int& x0) {
    x0 = (tile % p.tiles_x) * kFirstPx;
    const int b = tile / p.tiles_x;
    y = b % p.H;
    n = b / p.H;
  };

  if (warp == 0) {
    // ===================== TMA producer
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int ws = it % kF2WinStages;
      const uint32_t ph = (it / kF2WinStages) & 1;
      int n, y, x0;
      tile_coord(tile, n, y, x0);
      mbar_wait(win_empty(ws), ph ^ 1, 910);
      if (elect_one()) {
        mbar_expect_tx(win_full(ws), kF2WinTx);
        tma_load_4d(base + kF2OffWin + ws * kF2WinBytes, &tmap_img, win_full(ws), x0 - kF2WinX0, y - 1, 0, n);
      }
      __syncwarp();
    }
  } else if (warp <= 4) {
    // ===================== builders: im2col row of pixel px of the tile
    const int px = tid - 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int ws = it % kF2WinStages, as = it & 1;
      int n, y, x0;
      tile_coord(tile, n, y, x0);
      // reflection = index remap inside the window (rows y-1..y+1 at 0..2, columns from x0-4)
      int ridx[3] = {0, 1, 2};
      if (y == 0) ridx[0] = 2;
      if (y == p.H - 1) ridx[2] = 0;
      int cidx[3];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int xx = x0 + px + s - 1;
        xx = xx < 0 ? -xx : (xx >= p.W ? 2 * p.W - 2 - xx : xx);
        int c = xx - (x0 - kF2WinX0);
        cidx[s] = c < 0 ? 0 : (c > kF2WinCols - 1 ? kF2WinCols - 1 : c);  // only for pixels past W
      }
      mbar_wait(win_full(ws), (it / kF2WinStages) & 1, 920);
      const float* win = reinterpret_cast<const float*>(gen + kF2OffWin + ws * kF2WinBytes);
      float v[28];
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const int tap = k / 3, ci = k - 3 * tap;
        const int r = tap / 3, s = tap - 3 * r;
        v[k] = win[(ci * 3 + ridx[r]) * kF2WinCols + cidx[s]];
      }
      v[27] = 0.f;
      // The window is rewritten by TMA (async proxy): this thread's generic-proxy reads must be
      // ordered before that write, which takes a proxy fence before the release (without it a
      // 32-pixel quarter of a tile came out wrong about once per 10^5 tiles).
      fence_async_smem();
      mbar_arrive(win_empty(ws));
      uint32_t pk[16];
#pragma unroll
      for (int k2 = 0; k2 < 14; ++k2) pk[k2] = pack16x2<T16>(v[2 * k2], v[2 * k2 + 1]);
      pk[14] = 0u, pk[15] = 0u;
      MBAR_WAIT_RELAXED(a_empty(as), ((it >> 1) & 1) ^ 1, 930);
      const uint32_t sA = base + kF2OffA + as * kF2ABytes;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t dst = sA + px * 128 + ((j ^ (px & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]),
                     "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                     : "memory");
      }
      fence_async_smem();  // this thread's generic writes -> visible to the tensor core
      mbar_arrive(a_full(as));
    }
  } else if (warp == 5) {
    // ===================== MMA issuer
    constexpr uint32_t idesc = make_idesc<T16, 64>();
    const uint64_t bdesc = make_kmajor_sw128_desc(base + kF2OffB);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(t_empty(as), ph ^ 1, 940);
      mbar_wait(a_full(as), ph, 941);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = make_kmajor_sw128_desc(base + kF2OffA + as * kF2ABytes);
        const uint32_t d = tmem_base + (uint32_t)(as * 64);
        umma_bf16(d, adesc, bdesc, idesc, 0u);
        umma_bf16(d, adesc + 2, bdesc + 2, idesc, 1u);
        umma_commit(a_empty(as));
        umma_commit(t_full(as));
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue: warp q owns TMEM lanes 32q..32q+31 = pixels 32q.. of the tile
    const int q = warp & 3;  // warps 6,7,8,9 -> lane quadrants 2,3,0,1
    const int px = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      int n, y, x0;
      tile_coord(tile, n, y, x0);
      MBAR_WAIT_RELAXED(t_full(as), (it >> 1) & 1, 950);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 64);
      uint32_t r0[32], r1[32];
      tmem_ld32(taddr, r0);
      tmem_ld32(taddr + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty(as));
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        pk[j] = pack16x2_relu<T16>(__uint_as_float(r0[2 * j]) + sbias[2 * j],
                                   __uint_as_float(r0[2 * j + 1]) + sbias[2 * j + 1]);
        pk[16 + j] = pack16x2_relu<T16>(__uint_as_float(r1[2 * j]) + sbias[32 + 2 * j],
                                        __uint_as_float(r1[2 * j + 1]) + sbias[33 + 2 * j]);
      }
      // per-warp staging (two buffers): the store issued two tiles ago must have read its buffer
      bulk_wait_read<1>();
      __syncwarp();
      const uint32_t sOut = base + kF2OffOut + as * kF2ABytes;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t dst = sOut + px * 128 + ((j ^ (px & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]),
                     "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                     : "memory");
      }
      const int x = x0 + px;
      if (x < p.W) store_aliases(p.out, n, y, x, 0, pk);
      fence_async_smem();
      __syncwarp();
      if (elect_one()) {
        tma_store_4d(&tmap_out, sOut + q * (32 * 128), 0, x0 + q * 32, y, n);  // clipped at W
        bulk_commit();
      }
      __syncwarp();
    }
    bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<128>(tmem_base);
}

// =====================================================================================
// Last decoder conv (net.py:34-35, 64 -> 3 channels, no ReLU, fp32 NCHW result).
// With only 3 output channels the tap-by-tap implicit GEMM reads its A operand nine times from
// shared memory for N = 16 columns each and is bound by shared-memory bandwidth (measured 0.56 ms
// for batch 32 @512^2, HBM floor 0.17 ms).  Here the taps move into the N dimension instead:
//   P[j, (tap, co)] = sum_c X[j, c] * W[tap][c][co]        one 1x1 GEMM, N = 27 (padded to 32), K = 64
//   out[y, x, co]   = bias[co] + sum_tap P[(y + r, x + s), (tap, co)]     9-point gather
// X is the {64 ch, 18 px, 10 rows} halo slab of an 8x16 output tile (180 pixels, ONE TMA load,
// every input pixel read from shared memory once per K step instead of nine times); two M = 128
// MMA chains cover slab pixels 0..127 and 128..255 (pixels >= 180 are whatever follows in shared
// memory; their accumulator rows are never read).  The epilogue moves P through shared memory
// (fp32, pitch 29 words: conflict-free both ways) and every thread gathers one output pixel.
// =====================================================================================
constexpr int kLSlabW = kTileW + 2, kLSlabH = kTileH + 2;
constexpr int kLSlabPx = kLSlabW * kLSlabH;            // 180
constexpr int kLSlabBytes = kLSlabPx * 128;            // 23040
constexpr int kLStageStride = (kLSlabBytes + 1023) / 1024 * 1024;  // 23552
constexpr int kLStages = 4;
constexpr int kLPitch = 29;                            // words per P row
constexpr int kLPBytes = (kLSlabPx * kLPitch * 4 + 127) / 128 * 128;
constexpr int kLOffB = kLStages * kLStageStride;       // B' 32 x 128 B (the second MMA chain of the
                                                       // last stage reads past its slab into here)
constexpr int kLOffP = kLOffB + 32 * 128 + 8192;       // + slack so slab + 32 KiB stays in bounds
constexpr int kLOffBar = kLOffP + 2 * kLPBytes;
constexpr int kLNumBars = 2 * kLStages + 4;
constexpr int kLSmem = 1024 + kLOffBar + 8 * kLNumBars + 16;
static_assert((kLStages - 1) * kLStageStride + 256 * 128 <= kLOffP, "second MMA chain must stay inside the buffer");

template <typename T16>
__global__ void __launch_bounds__(kThreadsUmma, 1)
    conv_last_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const T16* __restrict__ wk,
                          ConvParams<T16> p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kLOffBar;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (kLStages + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * kLStages + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * kLStages + 2 + s); };
  const uint32_t tmem_slot = bar0 + 8u * kLNumBars;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // B'[n = tap*3 + co][k = c] from the packed weights wk[co][tap*64 + c]; rows 27..31 are zero
  for (int i = threadIdx.x; i < 32 * 8; i += kThreadsUmma) {
    const int n = i >> 3, j = i & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const int tap = n / 3, co = n - 3 * tap;
    if (n < 27 && co < p.Cout) v = *reinterpret_cast<const uint4*>(wk + (size_t)co * (9 * kBlockK) + tap * kBlockK + j * 8);
    *reinterpret_cast<uint4*>(gen + kLOffB + n * 128 + ((j ^ (n & 7)) << 4)) = v;
  }
  if (warp == 0 && lane == 0) prefetch_tmap(&tmap_a);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kLStages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<128>(tmem_slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + kLOffBar + 8 * kLNumBars);

  if (warp == 0) {
    // ===================== TMA producer: one halo slab per tile
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const TileCoord t = decode_tile<1>(p, tile, 0);
      const int s = it % kLStages;
      MBAR_WAIT_RELAXED(a_empty(s), ((it / kLStages) & 1) ^ 1, 700 + s);
      if (elect_one()) {
        mbar_expect_tx(a_full(s), kLSlabBytes);
        tma_load_4d(base + s * kLStageStride, &tmap_a, a_full(s), 0, t.x0, t.y0, t.n);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 2 chains (slab pixels 0..127, 128..255) x 4 K steps, N = 32
    constexpr uint32_t idesc = make_idesc<T16, 32>();
    const uint64_t bdesc = make_kmajor_sw128_desc(base + kLOffB);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int s = it % kLStages, acs = it & 1;
      mbar_wait(t_empty(acs), ((it >> 1) & 1) ^ 1, 710 + acs);
      mbar_wait(a_full(s), (it / kLStages) & 1, 720 + s);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t adesc = make_kmajor_sw128_desc(base + s * kLStageStride + h * (kBlockM * 128));
          const uint32_t d = tmem_base + (uint32_t)(acs * 64 + h * 32);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
        }
        umma_commit(a_empty(s));
        umma_commit(t_full(acs));
      }
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: group g takes tiles g, g+2, ...
    const int grp = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // slab pixel (first chain) / output pixel of the tile
    const int py = row / kTileW, px = row % kTileW;
    float* P = reinterpret_cast<float*>(gen + kLOffP + grp * kLPBytes);
    float bias[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bias[c] = c < p.Cout ? p.bias[c] : 0.f;
    for (int it = grp;; it += 2) {
      const long long tile_ll = (long long)blockIdx.x + (long long)it * gridDim.x;
      if (tile_ll >= p.total_tiles) break;
      const TileCoord t = decode_tile<1>(p, (int)tile_ll, 0);
      const int acs = it & 1;
      MBAR_WAIT_RELAXED(t_full(acs), (it >> 1) & 1, 730 + acs);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acs * 64);
      uint32_t r0[32], r1[32];
      tmem_ld32(taddr, r0);
      tmem_ld32(taddr + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty(acs));
#pragma unroll
      for (int c = 0; c < 27; ++c) P[row * kLPitch + c] = __uint_as_float(r0[c]);
      if (row + kBlockM < kLSlabPx) {
#pragma unroll
        for (int c = 0; c < 27; ++c) P[(row + kBlockM) * kLPitch + c] = __uint_as_float(r1[c]);
      }
      epi_barrier(grp);
      float acc[3] = {bias[0], bias[1], bias[2]};
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int r = tap / 3, s = tap - 3 * r;
        const float* src = P + ((py + r) * kLSlabW + px + s) * kLPitch + tap * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += src[c];
      }
      const int y = t.y0 + py, x = t.x0 + px;
      if (y < p.H && x < p.W) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (c < p.Cout) {
            float v = acc[c];
            if (p.relu) v = fmaxf(v, 0.f);
            if (p.out_u8) p.out_u8[(((size_t)t.n * p.H + y) * p.W + x) * p.Cout + c] = quantize_u8(v);
            else p.out_nchw[(((size_t)t.n * p.Cout + c) * p.H + y) * p.W + x] = v;
          }
        }
      }
      epi_barrier(grp);  // P is rewritten by this group's next tile
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<128>(tmem_base);
}

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
constexpr int kU4AStages = 2;
constexpr int kU4OffB = kU4AStages * kU4SlabBytes;           // 16 tiles x 8 KiB
constexpr int kU4OffStore = kU4OffB + 16 * 8192;
constexpr int kU4StoreBytes = 16384;                          // 120 rows x 128 B, rounded
constexpr int kU4OffBias = kU4OffStore + 2 * kU4StoreBytes;
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
__global__ void __launch_bounds__(kThreadsUmma, 1)
    conv_ups4_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ OutMaps tmap_out, ConvParams<T16> p) {
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
  auto tile_of = [&](int tile, int& n, int& y0, int& x0) {
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
      mbar_init(t_empty(s), 4);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + kU4OffBar + 8 * kU4NumBars);
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer: resident phase weights once, then one slab per tile
    if (elect_one()) {
      mbar_expect_tx(bres_bar, 16 * 8192);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        tma_load_2d(base + kU4OffB + i * 8192, &tmap_b, bres_bar, kU4TileTap[i] * kBlockK, kU4TilePh[i] * 64);
    }
    __syncwarp();
    pdl_wait();
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int n, y0, x0;
      tile_of(tile, n, y0, x0);
      MBAR_WAIT_RELAXED(a_empty(s), ph ^ 1, 900 + s);
      if (elect_one()) {
        mbar_expect_tx(a_full(s), kU4SlabBytes);
        // slab position (jy, jx) = padded pixel (y0 + jy, x0 + jx) = source (y0 - 1 + jy, x0 - 1 + jx)
        tma_load_4d(base + s * kU4SlabBytes, &tmap_a, a_full(s), 0, x0, y0, n);
      }
      __syncwarp();
      if (++s == kU4AStages) s = 0, ph ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 10 operand views x 4 K steps per tile
    mbar_wait(bres_bar, 0, 905);
    tc_fence_after();
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acs = it & 1;
      mbar_wait(t_empty(acs), ((it >> 1) & 1) ^ 1, 910 + acs);
      mbar_wait(a_full(s), ph, 920 + s);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc0 = make_kmajor_sw128_desc(base + s * kU4SlabBytes);
        const uint64_t bdesc0 = make_kmajor_sw128_desc(base + kU4OffB);
        const uint32_t d0 = tmem_base + (uint32_t)(acs * 256);
#pragma unroll
        for (int o = 0; o < 10; ++o) {
          const uint64_t adesc = adesc0 + (uint64_t)((kU4Ops[o].R * kU4BoxW + kU4Ops[o].S) * 128 >> 4);
          const uint64_t bdesc = bdesc0 + (uint64_t)(kU4Ops[o].first * (8192 >> 4));
          const uint32_t d = d0 + (uint32_t)(kU4Ops[o].slot * 64);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint32_t acc = (o | k) ? 1u : 0u;
            if (kU4Ops[o].ntiles == 4) umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, make_idesc<T16, 256>(), acc);
            else if (kU4Ops[o].ntiles == 2) umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, make_idesc<T16, 128>(), acc);
            else umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, make_idesc<T16, 64>(), acc);
          }
        }
        umma_commit(a_empty(s));
        umma_commit(t_full(acs));
      }
      __syncwarp();
      if (++s == kU4AStages) s = 0, ph ^= 1;
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: group g drains accumulator stage g; warp <-> tile row, lane <-> column
    const int grp = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;
    const bool issuer_warp = (quad == 0);
    const uint32_t sbuf = base + kU4OffStore + grp * kU4StoreBytes;
    const int srow = quad * kU4OutW + lane;
    for (int it = grp;; it += 2) {
      const long long tile_ll = (long long)blockIdx.x + (long long)it * gridDim.x;
      if (tile_ll >= p.total_tiles) break;
      int n, y0, x0;
      tile_of((int)tile_ll, n, y0, x0);
      const int acs = it & 1;
      const int y = y0 + quad, x = x0 + lane;
      const bool col_ok = lane < kU4OutW;
      const bool valid = col_ok && y < p.H && x < p.W;
      MBAR_WAIT_RELAXED(t_full(acs), (it >> 1) & 1, 930 + acs);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acs * 256);
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        const int phs = kU4SlotPh[ch], a = phs >> 1, b = phs & 1;
        uint32_t r[64];
        {
          uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
          uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
          tmem_ld32(taddr + ch * 64, r0);
          tmem_ld32(taddr + ch * 64 + 32, r1);
        }
        tmem_ld_wait();
        if (ch == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty(acs));
        }
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v0 = __uint_as_float(r[2 * j]) + s_bias[2 * j];
          const float v1 = __uint_as_float(r[2 * j + 1]) + s_bias[2 * j + 1];
          pk[j] = p.relu ? pack16x2_relu<T16>(v0, v1) : pack16x2<T16>(v0, v1);
        }
        if (issuer_warp) bulk_wait_read<0>();  // the staging buffer has been read out by its TMA store
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
        if (issuer_warp && elect_one()) {
          tma_store_4d(&tmap_out.m[phs], sbuf, 0, x0, y0, n);
          bulk_commit();
        }
      }
    }
    if (issuer_warp) bulk_wait_all();
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// =====================================================================================
// Last decoder conv, second form: filter ROWS by operand shifts, filter COLUMNS in N.
// The gather form above is bound by its epilogue (54 shared-memory stores + 27 loads per pixel).
// Here the tile is 4 rows x 32 columns of a linear slab {64 ch, 32 px, 6 rows} (one TMA load):
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

template <typename T16>
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
  auto tile_of = [&](int tile, int& n, int& y0, int& x0) {
    x0 = (tile % p.tiles_x) * kROutW;
    tile /= p.tiles_x;
    y0 = (tile % p.tiles_y) * kRRows;
    n = tile / p.tiles_y;
  };

  // B_r[n = s*4 + co][k = c] = W[co][c][r][s] from the packed weights wk[co][(r*3+s)*64 + c]; K-major
  // rows of 128 B with the 128-byte swizzle; unused rows are zero
  for (int i = threadIdx.x; i < 3 * 16 * 8; i += kThreadsUmma) {
    const int r = i / 128, n = (i >> 3) & 15, j = i & 7;
    const int sc = n >> 2, co = n & 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (sc < 3 && co < p.Cout && co < 3)
      v = *reinterpret_cast<const uint4*>(wk + (size_t)co * (9 * kBlockK) + (r * 3 + sc) * kBlockK + j * 8);
    *reinterpret_cast<uint4*>(gen + kROffB + r * 2048 + n * 128 + ((j ^ (n & 7)) << 4)) = v;
  }
  if (warp == 0 && lane == 0) prefetch_tmap(&tmap_a);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRStages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<32>(tmem_slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + kROffBar + 8 * kRNumBars);
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer: one slab per tile
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int n, y0, x0;
      tile_of(tile, n, y0, x0);
      MBAR_WAIT_RELAXED(a_empty(s), ph ^ 1, 800 + s);
      if (elect_one()) {
        mbar_expect_tx(a_full(s), kRSlabBytes);
        tma_load_4d(base + s * kRSlabBytes, &tmap_a, a_full(s), 0, x0, y0, n);
      }
      __syncwarp();
      if (++s == kRStages) s = 0, ph ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 3 filter rows x 4 K steps, M = 128, N = 16
    constexpr uint32_t idesc = make_idesc<T16, 16>();
    const uint64_t bdesc0 = make_kmajor_sw128_desc(base + kROffB);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acs = it & 1;
      mbar_wait(t_empty(acs), ((it >> 1) & 1) ^ 1, 810 + acs);
      mbar_wait(a_full(s), ph, 820 + s);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc0 = make_kmajor_sw128_desc(base + s * kRSlabBytes);
        const uint32_t d = tmem_base + (uint32_t)(acs * 16);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(d, adesc0 + (uint64_t)(r * (kRBoxW * 128 >> 4) + 2 * k), bdesc0 + (uint64_t)(r * (2048 >> 4) + 2 * k),
                      idesc, (r | k) ? 1u : 0u);
        umma_commit(a_empty(s));
        umma_commit(t_full(acs));
      }
      __syncwarp();
      if (++s == kRStages) s = 0, ph ^= 1;
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: group g takes tiles g, g+2, ...; warp <-> tile row, lane <-> column
    const int grp = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;
    float bias[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bias[c] = c < p.Cout ? p.bias[c] : 0.f;
    for (int it = grp;; it += 2) {
      const long long tile_ll = (long long)blockIdx.x + (long long)it * gridDim.x;
      if (tile_ll >= p.total_tiles) break;
      int n, y0, x0;
      tile_of((int)tile_ll, n, y0, x0);
      const int acs = it & 1;
      MBAR_WAIT_RELAXED(t_full(acs), (it >> 1) & 1, 830 + acs);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acs * 16), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty(acs));
      const int y = y0 + quad, x = x0 + lane;
      const bool ok = lane < kROutW && y < p.H && x < p.W;
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
  __syncthreads();
  if (warp == 2) tmem_dealloc<32>(tmem_base);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

bool pdl_on() {
  // measured (batch 32 @512^2, 20 steps): 5.917 vs 5.941 ms per step -- the prologues are already
  // cheap next to the tails, so it stays off unless CCST_PDL=1
  static const bool on = [] { const char* e = getenv("CCST_PDL"); return e && e[0] == '1'; }();
  return on;
}

// launch with the programmatic-dependent-launch attribute (and the cluster dimension for CTA pairs)
template <typename... KArgs, typename... Args>
cudaError_t launch_conv(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, int cluster,
                        Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_on()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr, cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename T16>
int make_act_map(CUtensorMap* m, const ActView<T16>& v, int box_w = kTileW, int box_h = kTileH + 2) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return CCST_ECUDA;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)(v.W + 2), (cuuint64_t)(v.H + 2),
                              (cuuint64_t)v.N};
  const cuuint64_t strides[3] = {(cuuint64_t)v.C * 2, (cuuint64_t)(v.W + 2) * v.C * 2,
                                 (cuuint64_t)(v.H + 2) * (v.W + 2) * v.C * 2};
  // slab = the tile plus the two extra rows the filter rows r = 1, 2 reach into
  const cuuint32_t box[4] = {kBlockK, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, Fmt16<T16>::kTmaType, 4, (void*)v.p, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(activation %dx%dx%dx%d) failed: CUresult %d", v.N, v.H, v.W,
              v.C, (int)r);
    return CCST_ECUDA;
  }
  return CCST_OK;
}

// 4-D map over the INTERIOR of an activation (halo excluded, so TMA clips ragged tiles at the
// image border): dims (C, W/sx, H/sy, N) starting at interior pixel (oy, ox), pixel step (sy, sx).
template <typename T16>
int make_out_map(CUtensorMap* m, const ActView<T16>& v, int oy, int ox, int sy, int sx, int box_w,
                 int box_h) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return CCST_ECUDA;
  }
  const size_t pitch_y = (size_t)(v.W + 2) * v.C, pitch_n = (size_t)(v.H + 2) * (v.W + 2) * v.C;
  T16* base = v.p + (size_t)(1 + oy) * pitch_y + (size_t)(1 + ox) * v.C;
  const cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)((v.W - ox + sx - 1) / sx),
                              (cuuint64_t)((v.H - oy + sy - 1) / sy), (cuuint64_t)v.N};
  const cuuint64_t strides[3] = {(cuuint64_t)sx * v.C * 2, (cuuint64_t)sy * pitch_y * 2,
                                 (cuuint64_t)pitch_n * 2};
  const cuuint32_t box[4] = {kBlockK, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, Fmt16<T16>::kTmaType, 4, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(output %dx%dx%dx%d) failed: CUresult %d", v.N, v.H, v.W, v.C,
              (int)r);
    return CCST_ECUDA;
  }
  return CCST_OK;
}

template <typename T16>
int make_weight_map(CUtensorMap* m, const T16* wk, int K, int CoutPad, int BN) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return CCST_ECUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)CoutPad};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {kBlockK, (cuuint32_t)BN};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, Fmt16<T16>::kTmaType, 2, (void*)wk, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weights K=%d Cout=%d) failed: CUresult %d", K, CoutPad,
              (int)r);
    return CCST_ECUDA;
  }
  return CCST_OK;
}

int lin_desc_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CCST_LIN_DESC");
    v = e ? atoi(e) : 0;
  }
  return v;
}

template <typename T16, int BN, int EPI, int BRES, int CG, int GEO = 0>
int launch_cfg(const CUtensorMap& ma0, const T16* wk, ConvParams<T16> p, cudaStream_t st) {
  constexpr bool UPS = (EPI == EPI_UPS);
  using Cfg = UmmaCfg<BN, BRES, CG, UPS, GEO>;
  if constexpr (!Cfg::kFits) {
    set_error("conv_umma: configuration BN=%d resident=%d pair=%d geo=%d does not fit shared memory", BN,
              BRES, CG, GEO);
    return CCST_EINVAL;
  } else {
  using G = Geo<GEO>;
  CUtensorMap ma = ma0;
  if (GEO != 0) {
    ActView<T16> in{const_cast<T16*>(p.in_ptr), p.N, p.H, p.W, p.Cin};
    if (int e = make_act_map(&ma, in, G::kBoxW, G::kSlabRows)) return e;
    p.tiles_x = (p.W + G::kOutW - 1) / G::kOutW;
    p.tiles_y = (p.H + G::kRows - 1) / G::kRows;
    const int64_t m_tiles = (int64_t)p.N * p.tiles_x * p.tiles_y;
    CCST_CHECK_ARG(m_tiles * p.n_tiles * (UPS ? 4 : 1) < (1ll << 31), "conv_umma: too many tiles");
    p.m_tiles = (int)m_tiles;
    p.desc_mode = lin_desc_mode();
  }
  {
    static const int ablate = [] { const char* e = getenv("CCST_ABLATE"); return e ? atoi(e) : 0; }();
    p.ablate = ablate;
  }
  CUtensorMap mb;
  if (int e = make_weight_map(&mb, wk, Cfg::kTaps * p.Cin, (UPS ? 4 : 1) * p.CoutPad, Cfg::kBRows)) return e;
  OutMaps mo;
  memset(&mo, 0, sizeof(mo));
  if (EPI == EPI_ACT || EPI == EPI_ACT_STATS) {
    if (int e = make_out_map(&mo.m[0], p.out, 0, 0, 1, 1, G::kOutW, G::kRows)) return e;
  } else if (EPI == EPI_ACT_POOL) {
    if (int e = make_out_map(&mo.m[0], p.out, 0, 0, 1, 1, G::kOutW / 2, G::kRows / 2)) return e;
  } else if (EPI == EPI_ACT_UP2 || EPI == EPI_UPS) {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        if (int e = make_out_map(&mo.m[a * 2 + b], p.out, a, b, 2, 2, G::kOutW, G::kRows)) return e;
  }
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_umma_kernel<T16, BN, EPI, BRES, CG, GEO>), Cfg::kSmemBytes));
  const int64_t units = ((int64_t)p.m_tiles + CG - 1) / CG * p.n_tiles * (UPS ? 4 : 1);
  CCST_CHECK_ARG(units < (1ll << 31), "conv_umma: too many tiles");
  p.total_tiles = (int)units;
  int slots = sm_count() / CG;  // persistent: one CTA (or CTA pair) per SM (pair)
  if (UPS && BRES) slots &= ~3;  // resident weights of ONE phase per CTA: unit stride % 4 == 0
  const int grid = (int)(units < slots ? units : slots) * CG;
  CCST_CUDA(launch_conv(conv_umma_kernel<T16, BN, EPI, BRES, CG, GEO>, grid, kThreadsUmma, Cfg::kSmemBytes, st, CG,
                        ma, mb, mo, p));
  CCST_LAUNCHED();
  return CCST_OK;
  }
}

// tile geometry per N (see Geo): CCST_GEO64 / CCST_GEO128 = 0, 1 or 2
int geo_mode(int bn) {
  static int v64 = -1, v128 = -1;
  if (v64 < 0) {
    const char* e = getenv("CCST_GEO64");
    v64 = e ? atoi(e) : 0;
    e = getenv("CCST_GEO128");
    v128 = e ? atoi(e) : 0;
  }
  return bn == 64 ? v64 : (bn == 128 ? v128 : 0);
}

template <typename T16, int BN, int BRES, int CG>
int launch_bn(const CUtensorMap& ma, const T16* wk, const ConvParams<T16>& p, int epi,
              cudaStream_t st) {
  if (BN == 64 || BN == 128) {
    constexpr int BNL = (BN == 64 || BN == 128) ? BN : 64;  // only these are instantiated
    const int geo = geo_mode(BN);
    if (geo == 1) {
      switch (epi) {
        case EPI_ACT: return launch_cfg<T16, BNL, EPI_ACT, BRES, CG, 1>(ma, wk, p, st);
        case EPI_ACT_POOL: return launch_cfg<T16, BNL, EPI_ACT_POOL, BRES, CG, 1>(ma, wk, p, st);
        case EPI_UPS: return launch_cfg<T16, BNL, EPI_UPS, BRES, CG, 1>(ma, wk, p, st);
        default: break;
      }
    } else if (geo == 2) {
      switch (epi) {
        case EPI_ACT: return launch_cfg<T16, BNL, EPI_ACT, BRES, CG, 2>(ma, wk, p, st);
        case EPI_UPS: return launch_cfg<T16, BNL, EPI_UPS, BRES, CG, 2>(ma, wk, p, st);
        default: break;
      }
    }
  }
  switch (epi) {
    case EPI_ACT:
      return launch_cfg<T16, BN, EPI_ACT, BRES, CG>(ma, wk, p, st);
    case EPI_ACT_UP2:
      return launch_cfg<T16, BN, EPI_ACT_UP2, BRES, CG>(ma, wk, p, st);
    case EPI_ACT_POOL:
      return launch_cfg<T16, BN, EPI_ACT_POOL, BRES, CG>(ma, wk, p, st);
    case EPI_UPS:
      return launch_cfg<T16, BN, EPI_UPS, BRES, CG>(ma, wk, p, st);
    case EPI_ACT_STATS:
      if constexpr (BN == 256 && BRES == 0) return launch_cfg<T16, BN, EPI_ACT_STATS, BRES, CG>(ma, wk, p, st);
      set_error("conv_umma: tile statistics are available for the N = 256 tiles only");
      return CCST_EINVAL;
    default:
      set_error("conv_umma: epilogue %d not available for BN=%d", epi, BN);
      return CCST_EINVAL;
  }
}

// CTA pairs (cta_group::2) are the default for the N >= 128 layers (measured on B200, batch 32
// @512^2: N=256 layers 1.50 -> 1.67 PFLOP/s, N=128 layers 1.15 -> 1.27); the N=64 layers are
// slower paired (0.96 -> 0.82) and stay single-CTA.  CCST_CTA_PAIR=0 forces single-CTA kernels
// everywhere, CCST_CTA_PAIR=2 pairs everywhere (A/B measurements).
int cta_pair_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CCST_CTA_PAIR");
    v = (e && e[0] == '0') ? 0 : ((e && e[0] == '2') ? 2 : 1);
  }
  return v;
}

bool bres128_on() {
  // (the 16 KiB tiles of a single-CTA N = 128 kernel do not fit resident: CTA pairs only)
  static const bool on = [] { const char* e = getenv("CCST_BRES128"); return !(e && e[0] == '0'); }();
  return on && cta_pair_mode() != 0;
}

template <typename T16, int BN, int BRES>
int launch_cg(const CUtensorMap& ma, const T16* wk, const ConvParams<T16>& p, int epi,
              cudaStream_t st) {
  const int mode = cta_pair_mode();
  const bool pair = mode == 2 || (mode == 1 && BN >= 128);
  return pair ? launch_bn<T16, BN, BRES, 2>(ma, wk, p, epi, st)
              : launch_bn<T16, BN, BRES, 1>(ma, wk, p, epi, st);
}

// CCST_SMERGE=0 falls back to the tap-by-tap kernel for the 64-channel layers; =2 uses CTA pairs
int smerge_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CCST_SMERGE");
    v = e ? atoi(e) : 1;
  }
  return v;
}

template <typename T16, int EPI, bool BRES, int CG, int NG = 2>
int launch_smerge_cfg(const CUtensorMap& ma, const T16* wk_sm, ConvParams<T16> p, cudaStream_t st) {
  using Cfg = SmergeCfg<BRES, CG, NG>;
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
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_smerge_kernel<T16, EPI, BRES, CG, NG>), Cfg::kSmemBytes));
  {
    static const int ablate = [] { const char* e = getenv("CCST_ABLATE"); return e ? atoi(e) : 0; }();
    p.ablate = ablate;
  }
  p.tiles_x = (p.W + kSmOutW - 1) / kSmOutW;
  p.n_tiles = 1;
  const int64_t m_tiles = (int64_t)p.N * p.tiles_x * p.tiles_y;
  CCST_CHECK_ARG(m_tiles < (1ll << 31), "conv_smerge: too many tiles");
  p.m_tiles = (int)m_tiles;
  const int64_t units = (m_tiles + CG - 1) / CG;
  p.total_tiles = (int)units;
  const int slots = sm_count() / CG;
  const int grid = (int)(units < slots ? units : slots) * CG;
  CCST_CUDA(launch_conv(conv_smerge_kernel<T16, EPI, BRES, CG, NG>, grid, Cfg::kThreads, Cfg::kSmemBytes, st, CG, ma,
                        mb, mo, p));
  CCST_LAUNCHED();
  return CCST_OK;
}

template <typename T16, bool BRES, int CG>
int launch_smerge_epi(const CUtensorMap& ma, const T16* wk_sm, const ConvParams<T16>& p, int epi,
                      cudaStream_t st) {
  static const int groups = [] { const char* e = getenv("CCST_SMERGE_GROUPS"); return e ? atoi(e) : 2; }();
  if constexpr (CG == 2) {
    if (groups == 3) {
      if (epi == EPI_ACT) return launch_smerge_cfg<T16, EPI_ACT, BRES, CG, 3>(ma, wk_sm, p, st);
      if (epi == EPI_ACT_POOL) return launch_smerge_cfg<T16, EPI_ACT_POOL, BRES, CG, 3>(ma, wk_sm, p, st);
    }
  }
  switch (epi) {
    case EPI_ACT:
      return launch_smerge_cfg<T16, EPI_ACT, BRES, CG>(ma, wk_sm, p, st);
    case EPI_ACT_UP2:
      return launch_smerge_cfg<T16, EPI_ACT_UP2, BRES, CG>(ma, wk_sm, p, st);
    case EPI_ACT_POOL:
      return launch_smerge_cfg<T16, EPI_ACT_POOL, BRES, CG>(ma, wk_sm, p, st);
    default:
      set_error("conv_smerge: epilogue %d not available", epi);
      return CCST_EINVAL;
  }
}

template <typename T16>
int launch_smerge(const CUtensorMap& ma, const T16* wk_sm, const ConvParams<T16>& p, int epi,
                  cudaStream_t st) {
  // CTA pairs for the streamed-weight layers (dec7: 0.30 -> 0.26 ms; each CTA stages half of every
  // weight tile); CCST_SMERGE_PAIR=0/1 forces singles / pairs everywhere
  static const int pair_env = [] { const char* e = getenv("CCST_SMERGE_PAIR"); return e ? atoi(e) : -1; }();
  const bool pair = pair_env >= 0 ? pair_env != 0 : true;
  if (p.Cin == kBlockK)
    return pair ? launch_smerge_epi<T16, true, 2>(ma, wk_sm, p, epi, st)
                : launch_smerge_epi<T16, true, 1>(ma, wk_sm, p, epi, st);
  return pair ? launch_smerge_epi<T16, false, 2>(ma, wk_sm, p, epi, st)
              : launch_smerge_epi<T16, false, 1>(ma, wk_sm, p, epi, st);
}

}  // namespace

template <typename T16>
int launch_conv_umma(ActView<T16> in, const T16* wk, const T16* wk_sm, const T16* wk_up,
                     const float* bias, int Cout, int CoutPad, int relu, int epi, ActView<T16> out,
                     float* out_nchw, uint8_t* out_u8, int halo_edge, cudaStream_t st, float2* tile_stats) {
  CCST_CHECK_ARG(in.C % kBlockK == 0, "conv_umma: Cin=%d must be a multiple of 64", in.C);
  int BN;
  if (epi == EPI_NCHW_F32) {
    CCST_CHECK_ARG(CoutPad == 16 && Cout <= 16, "conv_umma: NCHW epilogue expects CoutPad == 16");
    BN = 16;
  } else {
    CCST_CHECK_ARG(CoutPad == Cout && Cout % 64 == 0, "conv_umma: Cout=%d must be a multiple of 64",
                   Cout);
    BN = Cout >= 256 ? 256 : Cout;  // 64, 128, 256
    CCST_CHECK_ARG(BN == 64 || BN == 128 || BN == 256, "conv_umma: unsupported Cout=%d", Cout);
  }
  ConvParams<T16> p;
  p.N = in.N, p.H = in.H, p.W = in.W, p.Cin = in.C;
  p.Cout = Cout, p.CoutPad = CoutPad;
  p.tiles_x = (in.W + kTileW - 1) / kTileW;
  p.tiles_y = (in.H + kTileH - 1) / kTileH;
  p.n_tiles = CoutPad / BN;
  const int64_t m_tiles = (int64_t)in.N * p.tiles_x * p.tiles_y;
  CCST_CHECK_ARG(m_tiles * p.n_tiles < (1ll << 31), "conv_umma: too many tiles");
  p.m_tiles = (int)m_tiles;
  p.total_tiles = 0;  // set per kernel variant (tiles or tile pairs)
  p.relu = relu;
  p.in_ptr = in.p;
  p.desc_mode = 0;
  p.ablate = 0;
  p.halo_edge = halo_edge;
  p.bias = bias;
  p.out = out;
  p.out_nchw = out_nchw;
  p.out_u8 = out_u8;
  p.tile_stats = tile_stats;
  CCST_CHECK_ARG((epi == EPI_ACT_STATS) == (tile_stats != nullptr), "conv_umma: tile_stats goes with EPI_ACT_STATS");
  CCST_CHECK_ARG(out_u8 == nullptr || epi == EPI_NCHW_F32, "conv_umma: uint8 store is the last conv's");
  CCST_CHECK_ARG(halo_edge == 1 || (halo_edge == 0 && epi == EPI_ACT),
                 "conv_umma: a replicate halo is only written by the plain epilogue");
  CUtensorMap ma;
  if (int e = make_act_map(&ma, in)) return e;
  if (epi == EPI_UPS) {
    // fused nearest-x2 upsample: `in` is the low-resolution map, `out` twice its size
    CCST_CHECK_ARG(wk_up != nullptr, "conv_umma: EPI_UPS needs the phase-packed weights");
    CCST_CHECK_ARG(out.H == 2 * in.H && out.W == 2 * in.W, "conv_umma: EPI_UPS output must be 2x the input");
    CCST_CHECK_ARG(m_tiles * p.n_tiles * 4 < (1ll << 31), "conv_umma: too many tiles");
    static const bool ups4_off = [] { const char* e = getenv("CCST_UPS4"); return e && e[0] == '0'; }();
    if (BN == 64 && in.C == kBlockK && Cout == 64 && !ups4_off) {
      // all four phases per tile over one linear slab (see conv_ups4_kernel)
      CUtensorMap m4, mb;
      if (int e = make_act_map(&m4, in, kU4BoxW, kU4Rows + 2)) return e;
      if (int e = make_weight_map(&mb, wk_up, 4 * in.C, 4 * Cout, 64)) return e;
      OutMaps mo;
      memset(&mo, 0, sizeof(mo));
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
          if (int e = make_out_map(&mo.m[a * 2 + b], out, a, b, 2, 2, kU4OutW, kU4Rows)) return e;
      CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_ups4_kernel<T16>), kU4Smem));
      p.tiles_x = (in.W + kU4OutW - 1) / kU4OutW;
      p.tiles_y = (in.H + kU4Rows - 1) / kU4Rows;
      const int64_t tiles = (int64_t)in.N * p.tiles_x * p.tiles_y;
      CCST_CHECK_ARG(tiles < (1ll << 31), "conv_ups4: too many tiles");
      p.m_tiles = p.total_tiles = (int)tiles;
      const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
      CCST_CUDA(launch_conv(conv_ups4_kernel<T16>, grid, kThreadsUmma, kU4Smem, st, 1, m4, mb, mo, p));
      CCST_LAUNCHED();
      return CCST_OK;
    }
    switch (BN) {
      case 64:
        return in.C == kBlockK ? launch_cg<T16, 64, true>(ma, wk_up, p, epi, st)
                               : launch_cg<T16, 64, false>(ma, wk_up, p, epi, st);
      case 128:
        // the 8 phase tiles (4 taps x 2 chunks) of the Cin = 128 layer stay resident
        return (in.C == 2 * kBlockK && bres128_on()) ? launch_cg<T16, 128, 2>(ma, wk_up, p, epi, st)
                                                     : launch_cg<T16, 128, 0>(ma, wk_up, p, epi, st);
      default:
        return launch_cg<T16, 256, false>(ma, wk_up, p, epi, st);
    }
  }
  switch (BN) {
    case 16: {
      // the last decoder conv (64 -> 3)
      CCST_CHECK_ARG(in.C == kBlockK, "conv_umma: the NCHW epilogue expects Cin == 64");
      static const bool gather_off = [] { const char* e = getenv("CCST_LAST_GATHER"); return e && e[0] == '0'; }();
      if (Cout > 3 || gather_off) return launch_cfg<T16, 16, EPI_NCHW_F32, true, 1>(ma, wk, p, st);
      static const bool rows_off = [] { const char* e = getenv("CCST_LAST_ROWS"); return e && e[0] == '0'; }();
      if (!rows_off) {
        // filter rows by operand shifts, filter columns in N + two shuffles (see conv_last_rows_kernel)
        CUtensorMap mr;
        if (int e = make_act_map(&mr, in, kRBoxW, kRRows + 2)) return e;
        CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_last_rows_kernel<T16>), kRSmem));
        p.tiles_x = (in.W + kROutW - 1) / kROutW;
        p.tiles_y = (in.H + kRRows - 1) / kRRows;
        const int64_t tiles = (int64_t)in.N * p.tiles_x * p.tiles_y;
        CCST_CHECK_ARG(tiles < (1ll << 31), "conv_last_rows: too many tiles");
        p.m_tiles = p.total_tiles = (int)tiles;
        const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
        CCST_CUDA(launch_conv(conv_last_rows_kernel<T16>, grid, kThreadsUmma, kRSmem, st, 1, mr, wk, p));
        CCST_LAUNCHED();
        return CCST_OK;
      }
      // taps in the N dimension + 9-point gather (see conv_last_umma_kernel)
      CUtensorMap ml;
      if (int e = make_act_map(&ml, in, kLSlabW, kLSlabH)) return e;
      CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_last_umma_kernel<T16>), kLSmem));
      p.total_tiles = p.m_tiles;
      const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
      conv_last_umma_kernel<T16><<<grid, kThreadsUmma, kLSmem, st>>>(ml, wk, p);
      CCST_LAUNCHED();
      return CCST_OK;
    }
    case 64:
      // s-merged kernel (N = 192 per operand view instead of 64) for every 64-channel layer: at N = 64
      // the tap-by-tap kernel is capped at 50 % of the tensor peak by the A-operand fetch.  Measured
      // batch 32 @512^2: conv1_2 (fused pool, pair) 0.635 -> 0.58 ms, dec7 (pair) 0.32 -> 0.25 ms.
      // CCST_SMERGE=0 disables it, =4 keeps the pooled layer on the tap-by-tap kernel.
      if (wk_sm && smerge_mode() != 0 && geo_mode(64) == 0 && (epi != EPI_ACT_POOL || smerge_mode() != 4))
        return launch_smerge<T16>(ma, wk_sm, p, epi, st);
      // 64 -> 64 layers keep all 9 weight tiles resident in shared memory
      return in.C == kBlockK ? launch_cg<T16, 64, true>(ma, wk, p, epi, st)
                             : launch_cg<T16, 64, false>(ma, wk, p, epi, st);
    case 128:
      // 64 -> 128 (conv2_1): the 9 weight tiles stay resident
      return (in.C == kBlockK && bres128_on()) ? launch_cg<T16, 128, 1>(ma, wk, p, epi, st)
                                               : launch_cg<T16, 128, 0>(ma, wk, p, epi, st);
    default:
      return launch_cg<T16, 256, false>(ma, wk, p, epi, st);
  }
}
// (one operand type per translation unit -- conv_umma_bf16.cu / conv_umma_f16.cu -- so that the two
// halves of the template instantiations compile in parallel)
#if CCST_INST_BF16
template int launch_conv_umma<__nv_bfloat16>(ActView<__nv_bfloat16>, const __nv_bfloat16*,
                                             const __nv_bfloat16*, const __nv_bfloat16*, const float*,
                                             int, int, int, int, ActView<__nv_bfloat16>, float*,
                                             uint8_t*, int, cudaStream_t, float2*);
#endif
#if CCST_INST_F16
template int launch_conv_umma<__half>(ActView<__half>, const __half*, const __half*, const __half*,
                                      const float*, int, int, int, int, ActView<__half>, float*,
                                      uint8_t*, int, cudaStream_t, float2*);
#endif

template <typename T16>
int launch_conv_first_umma(const float* img, int N, int H, int W, const T16* wk, const float* bias,
                           ActView<T16> out, cudaStream_t st) {
  FirstParams<T16> p;
  p.img = img, p.N = N, p.H = H, p.W = W, p.wk = wk, p.bias = bias, p.out = out;
  p.tiles_x = (W + kFirstPx - 1) / kFirstPx;
  const int64_t total = (int64_t)N * H * p.tiles_x;
  CCST_CHECK_ARG(total < (1ll << 31), "conv_first_umma: too many tiles");
  p.total_tiles = (int)total;
  CUtensorMap mo;
  if (int e = make_out_map(&mo, out, 0, 0, 1, 1, 32, 1)) return e;  // one warp's quarter
  static const bool ws_off = [] { const char* e = getenv("CCST_FIRST_WS"); return e && e[0] == '0'; }();
  if (W % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0 && !ws_off) {
    // rows are 16-byte aligned: TMA-fed warp-specialised kernel
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return CCST_ECUDA;
    }
    CUtensorMap mi;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
    const cuuint32_t box[4] = {kF2WinCols, 3, 3, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&mi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)img, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(image %dx3x%dx%d) failed: CUresult %d", N, H, W, (int)r);
      return CCST_ECUDA;
    }
    CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_first_umma_ws_kernel<T16>), kF2Smem));
    const int64_t cap2 = (int64_t)sm_count() * 2;
    const int grid2 = (int)(total < cap2 ? total : cap2);
    CCST_CUDA(launch_conv(conv_first_umma_ws_kernel<T16>, grid2, kF2Threads, kF2Smem, st, 1, mi, mo, p));
    CCST_LAUNCHED();
    return CCST_OK;
  }
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_first_umma_kernel<T16>), kFirstSmem));
  const int64_t cap = (int64_t)sm_count() * 4;
  const int grid = (int)(total < cap ? total : cap);
  conv_first_umma_kernel<T16><<<grid, kFirstPx, kFirstSmem, st>>>(mo, p);
  CCST_LAUNCHED();
  return CCST_OK;
}
#if CCST_INST_BF16
template int launch_conv_first_umma<__nv_bfloat16>(const float*, int, int, int,
                                                   const __nv_bfloat16*, const float*,
                                                   ActView<__nv_bfloat16>, cudaStream_t);
#endif
#if CCST_INST_F16
template int launch_conv_first_umma<__half>(const float*, int, int, int, const __half*,
                                            const float*, ActView<__half>, cudaStream_t);
#endif

}  // namespace ccst
