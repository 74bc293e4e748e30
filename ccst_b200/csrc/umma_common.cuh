// tcgen05 / TMEM / TMA building blocks shared by the convolution kernels (sm_100a): PTX wrappers,
// shared-memory / instruction descriptors, tensor-map construction and the launch helper.
#pragma once
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <utility>

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

// Measurement switches (skip pipeline stages, programmatic dependent launch) exist only in builds
// made with -DCCST_DEV (tools/); the shipped library reads nothing from the environment.
#ifdef CCST_DEV
inline int dev_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
#define CCST_ABLATE_BITS(p) ((p).ablate)
#else
#define CCST_ABLATE_BITS(p) 0
#endif

template <typename T16>
struct ConvParams {
  int N, H, W, Cin;
  int Cout, CoutPad;
  int tiles_x, tiles_y, n_tiles, m_tiles;  // m_tiles = pixel tiles (N * tiles_y * tiles_x)
  int total_tiles;                        // work units: (pixel tile | pair of pixel tiles) x n_tiles
  int relu;
  int halo_edge;  // halo written around `out`: 1 reflection, 0 replicate (for_each_halo_alias)
  // per-sample weights (AdaIN folded into the conv that follows it, dec1): image n reads the weight
  // rows n * w_rows_per_n + ... of the weight map and the bias bias[n * bias_per_n + co]; both 0 otherwise
  int w_rows_per_n, bias_per_n;
  int m_tiles_img;  // pixel tiles per image (padded to even for CTA pairs when weights are per sample)
  int contig_units; // s-merged kernel: every CTA takes a contiguous range of work units instead of a strided one
  // f16x3 engine (conv_x3.cuh): the accumulator is multiplied by `out_scale`, the exact power of two the
  // split weights were scaled by, before the bias is added
  float out_scale;
  const float* bias;
  ActView<T16> out;
  float* out_nchw;
  uint8_t* out_u8;  // last conv only: NHWC uint8 store quantised like torchvision's save_image
  float2* tile_stats;  // EPI_ACT_STATS: [(pixel tile * 4 + row quarter) * Cout + channel] {mean, M2}
  unsigned int* sat_count;  // f16 stores that hit the +-65504 clamp (one atomic per CTA at exit); may be NULL
#ifdef CCST_DEV
  int ablate;  // 1 skip the epilogue's work, 2 skip the MMAs, 4 skip the A loads (tools/ablate.sh)
#endif
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
// kind::f16 instruction descriptor: D = fp32, A = B = bf16 or f16, both K-major, M = 128, N = BN
template <typename T16, int BN, int CG = 1>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) /*D fp32*/ | (Fmt16<T16>::kIdescFmt << 7) /*A*/ | (Fmt16<T16>::kIdescFmt << 10) /*B*/ |
         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((CG * kBlockM) >> 4) << 24);
}

// two fp32 additions in one instruction (FADD2); each half rounds like a scalar add.rn.f32
__device__ __forceinline__ float2 add2_f32(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

struct TileCoord {
  int n, y0, x0, nt, ph;
  int nw;  // image whose weights the tile uses (per-sample weights only)
};

// f16 activations are stored with saturation (+-65504) instead of overflowing to inf; weights with a
// larger dynamic range than the f16 format silently lose information there, so every epilogue keeps
// the running maximum |value| of what it packs (one LOP3 + one HMNMX2 per packed pair) and reports
// the threads that hit the clamp with one atomic per warp at exit.  bf16 has the fp32 range: no-op.
template <typename T16>
struct SatTracker {
  __device__ __forceinline__ void track(uint32_t) {}
  __device__ __forceinline__ void track_nonneg(uint32_t) {}
  template <int N>
  __device__ __forceinline__ void track_block(const uint32_t (&)[N], bool) {}
  __device__ __forceinline__ void flush(unsigned int*) {}
};
template <>
struct SatTracker<__half> {
  uint32_t m = 0u;
  __device__ __forceinline__ void track(uint32_t w) { m = max16x2<__half>(m, w & 0x7fff7fffu); }
  // values known to be >= 0 (packed with ReLU): one instruction
  __device__ __forceinline__ void track_nonneg(uint32_t w) { m = max16x2<__half>(m, w); }
  // a block of N packed words: four independent chains instead of one N-deep HMNMX2 chain, and the sign mask only
  // when the values may be negative (`nonneg` is warp-uniform: one branch per block, not one LOP3 per word)
  template <int N>
  __device__ __forceinline__ void track_block(const uint32_t (&w)[N], bool nonneg) {
    uint32_t a[4] = {0u, 0u, 0u, 0u};
    if (nonneg) {
#pragma unroll
      for (int j = 0; j < N; ++j) a[j & 3] = max16x2<__half>(a[j & 3], w[j]);
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) a[j & 3] = max16x2<__half>(a[j & 3], w[j] & 0x7fff7fffu);
    }
    m = max16x2<__half>(max16x2<__half>(m, a[0]), max16x2<__half>(max16x2<__half>(a[1], a[2]), a[3]));
  }
  __device__ __forceinline__ void flush(unsigned int* counter) {
    // bit patterns of non-negative halves order like the values; 0x7bff = 65504, above = inf / NaN
    const bool hit = (m & 0xffffu) >= 0x7bffu || (m >> 16) >= 0x7bffu;
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (ballot != 0u && counter != nullptr && (threadIdx.x & 31) == 0) atomicAdd(counter, (unsigned)__popc(ballot));
  }
};

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
  // cheap next to the tails, so it is off in the shipped library (CCST_DEV builds: CCST_PDL=1)
#ifdef CCST_DEV
  static const bool on = dev_env_int("CCST_PDL", 0) == 1;
  return on;
#else
  return false;
#endif
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

}  // namespace
}  // namespace ccst
