// Per-plane feature statistics and fused AdaIN on NCHW fp32 tensors (the reference's layout).
//
// Replaces: calc_mean_std (function.py:4-13), adaIN_StyleStat_ContentFeat (function.py:26-33),
// adaptive_instance_normalization (function.py:16-24), calc_sum + accumulation + finalise
// (mean_std_computation_effcientMem.py:103-137).
//
// Both kernels are HBM-bound.  One (n,c) plane is HW contiguous floats; a *group* of G threads
// (a warp for small planes, a 256-thread CTA for 64x64 planes) loads its plane with 128-bit
// streaming loads and keeps it in registers, so statistics and the affine re-normalisation cost
// exactly one HBM read (+ one write for AdaIN).  Each thread reduces its registers to an exact
// local (count, mean, M2) triple, the triples are merged with Chan's formula by warp shuffles and
// (for CTA groups) a shared-memory combine.  Planes that do not fit in registers, or whose size /
// alignment rules out 128-bit access, take a streaming variant of the same algorithm.
#include <cstdlib>
#include <dlfcn.h>

#include "common.cuh"

namespace ccst {

namespace {

constexpr int kThreads = 256;

enum StatOut { OUT_MEAN_STD = 0, OUT_MEAN_M2 = 1 };

struct StatsArgs {
  const float* x;
  int64_t planes;
  int64_t hw;
  float eps;
  int unbiased;
  // OUT_MEAN_STD
  float* mean;
  float* stdv;
  // OUT_MEAN_M2: float2 {mean, M2} per plane
  float2* raw;
};

struct AdainArgs {
  const float* x;
  float* out;
  int64_t planes;
  int64_t hw;
  int C;
  const float* mu_s;
  const float* sigma_s;
  int64_t stat_batch_stride;  // 0: stats are [C]; C: stats are [N,C]
  float alpha;
  float eps;
};

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// Merge across the G threads of a group.  G == 32: pure shuffles.  G == 256: shuffles, then the 8
// warp results are exchanged through shared memory and merged in a fixed order by every thread.
template <int G>
__device__ __forceinline__ Wf group_reduce(Wf v, Wf* smem /* >= 8 entries, only G==256 */) {
  v = wf_warp_reduce(v);
  if (G == 32) return v;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect smem reuse across loop iterations
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  Wf r = smem[0];
#pragma unroll
  for (int w = 1; w < G / 32; ++w) r = wf_merge(r, smem[w]);
  return r;
}

__device__ __forceinline__ float finish_std(Wf s, float eps, int unbiased) {
  // unbiased with n == 1 gives 0/0 = NaN exactly like torch.var (function.py:9)
  float denom = unbiased ? (s.n - 1.f) : s.n;
  return sqrtf(s.m2 / denom + eps);
}

// Register-resident plane: thread holds V float4 (element index = (i*G + t)*4 .. +3).
template <int G, int V>
struct PlaneRegs {
  float4 v[V];
  int cnt;  // valid elements held by this thread

  __device__ __forceinline__ void load(const float* plane, int64_t hw, int t) {
    const float4* p4 = reinterpret_cast<const float4*>(plane);
    const int n4 = (int)(hw >> 2);
    cnt = 0;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      int idx = i * G + t;
      if (idx < n4) {
        v[i] = ld_stream(p4 + idx);
        cnt += 4;
      } else {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }

  __device__ __forceinline__ Wf local() const {
    const int n4 = cnt >> 2;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
      if (i < n4) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    Wf r;
    r.n = (float)cnt;
    r.mean = cnt ? s / (float)cnt : 0.f;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
      if (i < n4) {
        float a = v[i].x - r.mean, b = v[i].y - r.mean, c = v[i].z - r.mean, d = v[i].w - r.mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    r.m2 = q;
    return r;
  }
};

template <int G, int V, int OUT>
__global__ void __launch_bounds__(kThreads) stats_regs_kernel(StatsArgs a) {
  __shared__ Wf smem[8];
  constexpr int kGroups = kThreads / G;
  const int t = threadIdx.x % G;
  const int g = threadIdx.x / G;
  for (int64_t plane = (int64_t)blockIdx.x * kGroups + g; plane < a.planes + (G == 256 ? 0 : 0);
       plane += (int64_t)gridDim.x * kGroups) {
    PlaneRegs<G, V> regs;
    regs.load(a.x + plane * a.hw, a.hw, t);
    Wf s = group_reduce<G>(regs.local(), smem);
    if (t == 0) {
      if (OUT == OUT_MEAN_STD) {
        if (a.mean) a.mean[plane] = s.mean;
        if (a.stdv) a.stdv[plane] = finish_std(s, a.eps, a.unbiased);
      } else {
        a.raw[plane] = make_float2(s.mean, s.m2);
      }
    }
  }
}

// Streaming Welford over one plane per CTA (any hw, any alignment).
__device__ __forceinline__ Wf stream_plane(const float* plane, int64_t hw, Wf* smem) {
  // peel to 16-byte alignment
  const uintptr_t addr = reinterpret_cast<uintptr_t>(plane);
  int64_t head = ((16 - (addr & 15)) & 15) >> 2;
  if (head > hw) head = hw;
  const int64_t n4 = (hw - head) >> 2;
  const int64_t tail0 = head + (n4 << 2);
  Wf acc{0.f, 0.f, 0.f};
  const float4* p4 = reinterpret_cast<const float4*>(plane + head);
  for (int64_t i = threadIdx.x; i < n4; i += kThreads) {
    float4 v = ld_stream(p4 + i);
    Wf c;
    c.n = 4.f;
    c.mean = 0.25f * ((v.x + v.y) + (v.z + v.w));
    float d0 = v.x - c.mean, d1 = v.y - c.mean, d2 = v.z - c.mean, d3 = v.w - c.mean;
    c.m2 = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    acc = wf_merge(acc, c);
  }
  // scalar head + tail elements
  for (int64_t i = threadIdx.x; i < head + (hw - tail0); i += kThreads) {
    int64_t e = i < head ? i : tail0 + (i - head);
    Wf c{1.f, __ldg(plane + e), 0.f};
    acc = wf_merge(acc, c);
  }
  return group_reduce<256>(acc, smem);
}

template <int OUT>
__global__ void __launch_bounds__(kThreads) stats_stream_kernel(StatsArgs a) {
  __shared__ Wf smem[8];
  for (int64_t plane = blockIdx.x; plane < a.planes; plane += gridDim.x) {
    Wf s = stream_plane(a.x + plane * a.hw, a.hw, smem);
    if (threadIdx.x == 0) {
      if (OUT == OUT_MEAN_STD) {
        if (a.mean) a.mean[plane] = s.mean;
        if (a.stdv) a.stdv[plane] = finish_std(s, a.eps, a.unbiased);
      } else {
        a.raw[plane] = make_float2(s.mean, s.m2);
      }
    }
  }
}

// out = (x - mu_c) * A + B,  A = alpha*sigma_s/sigma_c + (1-alpha),  B = alpha*mu_s + (1-alpha)*mu_c
__device__ __forceinline__ void adain_coeffs(const AdainArgs& a, int64_t plane, Wf s, float& A,
                                             float& B) {
  const int64_t n = plane / a.C;
  const int c = (int)(plane - n * a.C);
  const int64_t si = n * a.stat_batch_stride + c;
  const float mu_s = __ldg(a.mu_s + si), sg_s = __ldg(a.sigma_s + si);
  const float sg_c = finish_std(s, a.eps, /*unbiased=*/1);
  A = a.alpha * (sg_s / sg_c) + (1.f - a.alpha);
  B = a.alpha * mu_s + (1.f - a.alpha) * s.mean;
}

template <int G, int V>
__global__ void __launch_bounds__(kThreads) adain_regs_kernel(AdainArgs a) {
  __shared__ Wf smem[8];
  constexpr int kGroups = kThreads / G;
  const int t = threadIdx.x % G;
  const int g = threadIdx.x / G;
  for (int64_t plane = (int64_t)blockIdx.x * kGroups + g; plane < a.planes;
       plane += (int64_t)gridDim.x * kGroups) {
    PlaneRegs<G, V> regs;
    regs.load(a.x + plane * a.hw, a.hw, t);
    Wf s = group_reduce<G>(regs.local(), smem);
    float A, B;
    adain_coeffs(a, plane, s, A, B);
    float4* o4 = reinterpret_cast<float4*>(a.out + plane * a.hw);
    const int n4 = (int)(a.hw >> 2);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      int idx = i * G + t;
      if (idx < n4) {
        float4 v = regs.v[i];
        v.x = fmaf(v.x - s.mean, A, B);
        v.y = fmaf(v.y - s.mean, A, B);
        v.z = fmaf(v.z - s.mean, A, B);
        v.w = fmaf(v.w - s.mean, A, B);
        st_stream(o4 + idx, v);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads) adain_stream_kernel(AdainArgs a) {
  __shared__ Wf smem[8];
  for (int64_t plane = blockIdx.x; plane < a.planes; plane += gridDim.x) {
    const float* x = a.x + plane * a.hw;
    float* o = a.out + plane * a.hw;
    Wf s = stream_plane(x, a.hw, smem);
    float A, B;
    adain_coeffs(a, plane, s, A, B);
    // second pass: the plane was just read, so this mostly hits L2
    for (int64_t i = threadIdx.x; i < a.hw; i += kThreads) o[i] = fmaf(x[i] - s.mean, A, B);
  }
}

// ---- per-channel merge over the batch and into the running state (fp64) ----
// state = {count, mean[C], M2[C]}.  Block = 32 channels x 8 batch lanes; each thread Chan-merges its
// strided share of the N per-plane results, the 8 partials are combined through shared memory.
// Folds the per-plane {mean, M2} of a batch over n and into the running fp64 state (Chan), fixed
// order.  Block = 8 channels (one per warp) x 32 lanes over n: lane l merges planes l, l + 32, ...,
// the lanes are combined by a shuffle butterfly, lane 0 merges the result into the state.
__device__ __forceinline__ void chan_merge(double& n, double& mean, double& m2, double nb, double mb,
                                           double m2b) {
  if (nb == 0.0) return;
  const double nn = n + nb, d = mb - mean;
  mean += d * (nb / nn);
  m2 += m2b + d * d * (n * nb / nn);
  n = nn;
}

// The running state is 2 + 2C doubles: {count, mean[C], M2[C], ticket}.  Every block of an accumulating
// kernel reads the OLD count, merges its channels, then takes a ticket (atomicInc wraps to 0 at the
// grid size, so the word is zero again afterwards); the block that draws the last ticket -- by then
// every block has read the old count -- stores the new count.  No second launch, no scratch counter.
__device__ __forceinline__ void bump_count_last_block(double* state, int C, double old_count, double add) {
  __threadfence();
  unsigned int* ticket = reinterpret_cast<unsigned int*>(state + 1 + 2 * C);
  if (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1) state[0] = old_count + add;
}

// Block = 8 channels x 32 batch lanes (thread t: channel t & 7, lane t >> 3): a warp reads 4 batch rows x 8
// adjacent channels (64 contiguous bytes per row) per step, every thread Chan-merges N / 32 planes, the 32
// lanes of a channel are combined through shared memory in fixed order.  (One warp per channel with a
// 32-long dependent fp64 chain per lane and 4 KiB-strided loads took 20 us at [1024,512]; this form is
// bound by the loads.)
__global__ void __launch_bounds__(256) merge_planes_kernel(const float2* __restrict__ raw, int N,
                                                            int C, double hw,
                                                            double* __restrict__ state) {
  __shared__ double s_n[8][8], s_mean[8][8], s_m2[8][8];
  const int ch = threadIdx.x & 7, nl = threadIdx.x >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 8 + ch;
  const double old_count = state[0];
  double n = 0.0, mean = 0.0, m2 = 0.0;
  if (c < C) {
    int i = nl;
    for (; i + 96 < N; i += 128) {  // four independent loads in flight
      const float2 r0 = raw[(size_t)i * C + c], r1 = raw[(size_t)(i + 32) * C + c];
      const float2 r2 = raw[(size_t)(i + 64) * C + c], r3 = raw[(size_t)(i + 96) * C + c];
      chan_merge(n, mean, m2, hw, (double)r0.x, (double)r0.y);
      chan_merge(n, mean, m2, hw, (double)r1.x, (double)r1.y);
      chan_merge(n, mean, m2, hw, (double)r2.x, (double)r2.y);
      chan_merge(n, mean, m2, hw, (double)r3.x, (double)r3.y);
    }
    for (; i < N; i += 32) {
      const float2 r = raw[(size_t)i * C + c];
      chan_merge(n, mean, m2, hw, (double)r.x, (double)r.y);
    }
  }
  // the warp's four batch lanes of a channel (lanes ch, ch + 8, ch + 16, ch + 24): butterfly in which both
  // partners compute (lower lane) <- (upper lane), so the result does not depend on who keeps it
#pragma unroll
  for (int off = 8; off < 32; off <<= 1) {
    const double nb = __shfl_xor_sync(0xffffffffu, n, off);
    const double mb = __shfl_xor_sync(0xffffffffu, mean, off);
    const double qb = __shfl_xor_sync(0xffffffffu, m2, off);
    double an = n, am = mean, aq = m2, bn = nb, bm = mb, bq = qb;
    if (lane & off) an = nb, am = mb, aq = qb, bn = n, bm = mean, bq = m2;
    chan_merge(an, am, aq, bn, bm, bq);
    n = an, mean = am, m2 = aq;
  }
  if (lane < 8) s_n[warp][ch] = n, s_mean[warp][ch] = mean, s_m2[warp][ch] = m2;
  __syncthreads();
  if (threadIdx.x < 8 && c < C) {
    double an = s_n[0][ch], am = s_mean[0][ch], aq = s_m2[0][ch];
    for (int l = 1; l < 8; ++l) chan_merge(an, am, aq, s_n[l][ch], s_mean[l][ch], s_m2[l][ch]);
    double sn = old_count, sm = state[1 + c], sq = state[1 + C + c];
    chan_merge(sn, sm, sq, an, am, aq);
    state[1 + c] = sm;
    state[1 + C + c] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) bump_count_last_block(state, C, old_count, hw * (double)N);
}

__global__ void finalize_kernel(const double* __restrict__ state, int C, float eps, int unbiased,
                                float* __restrict__ mean, float* __restrict__ stdv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = state[0] - (unbiased ? 1.0 : 0.0);  // n == 1, unbiased: 0/0 = NaN as torch.var
  if (mean) mean[c] = (float)state[1 + c];
  if (stdv) stdv[c] = (float)sqrt(state[1 + C + c] / n + (double)eps);
}

__global__ void to_sums_kernel(const double* __restrict__ state, int C, float* __restrict__ sum,
                               float* __restrict__ sqsum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = state[0], m = state[1 + c];
  if (sum) sum[c] = (float)(n * m);
  if (sqsum) sqsum[c] = (float)(state[1 + C + c] + n * m * m);
}

__global__ void to_moments_kernel(const double* __restrict__ state, int C,
                                  double* __restrict__ mom) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const double n = state[0];
  if (c == 0) mom[0] = n;
  if (c >= C) return;
  const double m = state[1 + c];
  mom[1 + c] = n * m;
  mom[1 + C + c] = state[1 + C + c] + n * m * m;
}

__global__ void from_moments_kernel(const double* __restrict__ mom, int C,
                                    double* __restrict__ state) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const double n = mom[0];
  if (c == 0) state[0] = n;
  if (c >= C) return;
  const double m = n > 0 ? mom[1 + c] / n : 0.0;
  state[1 + c] = m;
  double m2 = mom[1 + C + c] - n * m * m;
  state[1 + C + c] = m2 > 0 ? m2 : 0.0;
}

// =====================================================================================
// Bulk-staged variant (the fast path): planes are contiguous in memory, so a CTA streams chunks of
// consecutive planes (4 .. 16 KiB) into a shared-memory ring with 1-D TMA bulk copies (cp.async.bulk +
// mbarrier complete_tx), 128 .. 200 KiB in flight per SM.  Each consumer warp owns its ring slots and
// reduces a chunk from shared memory with an exact two-pass (sum, then sum of squared deviations;
// warp shuffles combine the lanes of a plane), then -- for AdaIN -- makes a third pass that applies
// the affine and writes 128-bit coalesced stores.  HBM sees exactly one read (+ one write).
// =====================================================================================
constexpr int kRingBytes = 131072;     // ring of the one-launch accumulation (welford_bulk_kernel)
constexpr int kSlotBytesMax = 16384;
// plane_bulk_kernel sizes its ring at launch: what bounds the statistics kernels on small inputs is the number
// of bytes in flight per SM (ncu on [64,512,28,28] with a 16 x 8 KiB ring, 77 % filled: the consumer warps
// spend 44 % of their samples waiting for a full slot, DRAM at 49 %), so the ring takes all the shared
// memory an SM has and its slots are whole multiples of the plane size.
constexpr int kRingBudget = 225 * 1024;
constexpr int kMaxSlots = 32;
constexpr int kMaxConsumers = 16;

struct BulkArgs {
  const float* x;
  float* out;      // AdaIN only
  int64_t planes;
  int hw;          // elements per plane, hw % 4 == 0, hw * 4 <= kSlotBytesMax
  int ppc;         // planes per chunk
  int slots;       // ring slots of ppc * hw * 4 bytes (<= kMaxSlots)
  int warps;       // consumer warps (<= kMaxConsumers): warp w takes the CTA's chunks w, w + warps, ...
  int64_t chunks;
  float eps;
  int unbiased;
  float* mean;     // MODE 0
  float* stdv;
  float2* raw;     // MODE 1
  int C;           // MODE 2 (AdaIN)
  const float* mu_s;
  const float* sigma_s;
  int64_t stat_batch_stride;
  float alpha;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ bool bar_try_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  if (bar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!bar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // never hang the GPU on a pipeline bug
      printf("ccst stats: mbarrier timeout block=%d thread=%d\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// G lanes cooperate on one plane (G = 8 for planes <= 4 KiB, else 32)
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// MODE 0: mean/std, 1: raw {mean, M2}, 2: AdaIN
// Ring of a.slots slots of a.ppc planes, a.warps consumer warps; the CTA's chunks are blockIdx.x,
// blockIdx.x + gridDim.x, ..., consumer warp w takes every a.warps-th of them and owns the slots
// w * depth .. w * depth + depth - 1 (depth = slots / warps), so a slot is always read by the same warp.
//   PROD = false (statistics): self-service -- every warp issues the bulk copies of its first `depth` chunks
//     itself and, after reducing a chunk, the copy of the chunk that reuses the slot.  A single producer lane
//     was the bound of the read-only modes on every shape with chunks below 16 KiB (ncu on [64,512,28,28]:
//     consumers 44 % of their samples waiting for a full slot, DRAM at 49 %); with the issue work spread
//     over 16 warps [1024,512,12,12] went from 4.9 to 6.0 TB/s, [64,512,28,28] from 3.7 to 4.5.
//   PROD = true (AdaIN): warp 0 is a dedicated producer running ahead through empty barriers; the consumers
//     spend most of their time storing, and loads issued only when a warp gets around to it leave the read
//     side idle (measured 5 .. 8 % slower on the 28^2 / 32^2 shapes).
// Small planes are bound by the per-plane latency chain (two shared-memory passes, shuffles, sqrt): they get
// 16 warps and few lanes per plane.
template <int MODE, int G, bool PROD>
__global__ void __launch_bounds__(32 * (kMaxConsumers + (PROD ? 1 : 0)), 1) plane_bulk_kernel(BulkArgs a) {
  extern __shared__ __align__(128) uint8_t ring[];  // slots * slot_bytes, then the barriers (full, empty)
  const int kSlots = a.slots, W = a.warps;
  const int kDepth = kSlots / W;  // slots (chunks in flight) per consumer warp
  const int kSlotBytes = a.ppc * a.hw * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)kSlots * kSlotBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bars[s])));  // full
      if (PROD) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bars[kSlots + s])));  // empty
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t plane_bytes = (int64_t)a.hw * 4;
  // bulk copy of chunk c into `slot` (one lane)
  auto issue = [&](int64_t c, int slot) {
    const int64_t p0 = c * a.ppc;
    const int64_t np = (a.planes - p0) < a.ppc ? (a.planes - p0) : a.ppc;
    const uint32_t bytes = (uint32_t)(np * plane_bytes);
    const uint32_t full = smem_addr(&bars[slot]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(ring + (size_t)slot * kSlotBytes)),
        "l"(a.x + p0 * a.hw), "r"(bytes), "r"(full)
        : "memory");
  };
  const int64_t cstep = (int64_t)W * gridDim.x;
  if (PROD && warp == 0) {
    // the CTA's i-th chunk belongs to warp i % W, round i / W: slot (i % W) * depth + (i / W) % depth
    if (lane == 0) {
      int cw = 0, round = 0;
      for (int64_t c = blockIdx.x; c < a.chunks; c += gridDim.x) {
        const int slot = cw * kDepth + round % kDepth;
        bar_wait(smem_addr(&bars[kSlots + slot]), (((uint32_t)(round / kDepth)) & 1) ^ 1);
        issue(c, slot);
        if (++cw == W) cw = 0, ++round;
      }
    }
    return;
  }
  const int w = warp - (PROD ? 1 : 0);
  const int64_t c_first = (int64_t)blockIdx.x + (int64_t)w * gridDim.x;
  if (!PROD) {
    if (lane == 0)
      for (int d0 = 0; d0 < kDepth; ++d0)
        if (c_first + d0 * cstep < a.chunks) issue(c_first + d0 * cstep, w * kDepth + d0);
    __syncwarp();
  }
  const int n4 = a.hw >> 2;              // float4 per plane
  constexpr int kGroups = 32 / G;        // planes processed concurrently by the warp
  const int grp = lane / G, sub = lane % G;
  int d = 0;          // slot of this warp the current chunk sits in
  uint32_t use = 0;   // how often the warp has been around its slots
  for (int64_t c = c_first; c < a.chunks; c += cstep) {
    const int slot = w * kDepth + d;
    const float4* buf = reinterpret_cast<const float4*>(ring + (size_t)slot * kSlotBytes);
    bar_wait(smem_addr(&bars[slot]), use & 1);
    const int64_t p0 = c * a.ppc;
    const int np = (int)((a.planes - p0) < a.ppc ? (a.planes - p0) : a.ppc);
    for (int pb = 0; pb < np; pb += kGroups) {
      const int pl = pb + grp;  // plane inside the chunk handled by this lane group
      const bool active = pl < np;
      const float4* src = buf + (size_t)(active ? pl : 0) * n4;
      // pass 1: mean (4 independent accumulators hide the shared-memory latency)
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      if (active) {
        int k = sub;
        for (; k + 3 * G < n4; k += 4 * G) {
          const float4 v0 = src[k], v1 = src[k + G], v2 = src[k + 2 * G], v3 = src[k + 3 * G];
          s0 += (v0.x + v0.y) + (v0.z + v0.w);
          s1 += (v1.x + v1.y) + (v1.z + v1.w);
          s2 += (v2.x + v2.y) + (v2.z + v2.w);
          s3 += (v3.x + v3.y) + (v3.z + v3.w);
        }
        for (; k < n4; k += G) {
          const float4 v = src[k];
          s0 += (v.x + v.y) + (v.z + v.w);
        }
      }
      const float mean = group_sum<G>((s0 + s1) + (s2 + s3)) / (float)a.hw;
      // pass 2: sum of squared deviations (exact two-pass, no cancellation)
      float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
      if (active) {
        auto sq = [mean](const float4 v) {
          const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
          return (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        };
        int k = sub;
        for (; k + 3 * G < n4; k += 4 * G) {
          const float4 v0 = src[k], v1 = src[k + G], v2 = src[k + 2 * G], v3 = src[k + 3 * G];
          q0 += sq(v0);
          q1 += sq(v1);
          q2 += sq(v2);
          q3 += sq(v3);
        }
        for (; k < n4; k += G) q0 += sq(src[k]);
      }
      const float q = group_sum<G>((q0 + q1) + (q2 + q3));
      const int64_t plane = p0 + pl;
      if (MODE == 0) {
        if (active && sub == 0) {
          const float denom = a.unbiased ? (float)(a.hw - 1) : (float)a.hw;
          if (a.mean) a.mean[plane] = mean;
          if (a.stdv) a.stdv[plane] = sqrtf(q / denom + a.eps);  // hw == 1, unbiased: 0/0 = NaN (function.py:9)
        }
      } else if (MODE == 1) {
        if (active && sub == 0) a.raw[plane] = make_float2(mean, q);
      } else {
        if (active) {
          const int64_t n = plane / a.C;
          const int ch = (int)(plane - n * a.C);
          const int64_t si = n * a.stat_batch_stride + ch;
          const float sg_c = sqrtf(q / (float)(a.hw - 1) + a.eps);
          const float A = a.alpha * (__ldg(a.sigma_s + si) / sg_c) + (1.f - a.alpha);
          const float B = a.alpha * __ldg(a.mu_s + si) + (1.f - a.alpha) * mean;
          float4* dst = reinterpret_cast<float4*>(a.out + plane * a.hw);
          auto tf = [mean, A, B](float4 v) {
            v.x = fmaf(v.x - mean, A, B);
            v.y = fmaf(v.y - mean, A, B);
            v.z = fmaf(v.z - mean, A, B);
            v.w = fmaf(v.w - mean, A, B);
            return v;
          };
          int k = sub;
          for (; k + 3 * G < n4; k += 4 * G) {
            const float4 v0 = src[k], v1 = src[k + G], v2 = src[k + 2 * G], v3 = src[k + 3 * G];
            st_stream(dst + k, tf(v0));
            st_stream(dst + k + G, tf(v1));
            st_stream(dst + k + 2 * G, tf(v2));
            st_stream(dst + k + 3 * G, tf(v3));
          }
          for (; k < n4; k += G) st_stream(dst + k, tf(src[k]));
        }
      }
    }
    // every lane's generic-proxy reads of the slot -> ordered before the bulk copy (async proxy) that refills it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      if (PROD) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&bars[kSlots + slot])) : "memory");
      else if (c + kDepth * cstep < a.chunks) issue(c + kDepth * cstep, slot);
    }
    if (++d == kDepth) d = 0, ++use;
  }
}

// =====================================================================================
// Welford accumulation in ONE launch (calc_sum + `all_* += ...`, mean_std_computation_effcientMem.py:
// 103-131): the bulk-staged ring of plane_bulk_kernel, but with the planes dealt to the CTAs so that a
// CTA owns whole CHANNELS: chunk = PPC consecutive planes (channels j*PPC .. of one image), and CTA b
// takes the chunk classes j = b, b + grid, ... of EVERY image (grid divides C / PPC).  The per-plane
// {mean, M2} never leave the SM: each consumer warp keeps a running fp64 Chan state per channel over its
// images, the warps of a CTA are merged through shared memory in fixed order, and the CTA folds its
// channels into the global state itself (count: bump_count_last_block).  Deterministic.
//   NSLOTS ring slots of 128 KiB / NSLOTS bytes, W = min(16, NSLOTS) consumer warps: warp w takes the CTA's
//   chunks i = w, w + W, ...; chunk i = (image i / J, class b + (i % J) * grid), J = C / PPC / grid classes per
//   CTA, W % J == 0 so a warp always sees the same class; G = 32 / PPC lanes work on one plane.
// =====================================================================================
bool bulk_ok(const void* p, int64_t hw) {
  return hw % 4 == 0 && hw * 4 <= kSlotBytesMax && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

struct AccArgs {
  const float* x;
  int N, C, hw;
  int J;        // chunk classes per CTA
  int slots;    // ring slots of one chunk (PPC planes) each, <= kMaxSlots
  int warps;    // consumer warps, <= kMaxConsumers, a multiple of J
  double* state;
};

template <int PPC>
__global__ void __launch_bounds__(32 * kMaxConsumers, 1) welford_bulk_kernel(AccArgs a) {
  const int NSLOTS = a.slots, W = a.warps;
  const int kDepth = NSLOTS / W;  // slots (chunks in flight) per warp: warp w owns slots w * kDepth ..
  const int kSlotBytes = PPC * a.hw * 4;
  constexpr int G = 32 / PPC;
  extern __shared__ __align__(128) uint8_t ring[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)NSLOTS * kSlotBytes);
  double* s_part = reinterpret_cast<double*>(bars + 2 * kMaxSlots);  // [W][PPC][3]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOTS; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bars[s])));  // full
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int grid = gridDim.x, b = blockIdx.x;
  const int cpi = a.C / PPC;                       // chunks per image
  const int64_t chunks = (int64_t)a.N * a.J;       // chunks of this CTA
  const uint32_t chunk_bytes = (uint32_t)PPC * a.hw * 4u;
  auto chunk_src = [&](int64_t i) {
    const int64_t n = i / a.J;
    const int j = b + (int)(i % a.J) * grid;
    return a.x + ((n * cpi + j) * PPC) * (int64_t)a.hw;
  };
  // self-service ring (see plane_bulk_kernel): every warp issues the bulk copies of its own chunks
  auto issue = [&](int64_t i, int slot) {
    const uint32_t full = smem_addr(&bars[slot]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(chunk_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(ring + (size_t)slot * kSlotBytes)),
        "l"(chunk_src(i)), "r"(chunk_bytes), "r"(full)
        : "memory");
  };
  const int w = warp;
  if (lane == 0)
    for (int d0 = 0; d0 < kDepth; ++d0)
      if (w + (int64_t)d0 * W < chunks) issue(w + (int64_t)d0 * W, w * kDepth + d0);
  __syncwarp();
  const double old_count = a.state[0];
  const int grp = lane / G, sub = lane % G;  // plane of the chunk, lane inside the plane's group
  const int n4 = a.hw >> 2;
  double rn = 0.0, rmean = 0.0, rm2 = 0.0;   // running state of channel (class of this warp, plane grp); lane sub == 0
  int d = 0;
  uint32_t use = 0;
  for (int64_t i = w; i < chunks; i += W) {
    const int slot = w * kDepth + d;
    bar_wait(smem_addr(&bars[slot]), use & 1);
    const float4* src = reinterpret_cast<const float4*>(ring + (size_t)slot * kSlotBytes) + (size_t)grp * n4;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = sub;
    for (; k + 3 * G < n4; k += 4 * G) {
      const float4 v0 = src[k], v1 = src[k + G], v2 = src[k + 2 * G], v3 = src[k + 3 * G];
      s0 += (v0.x + v0.y) + (v0.z + v0.w);
      s1 += (v1.x + v1.y) + (v1.z + v1.w);
      s2 += (v2.x + v2.y) + (v2.z + v2.w);
      s3 += (v3.x + v3.y) + (v3.z + v3.w);
    }
    for (; k < n4; k += G) {
      const float4 v = src[k];
      s0 += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = group_sum<G>((s0 + s1) + (s2 + s3)) / (float)a.hw;
    auto sq = [mean](const float4 v) {
      const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
      return (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    };
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
    k = sub;
    for (; k + 3 * G < n4; k += 4 * G) {
      const float4 v0 = src[k], v1 = src[k + G], v2 = src[k + 2 * G], v3 = src[k + 3 * G];
      q0 += sq(v0);
      q1 += sq(v1);
      q2 += sq(v2);
      q3 += sq(v3);
    }
    for (; k < n4; k += G) q0 += sq(src[k]);
    const float q = group_sum<G>((q0 + q1) + (q2 + q3));
    // every lane's generic-proxy reads of the slot -> ordered before the bulk copy (async proxy) that refills it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0 && i + (int64_t)kDepth * W < chunks) issue(i + (int64_t)kDepth * W, slot);
    if (++d == kDepth) d = 0, ++use;
    if (sub == 0) chan_merge(rn, rmean, rm2, (double)a.hw, (double)mean, (double)q);
  }
  // warps of the CTA -> shared memory -> fixed-order merge per channel -> global state
  if (sub == 0) {
    double* d = s_part + ((size_t)w * PPC + grp) * 3;
    d[0] = rn, d[1] = rmean, d[2] = rm2;
  }
  __syncthreads();
  if (w == 0) {
    if (lane < a.J * PPC) {
      const int jj = lane / PPC, pl = lane % PPC;
      double n = 0.0, mean = 0.0, m2 = 0.0;
      for (int ww = jj; ww < W; ww += a.J) {  // warps whose class is jj, increasing
        const double* d = s_part + ((size_t)ww * PPC + pl) * 3;
        chan_merge(n, mean, m2, d[0], d[1], d[2]);
      }
      const int c = (b + jj * grid) * PPC + pl;
      double sn = old_count, sm = a.state[1 + c], sq2 = a.state[1 + a.C + c];
      chan_merge(sn, sm, sq2, n, mean, m2);
      a.state[1 + c] = sm;
      a.state[1 + a.C + c] = sq2;
    }
    __syncwarp();
    if (lane == 0) bump_count_last_block(a.state, a.C, old_count, (double)a.hw * (double)a.N);
  }
}

// (slots, consumer warps) of a ring.  A slot must always be consumed by the SAME warp -- a parity wait may run
// at most one phase ahead of the barrier, which only program order inside one warp guarantees -- so warps | slots.
struct RingPlan {
  int slots = 0, warps = 0;
};

constexpr int kAccTail = 2 * kMaxSlots * 8 + kMaxConsumers * 4 * 3 * 8;  // barriers + [W][PPC <= 4][3] doubles

template <int PPC>
int launch_acc_cfg(const AccArgs& a, int grid, cudaStream_t st) {
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(welford_bulk_kernel<PPC>), kRingBudget + kAccTail));
  const int smem = a.slots * PPC * a.hw * 4 + kAccTail;
  welford_bulk_kernel<PPC><<<grid, 32 * a.warps, smem, st>>>(a);
  CCST_LAUNCHED();
  return CCST_OK;
}

// Picks (PPC, slots, warps, grid) for the one-launch accumulation, or returns false (-> two-launch path).
bool launch_acc(const float* x, int N, int C, int64_t hw, double* state, cudaStream_t st, int* rc) {
  if (!bulk_ok(x, hw) || C % 4 != 0) return false;
  const int sms = sm_count();
  int force_ppc = 0, force_slots = 0, force_warps = 0;
#ifdef CCST_DEV
  if (const char* e = getenv("CCST_ACC_PLAN"))  // "ppc,slots,warps" (tools/bulk_sweep.sh)
    if (sscanf(e, "%d,%d,%d", &force_ppc, &force_slots, &force_warps) != 3) force_ppc = 0;
#endif
  for (int ppc = 4; ppc >= 1; ppc >>= 1) {
    if (force_ppc && ppc != force_ppc) continue;
    const int64_t cb = (int64_t)ppc * hw * 4;
    if (C % ppc != 0 || cb > kSlotBytesMax) continue;
    // bulk copies of less than 4 KiB cost more to issue than they move (measured on [1024,512,12,12]: 2.3 KiB
    // chunks ran 3x slower than the two-launch path, whose chunks are channel-agnostic and fill a ring slot)
    if (cb < 4096) return false;
    const int cpi = C / ppc;
    int grid = 0;  // largest divisor of cpi that fits the SMs
    for (int g = (cpi < sms ? cpi : sms); g >= 1; --g)
      if (cpi % g == 0) {
        grid = g;
        break;
      }
    const int J = cpi / grid;
    // one chunk in flight per warp; measured (tools/acc_sweep.sh, profiles/r02_bulk_sweep.txt): 16 warps when 16
    // chunks fit the ring, else 12 for the J = 4 layout of 16 KiB planes, else 8 -- the kernel is insensitive to
    // the ring beyond that (its small-input floor is the 128-CTA grid, the fp64 merges and the launch)
    RingPlan rp;
    for (int wc : {16, 12, 8, 4, 2, 1})
      if ((int64_t)wc * cb <= kRingBudget && wc % J == 0 && (wc != 12 || J == 4)) {
        rp.slots = rp.warps = wc;
        break;
      }
    if (force_ppc && force_slots >= 1 && force_slots <= kMaxSlots && force_warps >= 1 && force_warps <= kMaxConsumers &&
        force_slots % force_warps == 0 && force_warps % J == 0 && (int64_t)force_slots * cb <= kRingBudget)
      rp.slots = force_slots, rp.warps = force_warps;
    // enough CTAs to pull the HBM bandwidth, classes a warp can own, one lane per channel in the final merge
    if (grid * 4 < sms * 3 || rp.warps < 1 || J * ppc > 32) continue;
    AccArgs a{x, N, C, (int)hw, J, rp.slots, rp.warps, state};
    if (ppc == 4) *rc = launch_acc_cfg<4>(a, grid, st);
    else if (ppc == 2) *rc = launch_acc_cfg<2>(a, grid, st);
    else *rc = launch_acc_cfg<1>(a, grid, st);
    return true;
  }
  return false;
}

template <int MODE, int G>
int launch_bulk_cfg(BulkArgs a, cudaStream_t st) {
  constexpr bool PROD = (MODE == 2);
  constexpr int kGroups = 32 / G;
  const int64_t plane_bytes = (int64_t)a.hw * 4;
  const int sms = sm_count();
  // measured geometries (tools/bulk_sweep.sh, profiles/r02_bulk_sweep.txt)
  if (PROD) {
    // AdaIN: 16 x 8 KiB slots / 16 warps up to 4 KiB planes, 8 x 16 KiB / 8 warps above
    const int slot_max = G == 32 ? 16384 : 8192;
    a.ppc = (int)(slot_max / plane_bytes);
    a.slots = a.warps = G == 32 ? 8 : 16;
  } else if (G == 32) {
    // one plane (<= 16 KiB) per chunk; inputs with few chunks per CTA ramp up faster on more warps
    a.ppc = 1;
    const int64_t per_cta = ceil_div64(a.planes, sms);
    const int fit = (int)(kRingBudget / plane_bytes);
    a.slots = a.warps = per_cta < 48 ? (fit < 12 ? fit : 12) : 8;
  } else {
    // chunk = the planes one warp works on at a time (32 / G), repeated up to >= 4 KiB; two chunks in
    // flight per warp when the ring has room for 32 of them
    const int64_t unit = kGroups * plane_bytes;
    a.ppc = kGroups * (int)ceil_div64(4096, unit);
    a.warps = 16;
    a.slots = 32 * a.ppc * plane_bytes <= kRingBudget ? 32 : 16;
  }
  CCST_CHECK_ARG(a.ppc >= 1 && (int64_t)a.slots * a.ppc * plane_bytes <= kRingBudget,
                 "plane_bulk: a plane of %d floats does not fit the ring", a.hw);
#ifdef CCST_DEV
  if (const char* e = getenv("CCST_BULK_PLAN")) {  // "slots,warps,ppc" (tools/bulk_sweep.sh)
    int sl = 0, w = 0, pp = 0;
    if (sscanf(e, "%d,%d,%d", &sl, &w, &pp) == 3 && sl >= 1 && sl <= kMaxSlots && w >= 1 && w <= kMaxConsumers &&
        sl % w == 0 && pp >= 1 && (int64_t)sl * pp * plane_bytes <= kRingBudget)
      a.slots = sl, a.warps = w, a.ppc = pp;
  }
#endif
  a.chunks = ceil_div64(a.planes, a.ppc);
  const int grid = (int)(a.chunks < sms ? a.chunks : sms);
  const int smem = (int)((int64_t)a.slots * a.ppc * plane_bytes) + 2 * kMaxSlots * 8;
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(plane_bulk_kernel<MODE, G, PROD>), kRingBudget + 2 * kMaxSlots * 8));
  plane_bulk_kernel<MODE, G, PROD><<<grid, 32 * (a.warps + (PROD ? 1 : 0)), smem, st>>>(a);
  CCST_LAUNCHED();
  return CCST_OK;
}

template <int MODE>
int launch_bulk(BulkArgs a, cudaStream_t st) {
  if (a.hw <= 256)  // up to 1 KiB planes (12x12 .. 16x16): 4 lanes per plane
    return launch_bulk_cfg<MODE, 4>(a, st);
  if (a.hw <= 1024)  // up to 4 KiB planes: 16 lanes per plane
    return launch_bulk_cfg<MODE, 16>(a, st);
  return launch_bulk_cfg<MODE, 32>(a, st);
}

int grid_for(int64_t groups) {
  int64_t cap = (int64_t)sm_count() * 8;  // 8 resident 256-thread CTAs per SM
  return (int)(groups < cap ? (groups > 0 ? groups : 1) : cap);
}

bool vec_ok(const void* p, int64_t hw) {
  return (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
}

template <int OUT>
int launch_stats(const StatsArgs& a, cudaStream_t st) {
  const int64_t P = a.planes, hw = a.hw;
  if (bulk_ok(a.x, hw)) {
    BulkArgs b{};
    b.x = a.x, b.planes = P, b.hw = (int)hw, b.eps = a.eps, b.unbiased = a.unbiased;
    b.mean = a.mean, b.stdv = a.stdv, b.raw = a.raw;
    return launch_bulk<OUT == OUT_MEAN_STD ? 0 : 1>(b, st);
  }
  if (vec_ok(a.x, hw) && hw <= 256) {
    stats_regs_kernel<32, 2, OUT><<<grid_for(ceil_div64(P, 8)), kThreads, 0, st>>>(a);
  } else if (vec_ok(a.x, hw) && hw <= 1024) {
    stats_regs_kernel<32, 8, OUT><<<grid_for(ceil_div64(P, 8)), kThreads, 0, st>>>(a);
  } else if (vec_ok(a.x, hw) && hw <= 4096) {
    stats_regs_kernel<256, 4, OUT><<<grid_for(P), kThreads, 0, st>>>(a);
  } else if (vec_ok(a.x, hw) && hw <= 16384) {
    stats_regs_kernel<256, 16, OUT><<<grid_for(P), kThreads, 0, st>>>(a);
  } else {
    stats_stream_kernel<OUT><<<grid_for(P), kThreads, 0, st>>>(a);
  }
  CCST_LAUNCHED();
  return CCST_OK;
}

int launch_adain(const AdainArgs& a, cudaStream_t st) {
  const int64_t P = a.planes, hw = a.hw;
  if (bulk_ok(a.x, hw) && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0) {
    BulkArgs b{};
    b.x = a.x, b.out = a.out, b.planes = P, b.hw = (int)hw, b.eps = a.eps, b.unbiased = 1;
    b.C = a.C, b.mu_s = a.mu_s, b.sigma_s = a.sigma_s, b.stat_batch_stride = a.stat_batch_stride;
    b.alpha = a.alpha;
    return launch_bulk<2>(b, st);
  }
  const bool v = vec_ok(a.x, hw) && vec_ok(a.out, hw);
  if (v && hw <= 256) {
    adain_regs_kernel<32, 2><<<grid_for(ceil_div64(P, 8)), kThreads, 0, st>>>(a);
  } else if (v && hw <= 1024) {
    adain_regs_kernel<32, 8><<<grid_for(ceil_div64(P, 8)), kThreads, 0, st>>>(a);
  } else if (v && hw <= 4096) {
    adain_regs_kernel<256, 4><<<grid_for(P), kThreads, 0, st>>>(a);
  } else if (v && hw <= 16384) {
    adain_regs_kernel<256, 16><<<grid_for(P), kThreads, 0, st>>>(a);
  } else {
    adain_stream_kernel<<<grid_for(P), kThreads, 0, st>>>(a);
  }
  CCST_LAUNCHED();
  return CCST_OK;
}

}  // namespace

int merge_raw_into_state(const float2* raw, int N, int C, int64_t hw, double* d_state,
                         cudaStream_t st) {
  const int blocks = (C + 7) / 8;
  // (the block that draws the last ticket stores the new count: see bump_count_last_block)
  merge_planes_kernel<<<blocks, 256, 0, st>>>(raw, N, C, (double)hw, d_state);
  CCST_LAUNCHED();
  return CCST_OK;
}

}  // namespace ccst

using namespace ccst;

extern "C" int ccst_stats_nchw_f32(const float* d_x, int64_t planes, int64_t hw, float eps,
                                   int unbiased, float* d_mean, float* d_std, void* stream) {
  CCST_CHECK_ARG(d_x != nullptr || planes == 0, "ccst_stats_nchw_f32: null input");
  CCST_CHECK_ARG(planes >= 0 && hw >= 1, "ccst_stats_nchw_f32: bad shape planes=%lld hw=%lld",
                 (long long)planes, (long long)hw);
  if (int e = require_sm100()) return e;
  if (planes == 0) return CCST_OK;
  StatsArgs a{d_x, planes, hw, eps, unbiased, d_mean, d_std, nullptr};
  return launch_stats<OUT_MEAN_STD>(a, (cudaStream_t)stream);
}

extern "C" int ccst_welford_accumulate_nchw_f32(const float* d_x, int N, int C, int64_t hw,
                                                double* d_state, float* d_scratch, void* stream) {
  CCST_CHECK_ARG(d_x && d_state && d_scratch, "ccst_welford_accumulate_nchw_f32: null pointer");
  CCST_CHECK_ARG(N >= 1 && C >= 1 && hw >= 1, "ccst_welford_accumulate_nchw_f32: bad shape");
  if (int e = require_sm100()) return e;
  int rc = CCST_OK;
  if (launch_acc(d_x, N, C, hw, d_state, (cudaStream_t)stream, &rc)) return rc;  // one launch, no scratch
  StatsArgs a{d_x, (int64_t)N * C, hw, 0.f, 0, nullptr, nullptr,
              reinterpret_cast<float2*>(d_scratch)};
  if (int e = launch_stats<OUT_MEAN_M2>(a, (cudaStream_t)stream)) return e;
  return merge_raw_into_state(reinterpret_cast<const float2*>(d_scratch), N, C, hw, d_state,
                              (cudaStream_t)stream);
}

extern "C" int ccst_welford_finalize(const double* d_state, int C, float eps, float* d_mean,
                                     float* d_std, void* stream) {
  CCST_CHECK_ARG(d_state && C >= 1, "ccst_welford_finalize: bad argument");
  if (int e = require_sm100()) return e;
  finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_state, C, eps, 0, d_mean, d_std);
  CCST_LAUNCHED();
  return CCST_OK;
}

extern "C" int ccst_welford_finalize_unbiased(const double* d_state, int C, float eps, float* d_mean,
                                              float* d_std, void* stream) {
  CCST_CHECK_ARG(d_state && C >= 1, "ccst_welford_finalize_unbiased: bad argument");
  if (int e = require_sm100()) return e;
  finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_state, C, eps, 1, d_mean, d_std);
  CCST_LAUNCHED();
  return CCST_OK;
}

extern "C" int ccst_welford_to_sums(const double* d_state, int C, float* d_sum, float* d_sqsum,
                                    void* stream) {
  CCST_CHECK_ARG(d_state && C >= 1, "ccst_welford_to_sums: bad argument");
  if (int e = require_sm100()) return e;
  to_sums_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_state, C, d_sum, d_sqsum);
  CCST_LAUNCHED();
  return CCST_OK;
}

extern "C" int ccst_welford_to_moments(const double* d_state, int C, double* d_moments,
                                       void* stream) {
  CCST_CHECK_ARG(d_state && d_moments && C >= 1, "ccst_welford_to_moments: bad argument");
  if (int e = require_sm100()) return e;
  to_moments_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_state, C, d_moments);
  CCST_LAUNCHED();
  return CCST_OK;
}

extern "C" int ccst_welford_from_moments(const double* d_moments, int C, double* d_state,
                                         void* stream) {
  CCST_CHECK_ARG(d_state && d_moments && C >= 1, "ccst_welford_from_moments: bad argument");
  if (int e = require_sm100()) return e;
  from_moments_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_moments, C, d_state);
  CCST_LAUNCHED();
  return CCST_OK;
}

// ncclAllReduce through a run-time binding: int ncclAllReduce(const void*, void*, size_t, int dtype, int op, void* comm,
// cudaStream_t); ncclFloat64 = 8, ncclSum = 0 (nccl.h, stable since NCCL 2.0)
namespace {
typedef int (*PFN_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*PFN_ncclGetErrorString)(int);
PFN_ncclAllReduce g_nccl_allreduce = nullptr;
PFN_ncclGetErrorString g_nccl_errstr = nullptr;
bool bind_nccl() {
  if (g_nccl_allreduce) return true;
  void* lib = nullptr;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD);  // the copy the process already uses (e.g. torch's)
    if (!lib) lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return false;
  g_nccl_allreduce = reinterpret_cast<PFN_ncclAllReduce>(dlsym(lib, "ncclAllReduce"));
  g_nccl_errstr = reinterpret_cast<PFN_ncclGetErrorString>(dlsym(lib, "ncclGetErrorString"));
  return g_nccl_allreduce != nullptr;
}
}  // namespace

extern "C" int ccst_allreduce_moments(void* nccl_comm, double* d_moments, int64_t count, void* stream) {
  CCST_CHECK_ARG(nccl_comm && d_moments && count >= 1, "ccst_allreduce_moments: bad argument");
  if (!bind_nccl()) {
    set_error("ccst_allreduce_moments: no NCCL in the process or on the loader path (libnccl.so.2)");
    return CCST_ESTATE;
  }
  const int rc = g_nccl_allreduce(d_moments, d_moments, (size_t)count, /*ncclFloat64=*/8, /*ncclSum=*/0, nccl_comm,
                                  (cudaStream_t)stream);
  if (rc != 0) {
    set_error("ccst_allreduce_moments: ncclAllReduce failed: %s", g_nccl_errstr ? g_nccl_errstr(rc) : "?");
    return CCST_ECUDA;
  }
  return CCST_OK;
}

extern "C" int ccst_adain_stat_nchw_f32(const float* d_x, int N, int C, int64_t hw,
                                        const float* d_mu_s, const float* d_sigma_s,
                                        int64_t stat_batch_stride, float alpha, float eps,
                                        float* d_out, void* stream) {
  CCST_CHECK_ARG(d_x && d_out && d_mu_s && d_sigma_s, "ccst_adain_stat_nchw_f32: null pointer");
  CCST_CHECK_ARG(N >= 1 && C >= 1 && hw >= 1, "ccst_adain_stat_nchw_f32: bad shape");
  CCST_CHECK_ARG(stat_batch_stride == 0 || stat_batch_stride == C,
                 "ccst_adain_stat_nchw_f32: stat_batch_stride must be 0 or C");
  CCST_CHECK_ARG(alpha >= 0.f && alpha <= 1.f, "ccst_adain_stat_nchw_f32: alpha outside [0,1]");
  if (int e = require_sm100()) return e;
  AdainArgs a{d_x, d_out, (int64_t)N * C, hw, C, d_mu_s, d_sigma_s, stat_batch_stride, alpha, eps};
  return launch_adain(a, (cudaStream_t)stream);
}

extern "C" int ccst_adain_feat_nchw_f32(const float* d_content, const float* d_style, int N, int C,
                                        int64_t hw_c, int64_t hw_s, float alpha, float eps,
                                        float* d_out, float* d_scratch, void* stream) {
  CCST_CHECK_ARG(d_content && d_style && d_out && d_scratch,
                 "ccst_adain_feat_nchw_f32: null pointer");
  CCST_CHECK_ARG(N >= 1 && C >= 1 && hw_c >= 1 && hw_s >= 1, "ccst_adain_feat_nchw_f32: bad shape");
  CCST_CHECK_ARG(alpha >= 0.f && alpha <= 1.f, "ccst_adain_feat_nchw_f32: alpha outside [0,1]");
  if (int e = require_sm100()) return e;
  float* mu = d_scratch;
  float* sg = d_scratch + (size_t)N * C;
  StatsArgs s{d_style, (int64_t)N * C, hw_s, eps, 1, mu, sg, nullptr};
  if (int e = launch_stats<OUT_MEAN_STD>(s, (cudaStream_t)stream)) return e;
  AdainArgs a{d_content, d_out, (int64_t)N * C, hw_c, C, mu, sg, (int64_t)C, alpha, eps};
  return launch_adain(a, (cudaStream_t)stream);
}
