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
// state = {count, mean[C], M2[C]}
__global__ void merge_planes_kernel(const float2* __restrict__ raw, int N, int C, double hw,
                                    double* __restrict__ state) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const double n0 = state[0];
  if (c < C) {
    double n = n0, mean = state[1 + c], m2 = state[1 + C + c];
    for (int i = 0; i < N; ++i) {
      float2 r = raw[(size_t)i * C + c];
      double nb = hw, d = (double)r.x - mean, nn = n + nb;
      mean += d * (nb / nn);
      m2 += (double)r.y + d * d * (n * nb / nn);
      n = nn;
    }
    state[1 + c] = mean;
    state[1 + C + c] = m2;
  }
  // every thread has read state[0] before anyone may overwrite it
  __syncthreads();
  if (gridDim.x == 1 && threadIdx.x == 0) state[0] = n0 + hw * N;
}
__global__ void bump_count_kernel(double* state, double add) { state[0] += add; }

__global__ void finalize_kernel(const double* __restrict__ state, int C, float eps,
                                float* __restrict__ mean, float* __restrict__ stdv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = state[0];
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
  const int threads = 128;
  const int blocks = (C + threads - 1) / threads;
  merge_planes_kernel<<<blocks, threads, 0, st>>>(raw, N, C, (double)hw, d_state);
  CCST_LAUNCHED();
  if (blocks > 1) {
    bump_count_kernel<<<1, 1, 0, st>>>(d_state, (double)hw * N);
    CCST_LAUNCHED();
  }
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
  finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_state, C, eps, d_mean, d_std);
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
