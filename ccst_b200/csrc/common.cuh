// Shared device/host helpers for libccst_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/ccst_b200.h"

namespace ccst {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
extern int64_t g_launches;

#define CCST_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      ccst::set_error(__VA_ARGS__);    \
      return CCST_EINVAL;              \
    }                                  \
  } while (0)

#define CCST_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ccst::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                \
                      cudaGetErrorString(_e));                                            \
      return CCST_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

// counts the launch and surfaces launch-configuration errors immediately
#define CCST_LAUNCHED()                                                                   \
  do {                                                                                    \
    ++ccst::g_launches;                                                                   \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ccst::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,            \
                      cudaGetErrorString(_e));                                            \
      return CCST_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

int require_sm100();  // 0 or CCST_EARCH for the current device
int sm_count();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is
// per device, and a process may own handles on several GPUs.
cudaError_t ensure_dyn_smem(const void* kernel, int bytes);

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- Welford
// (count, mean, M2) triple; `add` is the classic one-sample update, `merge` is
// Chan et al.'s pairwise combination.  All fp32 unless the double variant.
struct Wf {
  float n, mean, m2;
};

__device__ __forceinline__ Wf wf_merge(Wf a, Wf b) {
  float n = a.n + b.n;
  if (n == 0.f) return a;
  float d = b.mean - a.mean;
  float f = __fdividef(b.n, n);
  Wf r;
  r.n = n;
  r.mean = fmaf(d, f, a.mean);
  r.m2 = a.m2 + b.m2 + d * d * a.n * f;
  return r;
}

__device__ __forceinline__ Wf wf_warp_reduce(Wf v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Wf o;
    o.n = __shfl_xor_sync(0xffffffffu, v.n, off);
    o.mean = __shfl_xor_sync(0xffffffffu, v.mean, off);
    o.m2 = __shfl_xor_sync(0xffffffffu, v.m2, off);
    v = wf_merge(v, o);
  }
  return v;
}

// ---------------------------------------------------------------- activations layout
// Internal activations are NHWC with a one-pixel reflection halo:
//   buffer[n][y+1][x+1][c],  y in [-1,H], x in [-1,W],  halo(-1) = interior(1), halo(H) = interior(H-2)
// so a 3x3 reflect-pad convolution reads plain shifted windows (TMA boxes) with no border logic.
template <typename T>
struct ActView {
  T* p;
  int N, H, W, C;  // interior size
  __host__ __device__ size_t pitch_x() const { return (size_t)C; }
  __host__ __device__ size_t pitch_y() const { return (size_t)(W + 2) * C; }
  __host__ __device__ size_t pitch_n() const { return (size_t)(H + 2) * (W + 2) * C; }
  __host__ __device__ size_t elems() const { return (size_t)N * pitch_n(); }
  // pointer to pixel (n, y, x) with y,x in [-1, H] / [-1, W]
  __device__ __forceinline__ T* px(int n, int y, int x) const {
    return p + (size_t)n * pitch_n() + (size_t)(y + 1) * pitch_y() + (size_t)(x + 1) * C;
  }
};

// Calls f(yy, xx) for (y, x) itself and for every halo position that mirrors it.
// edge = 1: reflection halo (halo(-1) = interior(1), the layout every 3x3 reflect-pad conv reads);
// edge = 0: replicate halo (halo(-1) = interior(0)), written by the layers whose consumer is the
// fused nearest-x2-upsample conv: reflect(upsample(S))[-1] = upsample(S)[1] = S[0].
template <typename F>
__device__ __forceinline__ void for_each_halo_alias(int y, int x, int H, int W, F&& f, int edge = 1) {
  int ys[3], xs[3];
  int ny = 0, nx = 0;
  ys[ny++] = y;
  if (y == edge) ys[ny++] = -1;
  if (y == H - 1 - edge) ys[ny++] = H;
  xs[nx++] = x;
  if (x == edge) xs[nx++] = -1;
  if (x == W - 1 - edge) xs[nx++] = W;
  for (int i = 0; i < ny; ++i)
    for (int j = 0; j < nx; ++j) f(ys[i], xs[j]);
}

// torchvision.utils.save_image: grid.mul(255).add_(0.5).clamp_(0, 255).to(uint8) -- two separately
// rounded fp32 operations (no FMA contraction), then truncation.
__device__ __forceinline__ uint8_t quantize_u8(float v) {
  const float t = __fadd_rn(__fmul_rn(v, 255.f), 0.5f);
  return (uint8_t)fminf(fmaxf(t, 0.f), 255.f);
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// fp16 activations saturate instead of overflowing to inf (fp32 accumulators can exceed 65504)
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) {
  return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
}

// Two fp32 -> one packed pair in ONE instruction (F2FP): round-to-nearest; fp16 saturates to
// +-65504 instead of overflowing to inf (fp32 accumulators can exceed the fp16 range); the `_relu`
// forms clamp negatives to zero in the same instruction (ReLU fused into the store conversion).
template <typename T>
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack16x2<__nv_bfloat16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <>
__device__ __forceinline__ uint32_t pack16x2<__half>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// elementwise max of two packed pairs (exact: rounding is monotonic, so pooling may run on the
// already-converted values and needs half as many shuffles as pooling in fp32)
template <typename T>
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b);
template <>
__device__ __forceinline__ uint32_t max16x2<__half>(uint32_t a, uint32_t b) {
  const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <>
__device__ __forceinline__ uint32_t max16x2<__nv_bfloat16>(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r =
      __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <typename T>
__device__ __forceinline__ uint32_t pack16x2_relu(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack16x2_relu<__nv_bfloat16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <>
__device__ __forceinline__ uint32_t pack16x2_relu<__half>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// packed pair -> two fp32
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

}  // namespace ccst
