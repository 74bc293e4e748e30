// Bandwidth-side layers of the encoder/decoder and the fp32 validation convolution.
//
//  * conv_first      : conv1_1 with the 1x1 colour conv folded in (net.py:39-42), NCHW fp32 image in,
//                      NHWC activation out (HBM-bound: K = 27)
//  * conv_ffma       : fp32 implicit-GEMM 3x3 reflect-pad conv on CUDA cores -- the "fp32 mode"
//                      whose output matches the reference within 1e-4 (net.py:6-69)
//  * pool            : MaxPool2d(2,2,ceil_mode=True) (net.py:46,53,66)
//  * adain_nhwc      : calc_mean_std + AdaIN + alpha blend on the arena layout
//                      (function.py:26-33, CCST_OverallStyleTransfer.py:44-45)
//  * stats_nhwc      : per-(n,c) Welford partials of relu4_1 for the overall-style loop
//                      (mean_std_computation_effcientMem.py:103-131)
//  * layout converters between the reference's NCHW fp32 tensors and the arena layout.
#include <type_traits>

#include "layers.h"

namespace ccst {

namespace {

__host__ __device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

template <typename T, int VEC>
struct Pack;  // VEC values of T moved as one vector
template <>
struct Pack<float, 4> {
  float4 v;
  __device__ __forceinline__ void set(int i, float f) { (&v.x)[i] = f; }
  __device__ __forceinline__ float get(int i) const { return (&v.x)[i]; }
};
template <>
struct Pack<__nv_bfloat16, 8> {
  uint4 v;
  __device__ __forceinline__ void set(int i, float f) {
    reinterpret_cast<__nv_bfloat16*>(&v)[i] = __float2bfloat16_rn(f);
  }
  __device__ __forceinline__ float get(int i) const {
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&v)[i]);
  }
};
template <>
struct Pack<__half, 8> {
  uint4 v;
  __device__ __forceinline__ void set(int i, float f) {
    reinterpret_cast<__half*>(&v)[i] = from_f32<__half>(f);
  }
  __device__ __forceinline__ float get(int i) const {
    return __half2float(reinterpret_cast<const __half*>(&v)[i]);
  }
};
template <typename T>
struct VecOf {
  static constexpr int value = 16 / sizeof(T);
};

// whole-vector conversions (two elements per instruction for the 16-bit types)
__device__ __forceinline__ void unpack_vec(const Pack<float, 4>& p, float (&f)[4]) {
  f[0] = p.v.x, f[1] = p.v.y, f[2] = p.v.z, f[3] = p.v.w;
}
__device__ __forceinline__ void unpack_vec(const Pack<__half, 8>& p, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&p.v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void unpack_vec(const Pack<__nv_bfloat16, 8>& p, float (&f)[8]) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&p.v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void pack_vec(Pack<float, 4>& p, const float (&f)[4]) {
  p.v = make_float4(f[0], f[1], f[2], f[3]);
}
__device__ __forceinline__ void pack_vec(Pack<__half, 8>& p, const float (&f)[8]) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&p.v);
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = pack16x2<__half>(f[2 * i], f[2 * i + 1]);
}
__device__ __forceinline__ void pack_vec(Pack<__nv_bfloat16, 8>& p, const float (&f)[8]) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&p.v);
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = pack16x2<__nv_bfloat16>(f[2 * i], f[2 * i + 1]);
}

// =====================================================================================
// conv1_1 (+ folded 1x1): one CTA = 64 consecutive pixels of one image row, 4 threads per pixel
// (16 output channels each).
// =====================================================================================
constexpr int kFirstTile = 64;

template <typename T>
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ img, int N, int H,
                                                         int W, const float* __restrict__ w27,
                                                         const float* __restrict__ bias64,
                                                         ActView<T> out) {
  __shared__ __align__(16) float sw[27 * 64];
  __shared__ float sb[64];
  __shared__ float sin[3][3][kFirstTile + 2];  // [ci][row][col]
  const int tiles_x = (W + kFirstTile - 1) / kFirstTile;
  int b = blockIdx.x;
  const int tx = b % tiles_x;
  b /= tiles_x;
  const int y = b % H;
  const int n = b / H;
  const int x0 = tx * kFirstTile;

  for (int i = threadIdx.x; i < 27 * 64; i += 256) sw[i] = w27[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = bias64[threadIdx.x];
  for (int i = threadIdx.x; i < 3 * 3 * (kFirstTile + 2); i += 256) {
    int col = i % (kFirstTile + 2);
    int row = (i / (kFirstTile + 2)) % 3;
    int ci = i / (3 * (kFirstTile + 2));
    int yy = reflect_idx(y + row - 1, H);
    int xx = x0 + col - 1;
    float v = 0.f;
    if (xx <= W) {  // xx == W is the right halo of the last pixel
      xx = reflect_idx(xx, W);
      v = __ldg(img + (((size_t)n * 3 + ci) * H + yy) * W + xx);
    }
    sin[ci][row][col] = v;
  }
  __syncthreads();

  const int p = threadIdx.x >> 2;        // pixel in tile
  const int cg = (threadIdx.x & 3) * 16;  // first output channel
  const int x = x0 + p;
  if (x >= W) return;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = sb[cg + j];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float v = sin[ci][r][p + s];
        const float4* w4 = reinterpret_cast<const float4*>(&sw[((r * 3 + s) * 3 + ci) * 64 + cg]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 w = w4[q];
          acc[q * 4 + 0] = fmaf(v, w.x, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(v, w.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(v, w.z, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(v, w.w, acc[q * 4 + 3]);
        }
      }
  constexpr int VEC = VecOf<T>::value;
  Pack<T, VEC> pk[16 / VEC];
#pragma unroll
  for (int j = 0; j < 16; ++j) pk[j / VEC].set(j % VEC, fmaxf(acc[j], 0.f));
  for_each_halo_alias(y, x, H, W, [&](int yy, int xx) {
    T* dst = out.px(n, yy, xx) + cg;
#pragma unroll
    for (int j = 0; j < 16 / VEC; ++j) reinterpret_cast<decltype(pk[0].v)*>(dst)[j] = pk[j].v;
  });
}

// The same conv1_1 with 4 consecutive pixels per thread (CTA = 256 pixels of one image row x 64 channels,
// thread = 4 pixels x 16 channels): every weight vector read from shared memory feeds 4 pixels (64 FMAs per
// 4 LDS.128 instead of 16), which is what the one-pixel form above is bound by.  Used by the f16x3 engine,
// whose first layer stays on the CUDA cores in fp32 and is stored as [hi | lo] f16 halves.
constexpr int kFirst4Tile = 256;

template <typename T16, bool SPLIT>
__global__ void __launch_bounds__(256)
    conv_first_px4_kernel(const float* __restrict__ img, int N, int H, int W, const float* __restrict__ w27,
                          const float* __restrict__ bias64, ActView<T16> out, unsigned int* sat_count) {
  __shared__ __align__(16) float sw[27 * 64];
  __shared__ float sb[64];
  __shared__ float sin[3][3][kFirst4Tile + 4];  // [ci][row][col]
  const int tiles_x = (W + kFirst4Tile - 1) / kFirst4Tile;
  int b = blockIdx.x;
  const int tx = b % tiles_x;
  b /= tiles_x;
  const int y = b % H;
  const int n = b / H;
  const int x0 = tx * kFirst4Tile;

  for (int i = threadIdx.x; i < 27 * 64; i += 256) sw[i] = w27[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = bias64[threadIdx.x];
  for (int i = threadIdx.x; i < 3 * 3 * (kFirst4Tile + 2); i += 256) {
    const int col = i % (kFirst4Tile + 2);
    const int row = (i / (kFirst4Tile + 2)) % 3;
    const int ci = i / (3 * (kFirst4Tile + 2));
    const int yy = reflect_idx(y + row - 1, H);
    int xx = x0 + col - 1;
    float v = 0.f;
    if (xx <= W) {  // xx == W is the right halo of the last pixel
      xx = reflect_idx(xx, W);
      v = __ldg(img + (((size_t)n * 3 + ci) * H + yy) * W + xx);
    }
    sin[ci][row][col] = v;
  }
  __syncthreads();

  const int pg = threadIdx.x >> 2;        // group of 4 pixels
  const int cg = (threadIdx.x & 3) * 16;  // first output channel
  const int xb = x0 + 4 * pg;
  if (xb >= W) return;
  float acc[4][16];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[q][j] = sb[cg + j];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      float v[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = sin[ci][r][4 * pg + c];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const float4* w4 = reinterpret_cast<const float4*>(&sw[((r * 3 + s) * 3 + ci) * 64 + cg]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 w = w4[k];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[q][k * 4 + 0] = fmaf(v[q + s], w.x, acc[q][k * 4 + 0]);
            acc[q][k * 4 + 1] = fmaf(v[q + s], w.y, acc[q][k * 4 + 1]);
            acc[q][k * 4 + 2] = fmaf(v[q + s], w.z, acc[q][k * 4 + 2]);
            acc[q][k * 4 + 3] = fmaf(v[q + s], w.w, acc[q][k * 4 + 3]);
          }
        }
      }
    }
  uint32_t hmax = 0u;  // running maximum of the stored (non-negative) high parts: f16 saturation guard
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int x = xb + q;
    if (x >= W) break;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v0 = fmaxf(acc[q][2 * j], 0.f), v1 = fmaxf(acc[q][2 * j + 1], 0.f);
      hi[j] = pack16x2<T16>(v0, v1);
      const float2 hf = unpack16x2<T16>(hi[j]);
      lo[j] = pack16x2<T16>(v0 - hf.x, v1 - hf.y);
      hmax = max16x2<T16>(hmax, hi[j]);
    }
    for_each_halo_alias(y, x, H, W, [&](int yy, int xx) {
      uint4* dst = reinterpret_cast<uint4*>(out.px(n, yy, xx) + cg);
      dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      dst[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      if (SPLIT) {
        uint4* dl = reinterpret_cast<uint4*>(out.px(n, yy, xx) + 64 + cg);
        dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
    });
  }
  if (std::is_same<T16, __half>::value && ((hmax & 0xffffu) >= 0x7bffu || (hmax >> 16) >= 0x7bffu) &&
      sat_count != nullptr)
    atomicAdd(sat_count, 1u);
}

// =====================================================================================
// fp32 FFMA implicit GEMM: CTA tile = 8x16 pixels x 64 output channels, K chunk = 16 channels
// (all 9 taps per chunk).  Thread = 8 consecutive pixels of one tile row x 4 output channels.
// =====================================================================================
constexpr int kFT_H = 8, kFT_W = 16, kFT_N = 64, kFT_K = 16;

template <int EPI>
__global__ void __launch_bounds__(256)
    conv_ffma_kernel(ActView<float> in, const float* __restrict__ w, const float* __restrict__ bias,
                     int Cout, int CoutPad, int relu, ActView<float> out,
                     float* __restrict__ out_nchw) {
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;                                          // [10][18][16]
  float* s_w = smem + (kFT_H + 2) * (kFT_W + 2) * kFT_K;       // [9][16][64]
  const int H = in.H, W = in.W, Cin = in.C;
  const int tiles_x = (W + kFT_W - 1) / kFT_W, tiles_y = (H + kFT_H - 1) / kFT_H;
  int b = blockIdx.x;
  const int ntile = b % (CoutPad / kFT_N);
  b /= (CoutPad / kFT_N);
  const int tx = b % tiles_x;
  b /= tiles_x;
  const int ty = b % tiles_y;
  const int n = b / tiles_y;
  const int y0 = ty * kFT_H, x0 = tx * kFT_W, co0 = ntile * kFT_N;

  const int t = threadIdx.x;
  const int cq = (t & 15) * 4;      // output channel quad inside the tile
  const int pg = t >> 4;            // pixel group 0..15
  const int prow = pg >> 1;         // tile row
  const int pcol = (pg & 1) * 8;    // first tile column
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += kFT_K) {
    __syncthreads();
    // input halo tile: (10 x 18) pixels x 16 channels
    for (int i = t; i < (kFT_H + 2) * (kFT_W + 2) * (kFT_K / 4); i += 256) {
      const int c4 = i % (kFT_K / 4);
      const int pix = i / (kFT_K / 4);
      const int col = pix % (kFT_W + 2), row = pix / (kFT_W + 2);
      const int yy = y0 - 1 + row, xx = x0 - 1 + col;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yy <= H && xx <= W)
        v = *reinterpret_cast<const float4*>(in.px(n, yy, xx) + c0 + c4 * 4);
      *reinterpret_cast<float4*>(s_in + (size_t)pix * kFT_K + c4 * 4) = v;
    }
    // weights: [tap][ci][64]
    for (int i = t; i < 9 * kFT_K * (kFT_N / 4); i += 256) {
      const int q = i % (kFT_N / 4);
      const int ci = (i / (kFT_N / 4)) % kFT_K;
      const int tap = i / (kFT_N / 4 * kFT_K);
      *reinterpret_cast<float4*>(s_w + ((size_t)tap * kFT_K + ci) * kFT_N + q * 4) =
          *reinterpret_cast<const float4*>(w + ((size_t)tap * Cin + c0 + ci) * CoutPad + co0 + q * 4);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const float* ip = s_in + ((size_t)(prow + r) * (kFT_W + 2) + pcol + s) * kFT_K;
        const float* wp = s_w + (size_t)(r * 3 + s) * kFT_K * kFT_N + cq;
#pragma unroll
        for (int c4 = 0; c4 < kFT_K / 4; ++c4) {
          float4 wv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            wv[q] = *reinterpret_cast<const float4*>(wp + (size_t)(c4 * 4 + q) * kFT_N);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 iv = *reinterpret_cast<const float4*>(ip + (size_t)j * kFT_K + c4 * 4);
            acc[j][0] = fmaf(iv.x, wv[0].x, acc[j][0]);
            acc[j][1] = fmaf(iv.x, wv[0].y, acc[j][1]);
            acc[j][2] = fmaf(iv.x, wv[0].z, acc[j][2]);
            acc[j][3] = fmaf(iv.x, wv[0].w, acc[j][3]);
            acc[j][0] = fmaf(iv.y, wv[1].x, acc[j][0]);
            acc[j][1] = fmaf(iv.y, wv[1].y, acc[j][1]);
            acc[j][2] = fmaf(iv.y, wv[1].z, acc[j][2]);
            acc[j][3] = fmaf(iv.y, wv[1].w, acc[j][3]);
            acc[j][0] = fmaf(iv.z, wv[2].x, acc[j][0]);
            acc[j][1] = fmaf(iv.z, wv[2].y, acc[j][1]);
            acc[j][2] = fmaf(iv.z, wv[2].z, acc[j][2]);
            acc[j][3] = fmaf(iv.z, wv[2].w, acc[j][3]);
            acc[j][0] = fmaf(iv.w, wv[3].x, acc[j][0]);
            acc[j][1] = fmaf(iv.w, wv[3].y, acc[j][1]);
            acc[j][2] = fmaf(iv.w, wv[3].z, acc[j][2]);
            acc[j][3] = fmaf(iv.w, wv[3].w, acc[j][3]);
          }
        }
      }
  }

  const float4 bv = *reinterpret_cast<const float4*>(bias + co0 + cq);
  const int y = y0 + prow;
  if (y >= H) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int x = x0 + pcol + j;
    if (x >= W) continue;
    float4 o = make_float4(acc[j][0] + bv.x, acc[j][1] + bv.y, acc[j][2] + bv.z, acc[j][3] + bv.w);
    if (relu) {
      o.x = fmaxf(o.x, 0.f);
      o.y = fmaxf(o.y, 0.f);
      o.z = fmaxf(o.z, 0.f);
      o.w = fmaxf(o.w, 0.f);
    }
    const int co = co0 + cq;
    if (EPI == EPI_ACT) {
      if (co < Cout)
        for_each_halo_alias(y, x, H, W, [&](int yy, int xx) {
          *reinterpret_cast<float4*>(out.px(n, yy, xx) + co) = o;
        });
    } else if (EPI == EPI_ACT_UP2) {
      if (co < Cout)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int bb = 0; bb < 2; ++bb)
            for_each_halo_alias(2 * y + a, 2 * x + bb, 2 * H, 2 * W, [&](int yy, int xx) {
              *reinterpret_cast<float4*>(out.px(n, yy, xx) + co) = o;
            });
    } else {  // EPI_NCHW_F32
      const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (co + q < Cout) out_nchw[(((size_t)n * Cout + co + q) * H + y) * W + x] = ov[q];
    }
  }
}

// =====================================================================================
// 2x2 stride-2 ceil-mode max pool, one thread per (pixel, 16-byte channel vector)
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(256) pool_kernel(ActView<T> in, ActView<T> out) {
  constexpr int VEC = VecOf<T>::value;
  const int cv = in.C / VEC;
  const size_t total = (size_t)out.N * out.H * out.W * cv;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int c = (int)(i % cv) * VEC;
    size_t p = i / cv;
    const int xo = (int)(p % out.W);
    p /= out.W;
    const int yo = (int)(p % out.H);
    const int n = (int)(p / out.H);
    float m[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) m[k] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = 2 * yo + dy, x = 2 * xo + dx;
        if (y < in.H && x < in.W) {
          Pack<T, VEC> v;
          v.v = *reinterpret_cast<const decltype(v.v)*>(in.px(n, y, x) + c);
#pragma unroll
          for (int k = 0; k < VEC; ++k) m[k] = fmaxf(m[k], v.get(k));
        }
      }
    Pack<T, VEC> o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) o.set(k, m[k]);
    for_each_halo_alias(yo, xo, out.H, out.W, [&](int yy, int xx) {
      *reinterpret_cast<decltype(o.v)*>(out.px(n, yy, xx) + c) = o.v;
    });
  }
}

// =====================================================================================
// NHWC statistics / AdaIN: CTA = (image n, 64-channel slab).  Thread = VEC channels x a strided
// subset of the pixels, streaming Welford with a uniform count, merged through shared memory.
// =====================================================================================
template <typename T>
struct NhwcGeom {
  static constexpr int VEC = VecOf<T>::value;
  static constexpr int CH_LANES = 64 / VEC;
  static constexpr int PX_LANES = 256 / CH_LANES;
};

// Statistics / AdaIN on the arena layout.  CTA = (image n, chunk of kNhwcChunk pixels) x ALL channels:
// a warp reads 512 contiguous bytes of one pixel (the earlier one-CTA-per-64-channel-slab mapping
// read 128-byte pieces at a 1 KiB stride from eight different SMs and reached 2.3 TB/s).  Every
// chunk CTA writes its {mean, M2} partials (the count follows from the geometry); consumers merge
// the partials of a plane with Chan's formula in fixed chunk order (deterministic).
constexpr int kNhwcChunk = 128;
constexpr int kNhwcBatch = 8;  // 16-byte loads in flight per thread

__host__ __device__ inline int nhwc_chunks(int HW) { return (HW + kNhwcChunk - 1) / kNhwcChunk; }

// grid (chunks, N): part[(n * chunks + chunk) * C + c] = {mean, M2} of the chunk.
// Thread = VEC channels x every PX_LANES-th pixel, as pivoted sums (d = x - x_first: the
// cancellation in  M2 = sum d^2 - (sum d)^2 / n  scales with the spread of the values, not with
// their mean; 3 instructions per element instead of a Welford update); pixel lanes are merged
// through shared memory (Chan).
// SPLIT (f16x3 engine): the map holds [hi | lo] halves of C = in.C / 2 logical channels; value = hi + lo.
template <typename T, bool SPLIT = false>
__global__ void __launch_bounds__(256)
    nhwc_stats_partial_kernel(ActView<T> in, float2* __restrict__ part) {
  constexpr int VEC = VecOf<T>::value;
  extern __shared__ float s_part[];  // [PX_LANES][C] mean, [PX_LANES][C] M2, [PX_LANES] count
  const int C = SPLIT ? in.C / 2 : in.C, ch_lanes = C / VEC, px_lanes = 256 / ch_lanes;
  float* s_mean = s_part;
  float* s_m2 = s_part + px_lanes * C;
  float* s_n = s_part + 2 * px_lanes * C;
  const int chunk = blockIdx.x, n = blockIdx.y;
  const int HW = in.H * in.W;
  const int p0 = chunk * kNhwcChunk, p1 = min(HW, p0 + kNhwcChunk);
  const int cl = threadIdx.x % ch_lanes, pl = threadIdx.x / ch_lanes;
  const T* src0 = in.px(n, 0, 0) + cl * VEC;
  float piv[VEC], s1[VEC], s2[VEC];
  float cnt = 0.f;
#pragma unroll
  for (int k = 0; k < VEC; ++k) piv[k] = 0.f, s1[k] = 0.f, s2[k] = 0.f;
  for (int pb = p0 + pl; pb < p1; pb += kNhwcBatch * px_lanes) {
    Pack<T, VEC> v[kNhwcBatch], v2[SPLIT ? kNhwcBatch : 1];
#pragma unroll
    for (int i = 0; i < kNhwcBatch; ++i) {  // all loads of the batch first
      const int p = pb + i * px_lanes;
      if (p < p1) {
        const int y = p / in.W, x = p - y * in.W;
        const T* src = src0 + (size_t)y * in.pitch_y() + (size_t)x * in.C;
        v[i].v = *reinterpret_cast<const decltype(v[i].v)*>(src);
        if (SPLIT) v2[SPLIT ? i : 0].v = *reinterpret_cast<const decltype(v[i].v)*>(src + C);
      }
    }
#pragma unroll
    for (int i = 0; i < kNhwcBatch; ++i) {
      const int p = pb + i * px_lanes;
      if (p < p1) {
        float f[VEC];
        unpack_vec(v[i], f);
        if (SPLIT) {
          float f2[VEC];
          unpack_vec(v2[SPLIT ? i : 0], f2);
#pragma unroll
          for (int k = 0; k < VEC; ++k) f[k] += f2[k];
        }
        if (cnt == 0.f) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) piv[k] = f[k];
        } else {
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float d = f[k] - piv[k];
            s1[k] += d;
            s2[k] = fmaf(d, d, s2[k]);
          }
        }
        cnt += 1.f;
      }
    }
  }
  const float inv = cnt > 0.f ? __frcp_rn(cnt) : 0.f;
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    s_mean[pl * C + cl * VEC + k] = fmaf(s1[k], inv, piv[k]);
    s_m2[pl * C + cl * VEC + k] = fmaxf(s2[k] - s1[k] * s1[k] * inv, 0.f);
  }
  if (cl == 0) s_n[pl] = cnt;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    Wf acc{s_n[0], s_mean[c], s_m2[c]};
    for (int l = 1; l < px_lanes; ++l) acc = wf_merge(acc, Wf{s_n[l], s_mean[l * C + c], s_m2[l * C + c]});
    part[((size_t)n * gridDim.x + chunk) * C + c] = make_float2(acc.mean, acc.m2);
  }
}

// Chan merge of the chunk partials of plane (n, c), fixed order
__device__ __forceinline__ Wf nhwc_merge_partials(const float2* __restrict__ part, int n, int c, int C,
                                                  int HW) {
  const int chunks = nhwc_chunks(HW);
  Wf acc{0.f, 0.f, 0.f};
  for (int ch = 0; ch < chunks; ++ch) {
    const float2 o = part[((size_t)n * chunks + ch) * C + c];
    const int cnt = min(kNhwcChunk, HW - ch * kNhwcChunk);
    acc = wf_merge(acc, Wf{(float)cnt, o.x, o.y});
  }
  return acc;
}

// raw[n * C + c] = {mean, M2} of the whole plane (input of merge_raw_into_state)
__global__ void __launch_bounds__(256)
    nhwc_stats_merge_kernel(const float2* __restrict__ part, float2* __restrict__ raw, int NC, int C,
                            int HW) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= NC) return;
  const Wf acc = nhwc_merge_partials(part, i / C, i % C, C, HW);
  raw[i] = make_float2(acc.mean, acc.m2);
}

// Chan merge, in fixed order, of the statistics the conv epilogue wrote per (8x16 pixel tile, quarter
// of the tile = two tile rows): tile_part[((n * tiles_y + ty) * tiles_x + tx) * 4 + q][c].
// Block = (image n, 32 channels) x 8 tile groups: group g merges tiles g, g + 8, ... (independent
// loads, 8x shorter dependency chains), then lanes of group 0 merge the 8 group results in order.
__device__ __forceinline__ Wf nhwc_merge_tiles_block(const float2* __restrict__ tp, int n, int c, int C, int H,
                                                     int W, Wf* s_grp /* [8][32] */) {
  const int tiles_x = (W + 15) / 16, tiles_y = (H + 7) / 8, tiles = tiles_x * tiles_y;
  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Wf acc{0.f, 0.f, 0.f};
  for (int t = g; t < tiles; t += 8) {
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int wv = min(16, W - tx * 16);
    const size_t base = ((size_t)n * tiles + t) * 4;
    float2 o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = tp[(base + q) * C + c];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int rows = max(0, min(2, H - (ty * 8 + 2 * q)));
      if (rows > 0) acc = wf_merge(acc, Wf{(float)(rows * wv), o[q].x, o[q].y});
    }
  }
  s_grp[g * 32 + lane] = acc;
  __syncthreads();
  Wf r = s_grp[lane];
  if (g == 0) {
#pragma unroll
    for (int k = 1; k < 8; ++k) r = wf_merge(r, s_grp[k * 32 + lane]);
  }
  return r;  // valid in group 0
}

// grid = N * C / 32 blocks of 256 threads
__global__ void __launch_bounds__(256)
    nhwc_tiles_merge_kernel(const float2* __restrict__ tp, float2* __restrict__ raw, int C, int H, int W) {
  __shared__ Wf s_grp[8 * 32];
  const int blocks_per_n = C / 32;
  const int n = blockIdx.x / blocks_per_n, c = (blockIdx.x % blocks_per_n) * 32 + (threadIdx.x & 31);
  const Wf acc = nhwc_merge_tiles_block(tp, n, c, C, H, W, s_grp);
  if (threadIdx.x < 32) raw[(size_t)n * C + c] = make_float2(acc.mean, acc.m2);
}

__global__ void __launch_bounds__(256)
    adain_nhwc_coef_tiles_kernel(const float2* __restrict__ tp, float4* __restrict__ coef, int C, int H, int W,
                                 const float* __restrict__ mu_s, const float* __restrict__ sigma_s,
                                 int64_t stat_batch_stride, float alpha, float eps) {
  __shared__ Wf s_grp[8 * 32];
  const int blocks_per_n = C / 32;
  const int n = blockIdx.x / blocks_per_n, c = (blockIdx.x % blocks_per_n) * 32 + (threadIdx.x & 31);
  const Wf st = nhwc_merge_tiles_block(tp, n, c, C, H, W, s_grp);
  if (threadIdx.x >= 32) return;
  const float sg_c = sqrtf(st.m2 / ((float)(H * W) - 1.f) + eps);
  const int64_t si = (int64_t)n * stat_batch_stride + c;
  const float ms = mu_s[si], ss = sigma_s[si];
  coef[(size_t)n * C + c] = make_float4(st.mean, alpha * (ss / sg_c) + (1.f - alpha),
                                        alpha * ms + (1.f - alpha) * st.mean, 0.f);
}

// The same statistics -> {A, B - mu_c * A} as two float arrays (folded AdaIN: out = x * A + B').
__global__ void __launch_bounds__(256)
    adain_fold_coef_tiles_kernel(const float2* __restrict__ tp, float* __restrict__ fold_a, float* __restrict__ fold_b,
                                 int C, int H, int W, const float* __restrict__ mu_s,
                                 const float* __restrict__ sigma_s, int64_t stat_batch_stride, float alpha, float eps) {
  __shared__ Wf s_grp[8 * 32];
  const int blocks_per_n = C / 32;
  const int n = blockIdx.x / blocks_per_n, c = (blockIdx.x % blocks_per_n) * 32 + (threadIdx.x & 31);
  const Wf st = nhwc_merge_tiles_block(tp, n, c, C, H, W, s_grp);
  if (threadIdx.x >= 32) return;
  const float sg_c = sqrtf(st.m2 / ((float)(H * W) - 1.f) + eps);
  const int64_t si = (int64_t)n * stat_batch_stride + c;
  const float A = alpha * (sigma_s[si] / sg_c) + (1.f - alpha);
  const float B = alpha * mu_s[si] + (1.f - alpha) * st.mean;
  fold_a[(size_t)n * C + c] = A;
  fold_b[(size_t)n * C + c] = fmaf(-st.mean, A, B);
}

// w_out[n][co][tap*Cin + c] = T16(w_k32[co][tap*Cin + c] * A[n][c]); thread = 8 consecutive k, grid (K*Cout/2048, N)
template <typename T16>
__global__ void __launch_bounds__(256)
    adain_fold_weights_kernel(const float* __restrict__ w_k32, const float* __restrict__ fold_a, int Cin, int K,
                              int Cout, T16* __restrict__ w_out, unsigned int* sat_count) {
  const int n = blockIdx.y;
  const size_t i8 = ((size_t)blockIdx.x * 256 + threadIdx.x) * 8;  // first of 8 elements of [Cout][K]
  uint32_t absmax = 0u;
  if (i8 < (size_t)Cout * K) {
    const int k = (int)(i8 % K);
    const int c = k % Cin;  // Cin % 8 == 0: the 8 elements share the tap and have consecutive channels
    const float4 w0 = *reinterpret_cast<const float4*>(w_k32 + i8), w1 = *reinterpret_cast<const float4*>(w_k32 + i8 + 4);
    const float4 a0 = *reinterpret_cast<const float4*>(fold_a + (size_t)n * Cin + c);
    const float4 a1 = *reinterpret_cast<const float4*>(fold_a + (size_t)n * Cin + c + 4);
    uint4 o;
    o.x = pack16x2<T16>(w0.x * a0.x, w0.y * a0.y);
    o.y = pack16x2<T16>(w0.z * a0.z, w0.w * a0.w);
    o.z = pack16x2<T16>(w1.x * a1.x, w1.y * a1.y);
    o.w = pack16x2<T16>(w1.z * a1.z, w1.w * a1.w);
    *reinterpret_cast<uint4*>(w_out + (size_t)n * Cout * K + i8) = o;
    if (sizeof(T16) == 2 && std::is_same<T16, __half>::value) {
      absmax = max16x2<__half>(max16x2<__half>(o.x & 0x7fff7fffu, o.y & 0x7fff7fffu),
                               max16x2<__half>(o.z & 0x7fff7fffu, o.w & 0x7fff7fffu));
    }
  }
  if (std::is_same<T16, __half>::value) {
    const bool hit = (absmax & 0xffffu) >= 0x7bffu || (absmax >> 16) >= 0x7bffu;
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (ballot != 0u && sat_count != nullptr && (threadIdx.x & 31) == 0) atomicAdd(sat_count, (unsigned)__popc(ballot));
  }
}

// b_out[n][co] = bias[co] + sum_c w_tapsum[co][c] * fold_b[n][c]; one warp per (n, co), fixed order
__global__ void __launch_bounds__(256)
    adain_fold_bias_kernel(const float* __restrict__ w_tapsum, const float* __restrict__ fold_b,
                           const float* __restrict__ bias, int Cin, int Cout, float* __restrict__ b_out) {
  const int n = blockIdx.y, co = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (co >= Cout) return;
  float acc = 0.f;
  for (int c = lane; c < Cin; c += 32) acc = fmaf(w_tapsum[(size_t)co * Cin + c], fold_b[(size_t)n * Cin + c], acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) b_out[(size_t)n * Cout + co] = bias[co] + acc;
}

// coef[n * C + c] = {mu_c, A, B}:  out = (x - mu_c) * A + B,
//   A = alpha * sigma_s / sigma_c + (1 - alpha),  B = alpha * mu_s + (1 - alpha) * mu_c
__global__ void __launch_bounds__(256)
    adain_nhwc_coef_kernel(const float2* __restrict__ part, float4* __restrict__ coef, int NC, int C,
                           int HW, const float* __restrict__ mu_s, const float* __restrict__ sigma_s,
                           int64_t stat_batch_stride, float alpha, float eps) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= NC) return;
  const int n = i / C, c = i % C;
  const Wf st = nhwc_merge_partials(part, n, c, C, HW);
  // unbiased like calc_mean_std (function.py:9); HW == 1 -> NaN as the reference
  const float sg_c = sqrtf(st.m2 / ((float)HW - 1.f) + eps);
  const int64_t si = (int64_t)n * stat_batch_stride + c;
  const float ms = mu_s[si], ss = sigma_s[si];
  coef[i] = make_float4(st.mean, alpha * (ss / sg_c) + (1.f - alpha), alpha * ms + (1.f - alpha) * st.mean,
                        0.f);
}

// grid (chunks, N): the affine on the chunk's pixels, out = x * A + (B - mu_c * A); same thread
// mapping as the statistics kernel (VEC channels x every PX_LANES-th pixel)
template <typename T>
__global__ void __launch_bounds__(256)
    adain_nhwc_apply_kernel(ActView<T> in, ActView<T> out, const float4* __restrict__ coef) {
  constexpr int VEC = VecOf<T>::value;
  const int C = in.C, ch_lanes = C / VEC, px_lanes = 256 / ch_lanes;
  const int chunk = blockIdx.x, n = blockIdx.y;
  const int HW = in.H * in.W;
  const int p0 = chunk * kNhwcChunk, p1 = min(HW, p0 + kNhwcChunk);
  const int cl = threadIdx.x % ch_lanes, pl = threadIdx.x / ch_lanes;
  const size_t coff = (size_t)cl * VEC;
  const T* src0 = in.px(n, 0, 0) + coff;
  float A[VEC], B[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const float4 q = coef[(size_t)n * C + coff + k];
    A[k] = q.y, B[k] = fmaf(-q.x, q.y, q.z);
  }
  for (int pb = p0 + pl; pb < p1; pb += kNhwcBatch * px_lanes) {
    Pack<T, VEC> v[kNhwcBatch];
#pragma unroll
    for (int i = 0; i < kNhwcBatch; ++i) {
      const int p = pb + i * px_lanes;
      if (p < p1) {
        const int y = p / in.W, x = p - y * in.W;
        v[i].v = *reinterpret_cast<const decltype(v[i].v)*>(src0 + (size_t)y * in.pitch_y() + (size_t)x * C);
      }
    }
#pragma unroll
    for (int i = 0; i < kNhwcBatch; ++i) {
      const int p = pb + i * px_lanes;
      if (p < p1) {
        const int y = p / in.W, x = p - y * in.W;
        float f[VEC];
        unpack_vec(v[i], f);
#pragma unroll
        for (int k = 0; k < VEC; ++k) f[k] = fmaf(f[k], A[k], B[k]);
        Pack<T, VEC> o;
        pack_vec(o, f);
        *reinterpret_cast<decltype(o.v)*>(out.px(n, y, x) + coff) = o.v;
        if (y == 1 || y == in.H - 2 || x == 1 || x == in.W - 2) {  // reflection-halo aliases
          for_each_halo_alias(y, x, in.H, in.W, [&](int yy, int xx) {
            if (yy != y || xx != x) *reinterpret_cast<decltype(o.v)*>(out.px(n, yy, xx) + coff) = o.v;
          });
        }
      }
    }
  }
}

// =====================================================================================
// layout converters (32 pixels x 32 channels tiles through shared memory)
// =====================================================================================
template <typename T, bool SPLIT = false>
__global__ void __launch_bounds__(256) act_to_nchw_kernel(ActView<T> in, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int HW = in.H * in.W;
  const int C = SPLIT ? in.C / 2 : in.C;  // SPLIT: [hi | lo] halves, value = hi + lo
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32, n = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < C) {
      const int y = p / in.W, x = p - y * in.W;
      v = to_f32(in.px(n, y, x)[c]);
      if (SPLIT) v += to_f32(in.px(n, y, x)[C + c]);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    if (p < HW && c < C) out[((size_t)n * C + c) * HW + p] = tile[tx][i];
  }
}

template <typename T, bool SPLIT = false>
__global__ void __launch_bounds__(256) nchw_to_act_kernel(const float* __restrict__ in,
                                                          ActView<T> out) {
  __shared__ float tile[32][33];
  const int HW = out.H * out.W;
  const int C = SPLIT ? out.C / 2 : out.C;  // SPLIT: the map holds [hi | lo] halves of C logical channels
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32, n = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (p < HW && c < C) ? in[((size_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (p < HW && c < C) {
      const int y = p / out.W, x = p - y * out.W;
      const T v = from_f32<T>(tile[tx][i]);
      const T lo = SPLIT ? from_f32<T>(tile[tx][i] - to_f32(v)) : v;
      for_each_halo_alias(y, x, out.H, out.W, [&](int yy, int xx) {
        out.px(n, yy, xx)[c] = v;
        if (SPLIT) out.px(n, yy, xx)[C + c] = lo;
      });
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_act_kernel(const float* __restrict__ in,
                                                          ActView<T> out, int edge) {
  const size_t total = (size_t)out.N * out.H * out.W * out.C;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int c = (int)(i % out.C);
    size_t p = i / out.C;
    const int x = (int)(p % out.W);
    p /= out.W;
    const int y = (int)(p % out.H);
    const int n = (int)(p / out.H);
    const T v = from_f32<T>(in[i]);
    for_each_halo_alias(y, x, out.H, out.W, [&](int yy, int xx) { out.px(n, yy, xx)[c] = v; }, edge);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) act_to_nhwc_kernel(ActView<T> in, float* __restrict__ out) {
  const size_t total = (size_t)in.N * in.H * in.W * in.C;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int c = (int)(i % in.C);
    size_t p = i / in.C;
    const int x = (int)(p % in.W);
    p /= in.W;
    const int y = (int)(p % in.H);
    const int n = (int)(p / in.H);
    out[i] = to_f32(in.px(n, y, x)[c]);
  }
}

// ToTensor of a uint8 HWC batch: out[n][c][y][x] = float(in[n][y][x][c]) / 255 (IEEE division, as
// torch's div).  One thread per pixel.
__global__ void __launch_bounds__(256) u8_nhwc_to_f32_nchw_kernel(const uint8_t* __restrict__ in,
                                                                  int C, size_t HW, size_t total_px,
                                                                  float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total_px; i += (size_t)gridDim.x * 256) {
    const size_t n = i / HW, p = i - n * HW;
    for (int c = 0; c < C; ++c)
      out[(n * C + c) * HW + p] = __fdiv_rn((float)in[i * C + c], 255.f);
  }
}

__global__ void __launch_bounds__(256) quantize_nchw_to_u8_nhwc_kernel(const float* __restrict__ in,
                                                                       int C, size_t HW, size_t total_px,
                                                                       uint8_t* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total_px; i += (size_t)gridDim.x * 256) {
    const size_t n = i / HW, p = i - n * HW;
    for (int c = 0; c < C; ++c) out[i * C + c] = quantize_u8(in[(n * C + c) * HW + p]);
  }
}

// `transforms.Resize(output_size)` on the output tensor (CCST_OverallStyleTransfer.py:134-135,154-155):
// torch's anti-aliased bilinear interpolation (aten UpSampleKernel / UpSampleBilinear2d antialias,
// align_corners = false), restated: separable triangle filter of support max(scale, 1), weights
// normalised per output index, horizontal pass first, everything in fp32.
struct AaAxis {
  float scale, support, invscale;
  int in_size;
};
__device__ __forceinline__ AaAxis aa_axis(int in_size, int out_size) {
  AaAxis a;
  a.in_size = in_size;
  a.scale = __fdiv_rn((float)in_size, (float)out_size);
  a.support = a.scale >= 1.f ? a.scale : 1.f;
  a.invscale = a.scale >= 1.f ? __fdiv_rn(1.f, a.scale) : 1.f;
  return a;
}
__device__ __forceinline__ void aa_window(const AaAxis& a, int i, int& lo, int& n, float& center) {
  center = __fmul_rn(a.scale, __fadd_rn((float)i, 0.5f));
  lo = max((int)__fadd_rn(__fsub_rn(center, a.support), 0.5f), 0);
  n = min((int)__fadd_rn(__fadd_rn(center, a.support), 0.5f), a.in_size) - lo;
}
__device__ __forceinline__ float aa_weight(const AaAxis& a, int j, int lo, float center) {
  const float x = fabsf(__fmul_rn(__fadd_rn(__fsub_rn((float)(j + lo), center), 0.5f), a.invscale));
  return x < 1.f ? __fsub_rn(1.f, x) : 0.f;
}

__global__ void __launch_bounds__(256) resize_aa_kernel(const float* __restrict__ in, size_t planes, int H,
                                                        int W, int OH, int OW, float* __restrict__ out) {
  const AaAxis ax = aa_axis(W, OW), ay = aa_axis(H, OH);
  const size_t total = planes * (size_t)OH * OW;
  for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const int ox = (int)(idx % OW);
    const int oy = (int)((idx / OW) % OH);
    const float* src = in + (idx / ((size_t)OW * OH)) * (size_t)H * W;
    int x0, nx, y0, ny;
    float cx, cy;
    aa_window(ax, ox, x0, nx, cx);
    aa_window(ay, oy, y0, ny, cy);
    float tx = 0.f, ty = 0.f;
    for (int j = 0; j < nx; ++j) tx = __fadd_rn(tx, aa_weight(ax, j, x0, cx));
    for (int j = 0; j < ny; ++j) ty = __fadd_rn(ty, aa_weight(ay, j, y0, cy));
    float acc = 0.f;
    for (int r = 0; r < ny; ++r) {
      const float* row = src + (size_t)(y0 + r) * W + x0;
      float h = 0.f;  // horizontal pass of this input row
      for (int j = 0; j < nx; ++j) {
        float w = aa_weight(ax, j, x0, cx);
        if (tx != 0.f) w = __fdiv_rn(w, tx);
        h = __fadd_rn(h, __fmul_rn(w, row[j]));
      }
      float wy = aa_weight(ay, r, y0, cy);
      if (ty != 0.f) wy = __fdiv_rn(wy, ty);
      acc = __fadd_rn(acc, __fmul_rn(wy, h));
    }
    out[idx] = acc;
  }
}

__global__ void __launch_bounds__(256)
    raw_to_mean_std_kernel(const float2* __restrict__ raw, int planes, float hw, float eps, int unbiased,
                           float* __restrict__ mean, float* __restrict__ stdv) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= planes) return;
  const float2 r = raw[i];
  if (mean) mean[i] = r.x;
  if (stdv) stdv[i] = sqrtf(r.y / (unbiased ? hw - 1.f : hw) + eps);  // hw == 1, unbiased: NaN as torch.var
}

// nn.MSELoss (net.py:104,124-136): mean of squared differences.  Stage 1: per-block double sums in a fixed
// thread order; stage 2: one block adds the partials in index order -> bit-reproducible.
constexpr int kMseBlocks = 1024;
__global__ void __launch_bounds__(256)
    mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double* __restrict__ part) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float d = a[i] - b[i];
    acc += (double)(d * d);  // the square in fp32 like torch's (input - target) ** 2, the sum in fp64
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}
__global__ void mse_final_kernel(const double* __restrict__ part, int blocks, double n, float* __restrict__ out) {
  double acc = 0.0;
  for (int i = 0; i < blocks; ++i) acc += part[i];
  out[0] = (float)(acc / n);
}

// Pillow's ImagingResampleHorizontal_8bpc / Vertical_8bpc (src/libImaging/Resample.c): 22-bit fixed-point
// coefficients, accumulator seeded with one half, result shifted and clamped to 0..255.  One thread per
// output byte; `stride` = distance in bytes between successive taps of the window.
constexpr int kPilPrecisionBits = 32 - 8 - 2;
__global__ void __launch_bounds__(256)
    resize_pil_pass_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t total, int out_len,
                           size_t inner, size_t in_len, const int* __restrict__ kk, const int* __restrict__ bounds,
                           int ks) {
  // layout of both tensors: [outer][len][inner] with len = in_len (input) / out_len (output) along the
  // resampled axis (inner = C for the horizontal pass, OW * C for the vertical one)
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const size_t in_idx = i % inner;
    const size_t t = i / inner;
    const int o = (int)(t % out_len);
    const size_t outer = t / out_len;
    const int first = bounds[2 * o], cnt = bounds[2 * o + 1];
    const uint8_t* src = in + (outer * in_len + first) * inner + in_idx;
    const int* k = kk + (size_t)o * ks;
    int ss = 1 << (kPilPrecisionBits - 1);
    for (int x = 0; x < cnt; ++x) ss += (int)src[(size_t)x * inner] * k[x];
    ss >>= kPilPrecisionBits;
    out[i] = (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
  }
}

int ew_grid(size_t total) {
  size_t blocks = (total + 255) / 256;
  size_t cap = (size_t)sm_count() * 16;
  return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace

// ------------------------------------------------------------------ launch wrappers
template <typename T>
int launch_conv_first(const float* img, int N, int H, int W, const float* w27, const float* bias64,
                      ActView<T> out, cudaStream_t st) {
  const int tiles_x = (W + kFirstTile - 1) / kFirstTile;
  const size_t blocks = (size_t)N * H * tiles_x;
  CCST_CHECK_ARG(blocks < (1ull << 31), "conv_first: grid too large");
  conv_first_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(img, N, H, W, w27, bias64, out);
  CCST_LAUNCHED();
  return CCST_OK;
}
template <typename T16>
int launch_conv_first_split(const float* img, int N, int H, int W, const float* w27, const float* bias64,
                            ActView<T16> out, cudaStream_t st, unsigned int* sat_count) {
  CCST_CHECK_ARG(out.C == 128, "conv_first(f16x3): the output map holds 64 + 64 channels");
  const int tiles_x = (W + kFirst4Tile - 1) / kFirst4Tile;
  const size_t blocks = (size_t)N * H * tiles_x;
  CCST_CHECK_ARG(blocks < (1ull << 31), "conv_first: grid too large");
  conv_first_px4_kernel<T16, true><<<(unsigned)blocks, 256, 0, st>>>(img, N, H, W, w27, bias64, out, sat_count);
  CCST_LAUNCHED();
  return CCST_OK;
}
template int launch_conv_first_split<__half>(const float*, int, int, int, const float*, const float*, ActView<__half>,
                                             cudaStream_t, unsigned int*);
template int launch_conv_first_split<__nv_bfloat16>(const float*, int, int, int, const float*, const float*,
                                                    ActView<__nv_bfloat16>, cudaStream_t, unsigned int*);
template int launch_conv_first<float>(const float*, int, int, int, const float*, const float*,
                                      ActView<float>, cudaStream_t);
template int launch_conv_first<__nv_bfloat16>(const float*, int, int, int, const float*,
                                              const float*, ActView<__nv_bfloat16>, cudaStream_t);
template int launch_conv_first<__half>(const float*, int, int, int, const float*,
                                              const float*, ActView<__half>, cudaStream_t);

int launch_conv_ffma(ActView<float> in, const float* w, const float* bias, int Cout, int CoutPad,
                     int relu, int epi, ActView<float> out, float* out_nchw, cudaStream_t st) {
  CCST_CHECK_ARG(in.C % kFT_K == 0 && CoutPad % kFT_N == 0, "conv_ffma: channel counts");
  const int tiles_x = (in.W + kFT_W - 1) / kFT_W, tiles_y = (in.H + kFT_H - 1) / kFT_H;
  const size_t blocks = (size_t)in.N * tiles_y * tiles_x * (CoutPad / kFT_N);
  CCST_CHECK_ARG(blocks < (1ull << 31), "conv_ffma: grid too large");
  const size_t smem =
      ((size_t)(kFT_H + 2) * (kFT_W + 2) * kFT_K + (size_t)9 * kFT_K * kFT_N) * sizeof(float);
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_ffma_kernel<EPI_ACT>), (int)smem));
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_ffma_kernel<EPI_ACT_UP2>), (int)smem));
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_ffma_kernel<EPI_NCHW_F32>), (int)smem));
  if (epi == EPI_ACT)
    conv_ffma_kernel<EPI_ACT><<<(unsigned)blocks, 256, smem, st>>>(in, w, bias, Cout, CoutPad, relu,
                                                                   out, out_nchw);
  else if (epi == EPI_ACT_UP2)
    conv_ffma_kernel<EPI_ACT_UP2><<<(unsigned)blocks, 256, smem, st>>>(in, w, bias, Cout, CoutPad,
                                                                       relu, out, out_nchw);
  else if (epi == EPI_NCHW_F32)
    conv_ffma_kernel<EPI_NCHW_F32><<<(unsigned)blocks, 256, smem, st>>>(in, w, bias, Cout, CoutPad,
                                                                        relu, out, out_nchw);
  else
    CCST_CHECK_ARG(false, "conv_ffma: unsupported epilogue %d", epi);
  CCST_LAUNCHED();
  return CCST_OK;
}

template <typename T>
int launch_pool(ActView<T> in, ActView<T> out, cudaStream_t st) {
  const size_t total = (size_t)out.N * out.H * out.W * (in.C / VecOf<T>::value);
  pool_kernel<T><<<ew_grid(total), 256, 0, st>>>(in, out);
  CCST_LAUNCHED();
  return CCST_OK;
}
template int launch_pool<float>(ActView<float>, ActView<float>, cudaStream_t);
template int launch_pool<__nv_bfloat16>(ActView<__nv_bfloat16>, ActView<__nv_bfloat16>,
                                        cudaStream_t);
template int launch_pool<__half>(ActView<__half>, ActView<__half>,
                                        cudaStream_t);

size_t nhwc_scratch_elems(int N, int C, int HW) { return (size_t)N * C * (nhwc_chunks(HW) + 2); }

template <typename T>
int nhwc_geometry_ok(const ActView<T>& in, const char* who) {
  constexpr int VEC = VecOf<T>::value;
  CCST_CHECK_ARG(in.C % VEC == 0 && in.C / VEC <= 256 && 256 % (in.C / VEC) == 0,
                 "%s: C=%d must be %d x a power of two <= 256", who, in.C, VEC);
  CCST_CHECK_ARG(in.N <= 65535, "%s: batch too large", who);
  return CCST_OK;
}
template <typename T>
size_t nhwc_stats_smem(const ActView<T>& in) {
  const int px_lanes = 256 / (in.C / VecOf<T>::value);
  return (size_t)(2 * px_lanes * in.C + px_lanes) * sizeof(float);
}

template <typename T>
int launch_adain_nhwc(ActView<T> in, ActView<T> out, const float* mu_s, const float* sigma_s,
                      int64_t stat_batch_stride, float alpha, float eps, float2* scratch,
                      cudaStream_t st) {
  if (int e = nhwc_geometry_ok(in, "adain_nhwc")) return e;
  const int chunks = nhwc_chunks(in.H * in.W);
  dim3 grid(chunks, in.N);
  const int NC = in.N * in.C;
  float4* coef = reinterpret_cast<float4*>(scratch);          // NC float4 = 2 NC float2
  float2* part = scratch + 2 * (size_t)NC;                    // NC * chunks float2
  nhwc_stats_partial_kernel<T><<<grid, 256, nhwc_stats_smem(in), st>>>(in, part);
  CCST_LAUNCHED();
  adain_nhwc_coef_kernel<<<(NC + 255) / 256, 256, 0, st>>>(part, coef, NC, in.C, in.H * in.W, mu_s,
                                                           sigma_s, stat_batch_stride, alpha, eps);
  CCST_LAUNCHED();
  adain_nhwc_apply_kernel<T><<<grid, 256, 0, st>>>(in, out, coef);
  CCST_LAUNCHED();
  return CCST_OK;
}
template int launch_adain_nhwc<float>(ActView<float>, ActView<float>, const float*, const float*,
                                      int64_t, float, float, float2*, cudaStream_t);
template int launch_adain_nhwc<__nv_bfloat16>(ActView<__nv_bfloat16>, ActView<__nv_bfloat16>,
                                              const float*, const float*, int64_t, float, float,
                                              float2*, cudaStream_t);
template int launch_adain_nhwc<__half>(ActView<__half>, ActView<__half>, const float*, const float*,
                                       int64_t, float, float, float2*, cudaStream_t);

// x3 engines: the affine on a [hi | lo] map of C = in.C / 2 logical channels; x = hi + lo, result split again.
template <typename T>
__global__ void __launch_bounds__(256)
    adain_nhwc_apply_split_kernel(ActView<T> in, ActView<T> out, const float4* __restrict__ coef) {
  constexpr int VEC = VecOf<T>::value;
  const int C = in.C / 2, ch_lanes = C / VEC, px_lanes = 256 / ch_lanes;
  const int chunk = blockIdx.x, n = blockIdx.y;
  const int HW = in.H * in.W;
  const int p0 = chunk * kNhwcChunk, p1 = min(HW, p0 + kNhwcChunk);
  const int cl = threadIdx.x % ch_lanes, pl = threadIdx.x / ch_lanes;
  const size_t coff = (size_t)cl * VEC;
  float A[VEC], B[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const float4 q = coef[(size_t)n * C + coff + k];
    A[k] = q.y, B[k] = fmaf(-q.x, q.y, q.z);
  }
  for (int p = p0 + pl; p < p1; p += px_lanes) {
    const int y = p / in.W, x = p - y * in.W;
    Pack<T, VEC> vh, vl;
    vh.v = *reinterpret_cast<const decltype(vh.v)*>(in.px(n, y, x) + coff);
    vl.v = *reinterpret_cast<const decltype(vl.v)*>(in.px(n, y, x) + C + coff);
    float fh[VEC], fl[VEC];
    unpack_vec(vh, fh);
    unpack_vec(vl, fl);
    Pack<T, VEC> oh, ol;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float v = fmaf(fh[k] + fl[k], A[k], B[k]);
      oh.set(k, v);
      ol.set(k, v - oh.get(k));
    }
    for_each_halo_alias(y, x, in.H, in.W, [&](int yy, int xx) {
      *reinterpret_cast<decltype(oh.v)*>(out.px(n, yy, xx) + coff) = oh.v;
      *reinterpret_cast<decltype(ol.v)*>(out.px(n, yy, xx) + C + coff) = ol.v;
    });
  }
}

template <typename T16>
int launch_adain_nhwc_split(ActView<T16> in, ActView<T16> out, const float* mu_s, const float* sigma_s,
                            int64_t stat_batch_stride, float alpha, float eps, float2* scratch, cudaStream_t st) {
  ActView<T16> logical = in;
  logical.C = in.C / 2;
  if (int e = nhwc_geometry_ok(logical, "adain_nhwc(x3)")) return e;
  const int chunks = nhwc_chunks(in.H * in.W);
  dim3 grid(chunks, in.N);
  const int NC = in.N * logical.C;
  float4* coef = reinterpret_cast<float4*>(scratch);  // NC float4 = 2 NC float2
  float2* part = scratch + 2 * (size_t)NC;            // NC * chunks float2
  nhwc_stats_partial_kernel<T16, true><<<grid, 256, nhwc_stats_smem(logical), st>>>(in, part);
  CCST_LAUNCHED();
  adain_nhwc_coef_kernel<<<(NC + 255) / 256, 256, 0, st>>>(part, coef, NC, logical.C, in.H * in.W, mu_s, sigma_s,
                                                           stat_batch_stride, alpha, eps);
  CCST_LAUNCHED();
  adain_nhwc_apply_split_kernel<T16><<<grid, 256, 0, st>>>(in, out, coef);
  CCST_LAUNCHED();
  return CCST_OK;
}
template int launch_adain_nhwc_split<__half>(ActView<__half>, ActView<__half>, const float*, const float*, int64_t,
                                             float, float, float2*, cudaStream_t);
template int launch_adain_nhwc_split<__nv_bfloat16>(ActView<__nv_bfloat16>, ActView<__nv_bfloat16>, const float*,
                                                    const float*, int64_t, float, float, float2*, cudaStream_t);

// Row-wise apply: CTA = (image n, row y), thread = (VEC channels, every PX_LANES-th pixel of the row).
// No per-pixel index arithmetic, the row's halo aliases are decided once per CTA / per x.
template <typename T>
__global__ void __launch_bounds__(256)
    adain_nhwc_apply_rows_kernel(ActView<T> in, ActView<T> out, const float4* __restrict__ coef) {
  constexpr int VEC = VecOf<T>::value;
  constexpr int kB = 8;
  const int C = in.C, ch_lanes = C / VEC, px_lanes = 256 / ch_lanes;
  const int y = blockIdx.x, n = blockIdx.y;
  const int cl = threadIdx.x % ch_lanes, pl = threadIdx.x / ch_lanes;
  const size_t coff = (size_t)cl * VEC;
  const T* src = in.px(n, y, 0) + coff;
  T* dst = out.px(n, y, 0) + coff;
  // rows that alias this one in the reflection halo: -1 for y == 1, H for y == H - 2
  T* dst_up = (y == 1) ? out.px(n, -1, 0) + coff : nullptr;
  T* dst_dn = (y == in.H - 2) ? out.px(n, in.H, 0) + coff : nullptr;
  float A[VEC], B[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const float4 q = coef[(size_t)n * C + coff + k];
    A[k] = q.y, B[k] = fmaf(-q.x, q.y, q.z);
  }
  for (int xb = pl; xb < in.W; xb += kB * px_lanes) {
    Pack<T, VEC> v[kB];
#pragma unroll
    for (int i = 0; i < kB; ++i) {
      const int x = xb + i * px_lanes;
      if (x < in.W) v[i].v = *reinterpret_cast<const decltype(v[i].v)*>(src + (size_t)x * C);
    }
#pragma unroll
    for (int i = 0; i < kB; ++i) {
      const int x = xb + i * px_lanes;
      if (x < in.W) {
        float f[VEC];
        unpack_vec(v[i], f);
#pragma unroll
        for (int k = 0; k < VEC; ++k) f[k] = fmaf(f[k], A[k], B[k]);
        Pack<T, VEC> o;
        pack_vec(o, f);
        // x's reflection-halo aliases: -1 for x == 1, W for x == W - 2 (both when W == 3)
        auto put = [&](T* row) {
          *reinterpret_cast<decltype(o.v)*>(row + (ptrdiff_t)x * C) = o.v;
          if (x == 1) *reinterpret_cast<decltype(o.v)*>(row - (ptrdiff_t)C) = o.v;
          if (x == in.W - 2) *reinterpret_cast<decltype(o.v)*>(row + (ptrdiff_t)in.W * C) = o.v;
        };
        put(dst);
        if (dst_up) put(dst_up);
        if (dst_dn) put(dst_dn);
      }
    }
  }
}

size_t nhwc_tile_scratch_elems(int N, int C, int H, int W) {
  return (size_t)N * C * 2 + (size_t)N * ((H + 7) / 8) * ((W + 15) / 16) * 4 * C;
}

template <typename T>
int launch_adain_nhwc_tiles(ActView<T> in, ActView<T> out, const float* mu_s, const float* sigma_s,
                            int64_t stat_batch_stride, float alpha, float eps, float2* scratch,
                            cudaStream_t st) {
  if (int e = nhwc_geometry_ok(in, "adain_nhwc")) return e;
  const int NC = in.N * in.C;
  float4* coef = reinterpret_cast<float4*>(scratch);
  adain_nhwc_coef_tiles_kernel<<<NC / 32, 256, 0, st>>>(scratch + 2 * (size_t)NC, coef, in.C, in.H, in.W, mu_s,
                                                        sigma_s, stat_batch_stride, alpha, eps);
  CCST_LAUNCHED();
  dim3 grid(in.H, in.N);
  adain_nhwc_apply_rows_kernel<T><<<grid, 256, 0, st>>>(in, out, coef);
  CCST_LAUNCHED();
  return CCST_OK;
}
template int launch_adain_nhwc_tiles<__nv_bfloat16>(ActView<__nv_bfloat16>, ActView<__nv_bfloat16>,
                                                    const float*, const float*, int64_t, float, float,
                                                    float2*, cudaStream_t);
template int launch_adain_nhwc_tiles<__half>(ActView<__half>, ActView<__half>, const float*, const float*,
                                             int64_t, float, float, float2*, cudaStream_t);

template <typename T16>
int launch_adain_fold(int N, int C, int H, int W, int Cout, float2* scratch, const float* mu_s, const float* sigma_s,
                      int64_t stat_batch_stride, float alpha, float eps, const float* w_k32, const float* w_tapsum,
                      const float* bias, T16* w_out, float* b_out, unsigned int* sat_count, cudaStream_t st) {
  CCST_CHECK_ARG(C % 32 == 0 && N <= 65535 && Cout % 8 == 0, "adain_fold: C=%d N=%d Cout=%d", C, N, Cout);
  const int NC = N * C, K = 9 * C;
  float* fold_a = reinterpret_cast<float*>(scratch);  // the coefficient area of the tile scratch: 4*NC floats
  float* fold_b = fold_a + NC;
  adain_fold_coef_tiles_kernel<<<NC / 32, 256, 0, st>>>(scratch + 2 * (size_t)NC, fold_a, fold_b, C, H, W, mu_s,
                                                        sigma_s, stat_batch_stride, alpha, eps);
  CCST_LAUNCHED();
  const size_t per_img = (size_t)Cout * K;
  dim3 gw((unsigned)((per_img / 8 + 255) / 256), N);
  adain_fold_weights_kernel<T16><<<gw, 256, 0, st>>>(w_k32, fold_a, C, K, Cout, w_out, sat_count);
  CCST_LAUNCHED();
  dim3 gb((Cout + 7) / 8, N);
  adain_fold_bias_kernel<<<gb, 256, 0, st>>>(w_tapsum, fold_b, bias, C, Cout, b_out);
  CCST_LAUNCHED();
  return CCST_OK;
}
template int launch_adain_fold<__nv_bfloat16>(int, int, int, int, int, float2*, const float*, const float*, int64_t,
                                              float, float, const float*, const float*, const float*,
                                              __nv_bfloat16*, float*, unsigned int*, cudaStream_t);
template int launch_adain_fold<__half>(int, int, int, int, int, float2*, const float*, const float*, int64_t, float,
                                       float, const float*, const float*, const float*, __half*, float*,
                                       unsigned int*, cudaStream_t);

int launch_raw_to_mean_std(const float2* raw, int planes, int64_t hw, float eps, int unbiased, float* mean,
                           float* stdv, cudaStream_t st) {
  raw_to_mean_std_kernel<<<(planes + 255) / 256, 256, 0, st>>>(raw, planes, (float)hw, eps, unbiased, mean, stdv);
  CCST_LAUNCHED();
  return CCST_OK;
}

int launch_mse(const float* a, const float* b, int64_t n, double* scratch, float* out, cudaStream_t st) {
  const int64_t want = (n + 255) / 256;
  const int blocks = (int)(want < kMseBlocks ? (want > 0 ? want : 1) : kMseBlocks);
  mse_partial_kernel<<<blocks, 256, 0, st>>>(a, b, n, scratch);
  CCST_LAUNCHED();
  mse_final_kernel<<<1, 1, 0, st>>>(scratch, blocks, (double)n, out);
  CCST_LAUNCHED();
  return CCST_OK;
}

int launch_stats_from_tiles(int N, int C, int H, int W, float2* scratch, cudaStream_t st) {
  const int NC = N * C;
  nhwc_tiles_merge_kernel<<<NC / 32, 256, 0, st>>>(scratch + 2 * (size_t)NC, scratch, C, H, W);
  CCST_LAUNCHED();
  return CCST_OK;
}

// scratch: nhwc_scratch_elems() float2; the merged per-plane {mean, M2} land in its first N*C entries
template <typename T>
int launch_stats_nhwc(ActView<T> in, float2* scratch, cudaStream_t st) {
  if (int e = nhwc_geometry_ok(in, "stats_nhwc")) return e;
  const int HW = in.H * in.W, chunks = nhwc_chunks(HW);
  const int NC = in.N * in.C;
  dim3 grid(chunks, in.N);
  nhwc_stats_partial_kernel<T><<<grid, 256, nhwc_stats_smem(in), st>>>(in, scratch + NC);
  CCST_LAUNCHED();
  nhwc_stats_merge_kernel<<<(NC + 255) / 256, 256, 0, st>>>(scratch + NC, scratch, NC, in.C, HW);
  CCST_LAUNCHED();
  return CCST_OK;
}
template <typename T16>
int launch_stats_nhwc_split(ActView<T16> in, float2* scratch, cudaStream_t st) {
  ActView<T16> logical = in;
  logical.C = in.C / 2;  // geometry and shared memory follow the logical channel count
  if (int e = nhwc_geometry_ok(logical, "stats_nhwc(f16x3)")) return e;
  const int HW = in.H * in.W, chunks = nhwc_chunks(HW);
  const int NC = in.N * logical.C;
  dim3 grid(chunks, in.N);
  nhwc_stats_partial_kernel<T16, true><<<grid, 256, nhwc_stats_smem(logical), st>>>(in, scratch + NC);
  CCST_LAUNCHED();
  nhwc_stats_merge_kernel<<<(NC + 255) / 256, 256, 0, st>>>(scratch + NC, scratch, NC, logical.C, HW);
  CCST_LAUNCHED();
  return CCST_OK;
}
template int launch_stats_nhwc_split<__half>(ActView<__half>, float2*, cudaStream_t);
template int launch_stats_nhwc_split<__nv_bfloat16>(ActView<__nv_bfloat16>, float2*, cudaStream_t);
template int launch_stats_nhwc<float>(ActView<float>, float2*, cudaStream_t);
template int launch_stats_nhwc<__nv_bfloat16>(ActView<__nv_bfloat16>, float2*, cudaStream_t);
template int launch_stats_nhwc<__half>(ActView<__half>, float2*, cudaStream_t);

template <typename T>
int launch_act_to_nchw(ActView<T> in, float* out_nchw, cudaStream_t st) {
  dim3 grid((in.H * in.W + 31) / 32, (in.C + 31) / 32, in.N);
  act_to_nchw_kernel<T><<<grid, 256, 0, st>>>(in, out_nchw);
  CCST_LAUNCHED();
  return CCST_OK;
}
template <typename T16>
int launch_act_to_nchw_split(ActView<T16> in, float* out_nchw, cudaStream_t st) {
  dim3 grid((in.H * in.W + 31) / 32, (in.C / 2 + 31) / 32, in.N);
  act_to_nchw_kernel<T16, true><<<grid, 256, 0, st>>>(in, out_nchw);
  CCST_LAUNCHED();
  return CCST_OK;
}
template int launch_act_to_nchw_split<__half>(ActView<__half>, float*, cudaStream_t);
template int launch_act_to_nchw_split<__nv_bfloat16>(ActView<__nv_bfloat16>, float*, cudaStream_t);
template int launch_act_to_nchw<float>(ActView<float>, float*, cudaStream_t);
template int launch_act_to_nchw<__nv_bfloat16>(ActView<__nv_bfloat16>, float*, cudaStream_t);
template int launch_act_to_nchw<__half>(ActView<__half>, float*, cudaStream_t);

template <typename T>
int launch_nchw_to_act(const float* in_nchw, ActView<T> out, cudaStream_t st) {
  dim3 grid((out.H * out.W + 31) / 32, (out.C + 31) / 32, out.N);
  nchw_to_act_kernel<T><<<grid, 256, 0, st>>>(in_nchw, out);
  CCST_LAUNCHED();
  return CCST_OK;
}
template <typename T16>
int launch_nchw_to_act_split(const float* in_nchw, ActView<T16> out, cudaStream_t st) {
  dim3 grid((out.H * out.W + 31) / 32, (out.C / 2 + 31) / 32, out.N);
  nchw_to_act_kernel<T16, true><<<grid, 256, 0, st>>>(in_nchw, out);
  CCST_LAUNCHED();
  return CCST_OK;
}
template int launch_nchw_to_act_split<__half>(const float*, ActView<__half>, cudaStream_t);
template int launch_nchw_to_act_split<__nv_bfloat16>(const float*, ActView<__nv_bfloat16>, cudaStream_t);
template int launch_nchw_to_act<float>(const float*, ActView<float>, cudaStream_t);
template int launch_nchw_to_act<__nv_bfloat16>(const float*, ActView<__nv_bfloat16>, cudaStream_t);
template int launch_nchw_to_act<__half>(const float*, ActView<__half>, cudaStream_t);

template <typename T>
int launch_nhwc_to_act(const float* in_nhwc, ActView<T> out, cudaStream_t st, int halo_edge) {
  const size_t total = (size_t)out.N * out.H * out.W * out.C;
  nhwc_to_act_kernel<T><<<ew_grid(total), 256, 0, st>>>(in_nhwc, out, halo_edge);
  CCST_LAUNCHED();
  return CCST_OK;
}
template int launch_nhwc_to_act<float>(const float*, ActView<float>, cudaStream_t, int);
template int launch_nhwc_to_act<__nv_bfloat16>(const float*, ActView<__nv_bfloat16>, cudaStream_t, int);
template int launch_nhwc_to_act<__half>(const float*, ActView<__half>, cudaStream_t, int);

int launch_u8_nhwc_to_f32_nchw(const uint8_t* in, int N, int C, int H, int W, float* out, cudaStream_t st) {
  const size_t hw = (size_t)H * W, total = (size_t)N * hw;
  u8_nhwc_to_f32_nchw_kernel<<<ew_grid(total), 256, 0, st>>>(in, C, hw, total, out);
  CCST_LAUNCHED();
  return CCST_OK;
}

int launch_quantize_nchw_to_u8_nhwc(const float* in, int N, int C, int H, int W, uint8_t* out, cudaStream_t st) {
  const size_t hw = (size_t)H * W, total = (size_t)N * hw;
  quantize_nchw_to_u8_nhwc_kernel<<<ew_grid(total), 256, 0, st>>>(in, C, hw, total, out);
  CCST_LAUNCHED();
  return CCST_OK;
}

int launch_resize_pil_u8(const uint8_t* in, int N, int H, int W, int C, int OH, int OW, const int* kx, const int* bx,
                         int ksx, const int* ky, const int* by, int ksy, uint8_t* tmp, uint8_t* out, cudaStream_t st) {
  // Pillow resamples horizontally first, then vertically (ImagingResample); a pass whose size does not
  // change is skipped there and is the identity here (one coefficient of 1 << 22 per output index)
  const size_t t1 = (size_t)N * H * OW * C, t2 = (size_t)N * OH * OW * C;
  resize_pil_pass_kernel<<<ew_grid(t1), 256, 0, st>>>(in, tmp, t1, OW, (size_t)C, (size_t)W, kx, bx, ksx);
  CCST_LAUNCHED();
  resize_pil_pass_kernel<<<ew_grid(t2), 256, 0, st>>>(tmp, out, t2, OH, (size_t)OW * C, (size_t)H, ky, by, ksy);
  CCST_LAUNCHED();
  return CCST_OK;
}

int launch_resize_aa(const float* in, int64_t planes, int H, int W, int OH, int OW, float* out, cudaStream_t st) {
  const size_t total = (size_t)planes * OH * OW;
  resize_aa_kernel<<<ew_grid(total), 256, 0, st>>>(in, (size_t)planes, H, W, OH, OW, out);
  CCST_LAUNCHED();
  return CCST_OK;
}

template <typename T>
int launch_act_to_nhwc(ActView<T> in, float* out_nhwc, cudaStream_t st) {
  const size_t total = (size_t)in.N * in.H * in.W * in.C;
  act_to_nhwc_kernel<T><<<ew_grid(total), 256, 0, st>>>(in, out_nhwc);
  CCST_LAUNCHED();
  return CCST_OK;
}
template int launch_act_to_nhwc<float>(ActView<float>, float*, cudaStream_t);
template int launch_act_to_nhwc<__nv_bfloat16>(ActView<__nv_bfloat16>, float*, cudaStream_t);
template int launch_act_to_nhwc<__half>(ActView<__half>, float*, cudaStream_t);

}  // namespace ccst
