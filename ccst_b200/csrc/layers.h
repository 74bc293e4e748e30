// Internal (non-ABI) launch wrappers shared by api.cu / layers.cu / conv_umma_impl.cuh.
#pragma once
#include "common.cuh"

namespace ccst {

enum Epilogue {
  EPI_ACT = 0,      // bias (+ReLU) -> NHWC activation with reflection halo
  EPI_ACT_UP2 = 1,  // same, every pixel replicated 2x2 (nearest upsample fused into the store)
  EPI_NCHW_F32 = 2, // bias (+ReLU) -> caller's NCHW fp32 tensor (last decoder conv)
  EPI_ACT_POOL = 3, // bias + ReLU + 2x2 ceil-mode max-pool fused into the store (tcgen05 path)
  EPI_ACT_STATS = 5,  // tcgen05 path, EPI_ACT + per-(8x16 pixel tile, row quarter, channel) {mean, M2} of the
                      // stored values written to `tile_stats` (relu4_1 statistics for AdaIN / the
                      // overall-style accumulation without another pass over the feature map)
  EPI_UPS = 4,      // tcgen05 path: the INPUT is the low-resolution activation (replicate halo) of a
                    // nearest-x2 upsample; conv(reflect_pad(upsample(S))) is computed as four 2x2
                    // phase convolutions with pre-summed weights (16 instead of 36 tap-GEMMs per source
                    // pixel) and stored like EPI_ACT at twice the input resolution
};

// conv1_1 with the 1x1 colour conv folded in.  img: NCHW fp32 [N,3,H,W]; w27: [27][64] fp32
// (k = (r*3+s)*3 + ci), bias64.
template <typename T>
int launch_conv_first(const float* img, int N, int H, int W, const float* w27, const float* bias64,
                      ActView<T> out, cudaStream_t st);
// x3 engines: the same fp32 FFMA conv1_1, result stored as [hi | lo] 16-bit halves (out.C == 128)
template <typename T16>
int launch_conv_first_split(const float* img, int N, int H, int W, const float* w27, const float* bias64,
                            ActView<T16> out, cudaStream_t st, unsigned int* sat_count = nullptr);

// fp32 FFMA implicit GEMM.  w: [9][Cin][CoutPad64] fp32, bias [CoutPad64].
int launch_conv_ffma(ActView<float> in, const float* w, const float* bias, int Cout, int CoutPad,
                     int relu, int epi, ActView<float> out, float* out_nchw, cudaStream_t st);

template <typename T>
int launch_pool(ActView<T> in, ActView<T> out, cudaStream_t st);

// float2 elements of the scratch buffer the two launchers below need (chunk partials + merged planes)
size_t nhwc_scratch_elems(int N, int C, int HW);

template <typename T>
int launch_adain_nhwc(ActView<T> in, ActView<T> out, const float* mu_s, const float* sigma_s,
                      int64_t stat_batch_stride, float alpha, float eps, float2* scratch,
                      cudaStream_t st);

// Variants fed by the tile statistics a conv wrote with EPI_ACT_STATS: scratch = [2*N*C float2 coef / raw]
// [tile_part: tiles*4*C float2], tiles = N * ceil(H/8) * ceil(W/16); the conv is given scratch + 2*N*C.
size_t nhwc_tile_scratch_elems(int N, int C, int H, int W);
template <typename T>
int launch_adain_nhwc_tiles(ActView<T> in, ActView<T> out, const float* mu_s, const float* sigma_s,
                            int64_t stat_batch_stride, float alpha, float eps, float2* scratch,
                            cudaStream_t st);
int launch_stats_from_tiles(int N, int C, int H, int W, float2* scratch, cudaStream_t st);

// AdaIN folded into the conv that consumes it (dec1): from the tile statistics in `scratch` (layout of
// nhwc_tile_scratch_elems) and the style statistics, per image n
//   w_out[n][co][k = tap*Cin + c] = w_k32[co][k] * A[n][c]                       (rounded once to T16)
//   b_out[n][co] = bias[co] + sum_c w_tapsum[co][c] * (B[n][c] - mu_c[n][c] * A[n][c])
// with out = (x - mu_c) * A + B the AdaIN + alpha-blend affine of function.py:26-33 /
// CCST_OverallStyleTransfer.py:44-45.  The first 2*N*C floats of `scratch` are overwritten.
template <typename T16>
int launch_adain_fold(int N, int C, int H, int W, int Cout, float2* scratch, const float* mu_s, const float* sigma_s,
                      int64_t stat_batch_stride, float alpha, float eps, const float* w_k32, const float* w_tapsum,
                      const float* bias, T16* w_out, float* b_out, unsigned int* sat_count, cudaStream_t st);

// per-(n,c) {mean, M2} of an activation -> scratch[0 .. N*C)
template <typename T>
int launch_stats_nhwc(ActView<T> in, float2* scratch, cudaStream_t st);
// x3 engines: the maps hold [hi | lo] halves of C / 2 logical channels; value = hi + lo
template <typename T16>
int launch_stats_nhwc_split(ActView<T16> in, float2* scratch, cudaStream_t st);
template <typename T16>
int launch_act_to_nchw_split(ActView<T16> in, float* out_nchw, cudaStream_t st);
template <typename T16>
int launch_nchw_to_act_split(const float* in_nchw, ActView<T16> out, cudaStream_t st);
// calc_mean_std + AdaIN + alpha blend on a [hi | lo] map (scratch: nhwc_scratch_elems of the logical channels)
template <typename T16>
int launch_adain_nhwc_split(ActView<T16> in, ActView<T16> out, const float* mu_s, const float* sigma_s,
                            int64_t stat_batch_stride, float alpha, float eps, float2* scratch, cudaStream_t st);

// raw[i] = {mean, M2} of a plane of `hw` values -> mean[i], std[i] = sqrt(M2 / (hw - unbiased) + eps)
int launch_raw_to_mean_std(const float2* raw, int planes, int64_t hw, float eps, int unbiased, float* mean,
                           float* stdv, cudaStream_t st);
// out[0] = mean((a - b)^2) over n floats (nn.MSELoss), deterministic two-stage sum; scratch: 1024 doubles
int launch_mse(const float* a, const float* b, int64_t n, double* scratch, float* out, cudaStream_t st);

int merge_raw_into_state(const float2* raw, int N, int C, int64_t hw, double* d_state,
                         cudaStream_t st);

template <typename T>
int launch_act_to_nchw(ActView<T> in, float* out_nchw, cudaStream_t st);

// Image I/O around the path: ToTensor of a uint8 HWC batch (cjm_util/data_helper.py:45,
// torchvision to_tensor: float(u) / 255) -> NCHW fp32, and save_image's quantisation of an NCHW fp32
// batch (torchvision utils.save_image: mul(255).add(0.5).clamp(0,255).to(uint8)) -> NHWC uint8.
int launch_u8_nhwc_to_f32_nchw(const uint8_t* in, int N, int C, int H, int W, float* out, cudaStream_t st);
int launch_quantize_nchw_to_u8_nhwc(const float* in, int N, int C, int H, int W, uint8_t* out, cudaStream_t st);
// Pillow's 8-bit bilinear resample of a uint8 NHWC batch (the loader's `transforms.Resize((S, S))` on the PIL
// image, cjm_util/data_helper.py:45-49): horizontal pass into `tmp` [N,H,OW,C], vertical pass into `out`
// [N,OH,OW,C]; kx / ky: fixed-point coefficients [out][ks], bx / by: {first input index, count} per output index.
int launch_resize_pil_u8(const uint8_t* in, int N, int H, int W, int C, int OH, int OW, const int* kx, const int* bx,
                         int ksx, const int* ky, const int* by, int ksy, uint8_t* tmp, uint8_t* out, cudaStream_t st);
// torch's anti-aliased bilinear resize of `planes` H x W fp32 planes to OH x OW (transforms.Resize on a tensor)
int launch_resize_aa(const float* in, int64_t planes, int H, int W, int OH, int OW, float* out, cudaStream_t st);
template <typename T>
int launch_nchw_to_act(const float* in_nchw, ActView<T> out, cudaStream_t st);
// halo_edge: 1 = reflection halo, 0 = replicate halo (see for_each_halo_alias)
template <typename T>
int launch_nhwc_to_act(const float* in_nhwc, ActView<T> out, cudaStream_t st, int halo_edge = 1);
template <typename T>
int launch_act_to_nhwc(ActView<T> in, float* out_nhwc, cudaStream_t st);

// tcgen05 / TMA implicit GEMM (conv_umma_impl.cuh, instantiated per operand type in
// conv_umma_{bf16,f16}.cu), T16 = __nv_bfloat16 or __half operands, fp32 accumulation in TMEM.
template <typename T16>
struct UmmaConvArgs {
  ActView<T16> in{}, out{};
  const T16* wk = nullptr;     // [CoutPad][9*Cin] K-major (k = tap*Cin + c); per_sample: [N][Cout][9*Cin]
  const T16* wk_sm = nullptr;  // Cout == 64 only: the same weights packed [192 = (s, co)][3*Cin = (r, c)]
  const T16* wk_up = nullptr;  // EPI_UPS only: phase weights [4 = (a, b)][Cout][4*Cin], k = (dy*2 + dx)*Cin + c
  const float* bias = nullptr; // fp32 [CoutPad]; per_sample: [N][Cout]
  int Cout = 0, CoutPad = 0;
  int relu = 1;
  int epi = EPI_ACT;
  int halo_edge = 1;           // halo the epilogue writes around `out` (1 reflection, 0 replicate; EPI_ACT only)
  float* out_nchw = nullptr;   // EPI_NCHW_F32 (last conv)
  uint8_t* out_u8 = nullptr;   // EPI_NCHW_F32 only, may be NULL: store NHWC uint8 quantised like save_image instead
  float2* tile_stats = nullptr;       // EPI_ACT_STATS
  unsigned int* sat_count = nullptr;  // device counter of f16 stores that hit the +-65504 clamp (may be NULL)
  bool per_sample = false;     // image n uses weights wk[n] / bias[n] (AdaIN folded into dec1)
  // x3 engines (conv_x3.cuh; EPI_ACT / _POOL / _UP2 / NCHW_F32): `in` and `out` hold [hi | lo] halves (2 x the
  // logical channel counts), wk_x3 = [ceil(Cout/64) tiles][hi | lo][64 co][9 * Cin] K-major (k = tap * Cin + c)
  // with hi + lo = w * 2^k, out_scale = 2^-k
  bool split = false;
  const T16* wk_x3 = nullptr;
  const T16* wk_x3_up = nullptr;  // EPI_UPS on the x3 engines: [4 phases][tiles][hi | lo][64][4 * Cin]
  float out_scale = 1.f;
};
template <typename T16>
int launch_conv_umma(const UmmaConvArgs<T16>& a, cudaStream_t st);

// conv1_1 (+ folded 1x1) on tcgen05: thread-built im2col rows (K = 27 padded to 32).
// wk: [64][32] T16 K-major, bias fp32 [64].  img_u8 != nullptr: the window is read from the loader's uint8 HWC batch
// [N,H,W,3] itself (ToTensor fused into the loader; first_u8_ok(img_u8, W) must hold) and `img` is unused.
template <typename T16>
int launch_conv_first_umma(const float* img, int N, int H, int W, const T16* wk, const float* bias,
                           ActView<T16> out, cudaStream_t st, unsigned int* sat_count = nullptr,
                           const uint8_t* img_u8 = nullptr);
// conv1_1 of the x3 engines on tcgen05: wk_x3 [128 = hi | lo rows][64 = 2 x (27 padded to 32) k] of w * 2^e,
// out_scale = 2^-e, out = [hi | lo] map (C == 128).  fp32 image: W % 4 == 0 and a 16-byte aligned base only.
template <typename T16>
int launch_conv_first_x3(const float* img, int N, int H, int W, const T16* wk_x3, float out_scale, const float* bias,
                         ActView<T16> out, cudaStream_t st, unsigned int* sat_count = nullptr,
                         const uint8_t* img_u8 = nullptr);
// uint8 HWC rows that TMA can fetch as 4-byte words: 3 W bytes per row a multiple of 16, 16-byte aligned base
inline bool first_u8_ok(const uint8_t* img_u8, int W) {
  return W % 16 == 0 && (reinterpret_cast<uintptr_t>(img_u8) & 15) == 0;
}

}  // namespace ccst
