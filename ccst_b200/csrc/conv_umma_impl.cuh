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
#include "conv_first.cuh"
#include "conv_last.cuh"
#include "conv_main.cuh"
#include "conv_smerge.cuh"
#include "conv_ups4.cuh"
#include "conv_x3.cuh"

namespace ccst {

template <typename T16>
int launch_conv_umma(const UmmaConvArgs<T16>& a, cudaStream_t st) {
  const ActView<T16>& in = a.in;
  const int Cout = a.Cout, epi = a.epi;
  CCST_CHECK_ARG(in.C % kBlockK == 0, "conv_umma: Cin=%d must be a multiple of 64", in.C);
  ConvParams<T16> p;
  memset(&p, 0, sizeof(p));
  p.N = in.N, p.H = in.H, p.W = in.W, p.Cin = in.C;
  p.Cout = Cout, p.CoutPad = a.CoutPad;
  p.tiles_x = (in.W + kTileW - 1) / kTileW;
  p.tiles_y = (in.H + kTileH - 1) / kTileH;
  p.relu = a.relu;
  p.halo_edge = a.halo_edge;
  p.bias = a.bias;
  p.out = a.out;
  p.out_nchw = a.out_nchw;
  p.out_u8 = a.out_u8;
  p.tile_stats = a.tile_stats;
  p.sat_count = a.sat_count;
#ifdef CCST_DEV
  p.ablate = dev_env_int("CCST_ABLATE", 0);
#endif
  CCST_CHECK_ARG((epi == EPI_ACT_STATS) == (a.tile_stats != nullptr), "conv_umma: tile_stats goes with EPI_ACT_STATS");
  CCST_CHECK_ARG(a.out_u8 == nullptr || epi == EPI_NCHW_F32, "conv_umma: uint8 store is the last conv's");
  CCST_CHECK_ARG(a.halo_edge == 1 || (a.halo_edge == 0 && epi == EPI_ACT),
                 "conv_umma: a replicate halo is only written by the plain epilogue");
  if (a.split) return launch_x3<T16>(a, p, st);  // x3 engines: split operands (conv_x3.cuh)
  if (epi == EPI_NCHW_F32) {
    // the last decoder conv (64 -> 3): filter rows by operand shifts, filter columns in N + two shuffles
    CCST_CHECK_ARG(a.CoutPad == 16 && Cout <= 3 && in.C == kBlockK && !a.per_sample,
                   "conv_umma: the NCHW epilogue is the 64 -> (<= 3) channel last conv");
    return launch_last_rows<T16>(in, a.wk, p, st);
  }
  CCST_CHECK_ARG(a.CoutPad == Cout && Cout % 64 == 0, "conv_umma: Cout=%d must be a multiple of 64", Cout);
  const int BN = Cout >= 256 ? 256 : Cout;  // 64, 128, 256
  CCST_CHECK_ARG(BN == 64 || BN == 128 || BN == 256, "conv_umma: unsupported Cout=%d", Cout);
  p.n_tiles = a.CoutPad / BN;
  const int64_t m_tiles = (int64_t)in.N * p.tiles_x * p.tiles_y;
  CCST_CHECK_ARG(m_tiles * p.n_tiles * 4 < (1ll << 31), "conv_umma: too many tiles");
  p.m_tiles = (int)m_tiles;
  if (a.per_sample) {
    // AdaIN folded into this conv: image n multiplies with its own weights wk[n] and adds bias[n]
    CCST_CHECK_ARG(epi == EPI_ACT && BN == 256, "conv_umma: per-sample weights exist for the plain N = 256 kernel");
    p.w_rows_per_n = a.CoutPad, p.bias_per_n = a.CoutPad;
  }
  CUtensorMap ma;
  if (int e = make_act_map(&ma, in)) return e;
  if (epi == EPI_UPS) {
    // fused nearest-x2 upsample: `in` is the low-resolution map, `out` twice its size
    CCST_CHECK_ARG(a.wk_up != nullptr, "conv_umma: EPI_UPS needs the phase-packed weights");
    CCST_CHECK_ARG(a.out.H == 2 * in.H && a.out.W == 2 * in.W, "conv_umma: EPI_UPS output must be 2x the input");
    // 64 -> 64: all four phases per tile over one linear slab (see conv_ups4_kernel)
    if (BN == 64 && in.C == kBlockK) return launch_ups4<T16>(in, a.wk_up, p, st);
    return launch_main<T16>(ma, a.wk_up, p, BN, epi, st);
  }
  // s-merged kernel (N = 192 per operand view instead of 64) for every 64-channel layer: at N = 64
  // the tap-by-tap kernel is capped at 50 % of the tensor peak by the A-operand fetch.  Measured
  // batch 32 @512^2: conv1_2 (fused pool, pair) 0.635 -> 0.58 ms, dec7 (pair) 0.32 -> 0.25 ms.
  if (BN == 64) {
    CCST_CHECK_ARG(a.wk_sm != nullptr, "conv_umma: 64-channel layers need the s-merged weight pack");
    return launch_smerge<T16>(ma, a.wk_sm, p, epi, st);
  }
  return launch_main<T16>(ma, a.wk, p, BN, epi, st);
}
// (one operand type per translation unit -- conv_umma_bf16.cu / conv_umma_f16.cu -- so that the two
// halves of the template instantiations compile in parallel)
#if CCST_INST_BF16
template int launch_conv_umma<__nv_bfloat16>(const UmmaConvArgs<__nv_bfloat16>&, cudaStream_t);
#endif
#if CCST_INST_F16
template int launch_conv_umma<__half>(const UmmaConvArgs<__half>&, cudaStream_t);
#endif

// TMA map of the input image for the warp-specialised conv1_1 kernels: the fp32 NCHW tensor (box {136 col, 3 row,
// 3 ch}) or, u8 != nullptr, the loader's uint8 HWC batch as rows of 3 W / 4 four-byte words (box {104 words, 3 rows})
inline int make_image_map(CUtensorMap* mi, const float* img, const uint8_t* u8, int N, int H, int W) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return CCST_ECUDA;
  }
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r;
  if (u8) {
    const cuuint64_t row = (cuuint64_t)W * 3;
    const cuuint64_t dims[4] = {row / 4, (cuuint64_t)H, 1, (cuuint64_t)N};
    const cuuint64_t strides[3] = {row, row * H, row * H};
    const cuuint32_t box[4] = {kU8RowBytes / 4, 3, 1, 1};
    r = enc(mi, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, (void*)u8, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
    const cuuint32_t box[4] = {kF2WinCols, 3, 3, 1};
    r = enc(mi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(image %dx3x%dx%d%s) failed: CUresult %d", N, H, W, u8 ? ", uint8 HWC" : "", (int)r);
    return CCST_ECUDA;
  }
  return CCST_OK;
}

// img: fp32 NCHW image, or (img_u8 != nullptr) the uint8 HWC batch read directly -- ccst_first_u8_ok() must hold
template <typename T16>
int launch_conv_first_umma(const float* img, int N, int H, int W, const T16* wk, const float* bias,
                           ActView<T16> out, cudaStream_t st, unsigned int* sat_count, const uint8_t* img_u8) {
  FirstParams<T16> p;
  p.img = img, p.N = N, p.H = H, p.W = W, p.wk = wk, p.bias = bias, p.out = out;
  p.sat_count = sat_count;
  p.tiles_x = (W + kFirstPx - 1) / kFirstPx;
  const int64_t total = (int64_t)N * H * p.tiles_x;
  CCST_CHECK_ARG(total < (1ll << 31), "conv_first_umma: too many tiles");
  p.total_tiles = (int)total;
  CUtensorMap mo;
  if (int e = make_out_map(&mo, out, 0, 0, 1, 1, 32, 1)) return e;  // one warp's quarter
  CCST_CHECK_ARG(img_u8 == nullptr || first_u8_ok(img_u8, W), "conv_first_umma: uint8 rows must be 16-byte aligned");
  if (img_u8 || (W % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0)) {
    // rows are 16-byte aligned: TMA-fed warp-specialised kernel
    CUtensorMap mi;
    if (int e = make_image_map(&mi, img, img_u8, N, H, W)) return e;
    const int64_t cap2 = (int64_t)sm_count() * 2;
    const int grid2 = (int)(total < cap2 ? total : cap2);
    if (img_u8) {
      CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_first_umma_ws_kernel<T16, true>), kF2Smem));
      CCST_CUDA(launch_conv(conv_first_umma_ws_kernel<T16, true>, grid2, kF2Threads, kF2Smem, st, 1, mi, mo, p));
    } else {
      CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_first_umma_ws_kernel<T16, false>), kF2Smem));
      CCST_CUDA(launch_conv(conv_first_umma_ws_kernel<T16, false>, grid2, kF2Threads, kF2Smem, st, 1, mi, mo, p));
    }
    CCST_LAUNCHED();
    return CCST_OK;
  }
  // image rows that are not 16-byte aligned (W % 4 != 0) cannot be fetched by TMA: cp.async windows
  CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_first_umma_kernel<T16>), kFirstSmem));
  const int64_t cap = (int64_t)sm_count() * 4;
  const int grid = (int)(total < cap ? total : cap);
  conv_first_umma_kernel<T16><<<grid, kFirstPx, kFirstSmem, st>>>(mo, p);
  CCST_LAUNCHED();
  return CCST_OK;
}

// conv1_1 of the x3 engines on the tensor pipe (conv_first_x3_ws_kernel): wk_x3 = [128 rows = hi | lo][64 k], the 27
// weights * 2^e repeated under k = 0..31 and k = 32..63; out = [hi | lo] map of 64 logical channels.  Needs TMA-
// fetchable image rows (fp32: W % 4 == 0 and a 16-byte aligned base; uint8: first_u8_ok), CCST_EINVAL otherwise
// (the caller falls back to the CUDA-core kernel).
template <typename T16>
int launch_conv_first_x3(const float* img, int N, int H, int W, const T16* wk_x3, float out_scale, const float* bias,
                         ActView<T16> out, cudaStream_t st, unsigned int* sat_count, const uint8_t* img_u8) {
  CCST_CHECK_ARG(out.C == 128 && (img_u8 ? first_u8_ok(img_u8, W)
                                         : (W % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0)),
                 "conv_first_x3: needs a [hi | lo] map of 64 channels and 16-byte aligned image rows");
  FirstParams<T16> p;
  p.img = img, p.N = N, p.H = H, p.W = W, p.wk = wk_x3, p.bias = bias, p.out = out;
  p.sat_count = sat_count;
  p.tiles_x = (W + kFirstPx - 1) / kFirstPx;
  const int64_t total = (int64_t)N * H * p.tiles_x;
  CCST_CHECK_ARG(total < (1ll << 31), "conv_first_x3: too many tiles");
  p.total_tiles = (int)total;
  CUtensorMap mo, mi;
  if (int e = make_out_map(&mo, out, 0, 0, 1, 1, 32, 1)) return e;  // one warp's quarter, 64 of the 128 channels
  if (int e = make_image_map(&mi, img, img_u8, N, H, W)) return e;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  if (img_u8) {
    CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_first_x3_ws_kernel<T16, true>), kF3Smem));
    CCST_CUDA(launch_conv(conv_first_x3_ws_kernel<T16, true>, grid, kF3Threads, kF3Smem, st, 1, mi, mo, p, out_scale));
  } else {
    CCST_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_first_x3_ws_kernel<T16, false>), kF3Smem));
    CCST_CUDA(launch_conv(conv_first_x3_ws_kernel<T16, false>, grid, kF3Threads, kF3Smem, st, 1, mi, mo, p, out_scale));
  }
  CCST_LAUNCHED();
  return CCST_OK;
}
#if CCST_INST_BF16
template int launch_conv_first_x3<__nv_bfloat16>(const float*, int, int, int, const __nv_bfloat16*, float, const float*,
                                                 ActView<__nv_bfloat16>, cudaStream_t, unsigned int*, const uint8_t*);
#endif
#if CCST_INST_F16
template int launch_conv_first_x3<__half>(const float*, int, int, int, const __half*, float, const float*,
                                          ActView<__half>, cudaStream_t, unsigned int*, const uint8_t*);
#endif

#if CCST_INST_BF16
template int launch_conv_first_umma<__nv_bfloat16>(const float*, int, int, int,
                                                   const __nv_bfloat16*, const float*,
                                                   ActView<__nv_bfloat16>, cudaStream_t, unsigned int*, const uint8_t*);
#endif
#if CCST_INST_F16
template int launch_conv_first_umma<__half>(const float*, int, int, int, const __half*,
                                            const float*, ActView<__half>, cudaStream_t, unsigned int*, const uint8_t*);
#endif

}  // namespace ccst
