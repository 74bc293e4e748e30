/*
 * ccst_b200.h -- C ABI of libccst_b200.so: the B200 (sm_100a) implementation of
 * the AdaIN style-transfer hot path of JeremyCJM/CCST.
 *
 * The reference has no FFI/plugin interface (it is pure Python calling PyTorch),
 * so these entry points are what a binding for the reference's hot-path callables
 * would bind; each one names the reference code it replaces (paths relative to
 * /root/reference/style_transfer/AdaIN).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller; h_* is a
 *     host pointer.  Tensors are contiguous NCHW fp32 exactly as the reference
 *     passes them (function.py:9 requires contiguity for .view()).
 *   - outputs are caller-allocated; the library never returns memory it owns,
 *     except through the opaque ccst_handle (packed weights + activation arena).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *     All work is enqueued asynchronously; no entry point synchronises unless
 *     its comment says so.
 *   - return value: 0 on success, a negative CCST_E* code otherwise;
 *     ccst_last_error() returns a thread-local message for the last failure.
 *   - there is no CPU fallback: on a device that is not compute capability 10.x
 *     every compute entry point returns CCST_EARCH.
 */
#ifndef CCST_B200_H_
#define CCST_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCST_ABI_VERSION 2

#define CCST_OK 0
#define CCST_EINVAL (-1)  /* bad shape / null pointer / bad argument          */
#define CCST_EARCH (-2)   /* device is not sm_100                             */
#define CCST_ECUDA (-3)   /* CUDA runtime/driver error (see ccst_last_error)  */
#define CCST_ESTATE (-4)  /* handle not ready (weights missing ...)           */

/* precision of the encoder/decoder convolutions */
#define CCST_PREC_FP32 0 /* fp32 activations + fp32 FFMA implicit GEMM (validation mode, 1e-4) */
#define CCST_PREC_BF16 1 /* bf16 activations + tcgen05/TMEM implicit GEMM fed by TMA (fp32 accum) */
#define CCST_PREC_FP16 2 /* same kernels with f16 operands (11-bit significand, saturating stores) */
/* x3 engines: fp32-grade results on the 16-bit tensor pipe, every entry point.  Activations and weights are split
 * into 16-bit high and low parts and a * w = (a_hi + a_lo) * (w_hi + w_lo) runs as one tcgen05 implicit GEMM over
 * 2x the channels and 2x the filter rows with fp32 accumulation (partial sums promoted to registers every 12 MMAs).
 * FP16X3: f16 halves, 22 significand bits -- relu4_1 statistics meet the 1e-5 bar of the fp32 reference and the
 * stylised image its 1e-4 bar at ~8x the rate of CCST_PREC_FP32; activations must stay inside the f16 range (the
 * saturation counter applies).  BF16X3: bf16 halves, 16 significand bits at the fp32 exponent range -- the
 * tensor-core mode for weights whose activations leave the f16 range; meets the 1e-2 image bar that single bf16
 * operands (CCST_PREC_BF16) miss. */
#define CCST_PREC_FP16X3 3
#define CCST_PREC_BF16X3 4

typedef struct ccst_handle ccst_handle;

int ccst_abi_version(void);
const char* ccst_last_error(void);
/* 0 if `device` is an sm_100 GPU, CCST_EARCH otherwise. */
int ccst_check_device(int device);
/* Make `device` current for the calling thread inside the library (the library links its own
 * CUDA runtime; hosts that switch devices call this before enqueueing work). */
int ccst_set_device(int device);

/* ------------------------------------------------------------------------
 * Feature statistics
 * ---------------------------------------------------------------------- */

/* calc_mean_std(feat, eps)  [function.py:4-13]
 * per (n,c) plane of `hw` contiguous floats: mean and sqrt(var + eps);
 * unbiased != 0 divides by (hw-1) as torch.var does (hw == 1 -> NaN, as the
 * reference), unbiased == 0 divides by hw.  d_mean/d_std hold `planes` floats.
 * Either output may be NULL. */
int ccst_stats_nchw_f32(const float* d_x, int64_t planes, int64_t hw, float eps, int unbiased,
                        float* d_mean, float* d_std, void* stream);

/* Running per-channel Welford state of one client
 * [mean_std_computation_effcientMem.py:117 `all_feat_sum, all_feat_square_sum,
 * all_count`], kept on the device as 2+2C doubles: {count, mean[C], M2[C], ticket}.  The last word is
 * scratch of the accumulating kernels (the block that finishes last publishes the new count through it;
 * it is zero between calls).  Zero the whole buffer (cudaMemset) to start a client. */

/* calc_sum(feat) + `all_* += ...`  [mean_std_computation_effcientMem.py:103-115,129-131]
 * Folds the batch d_x = [N,C,HW] into d_state with one pass over d_x (per-plane two-pass statistics in
 * fp32, Chan merges over N and into the state in fp64, fixed order).  Planes of up to 16 KiB with
 * 16-byte alignment run as ONE launch in which every CTA owns whole channels; other shapes take a
 * statistics launch + a merge launch through d_scratch, which must hold 2*N*C floats. */
int ccst_welford_accumulate_nchw_f32(const float* d_x, int N, int C, int64_t hw, double* d_state,
                                     float* d_scratch, void* stream);

/* finalise  [mean_std_computation_effcientMem.py:135-137, CCST_SingleStyleTransfer.py:201-203]
 * mean = state.mean, std = sqrt(M2/count + eps) (biased), as fp32 [C]. */
int ccst_welford_finalize(const double* d_state, int C, float eps, float* d_mean, float* d_std,
                          void* stream);

/* the batch-wide `calc_mean_std` of mean_std_computation_effcientMem.py:89-101 (defined there, never
 * called): per channel over N*H*W, torch's default UNBIASED variance: std = sqrt(M2/(count-1) + eps). */
int ccst_welford_finalize_unbiased(const double* d_state, int C, float eps, float* d_mean,
                                   float* d_std, void* stream);

/* calc_sum's return values from a state: sum = n*mean, sqsum = M2 + n*mean^2 (fp32 [C]). */
int ccst_welford_to_sums(const double* d_state, int C, float* d_sum, float* d_sqsum, void* stream);

/* State <-> exactly-summable moments {n, n*mean[C], M2[C]+n*mean[C]^2} (fp64), the payload of the
 * single all-reduce(sum) that merges the per-GPU partials of one client. */
int ccst_welford_to_moments(const double* d_state, int C, double* d_moments, void* stream);
int ccst_welford_from_moments(const double* d_moments, int C, double* d_state, void* stream);
/* The collective itself for a host without torch.distributed (the reference's processes
 * only meet on disk, README.md:28-37): in-place ncclAllReduce(sum, fp64) of `count` doubles (1 + 2 C moments, plus
 * whatever counters the caller appends) over the caller's communicator `nccl_comm` (an ncclComm_t) on `stream`.
 * NCCL is resolved at run time (the copy already loaded into the process, else libnccl.so.2 on the loader path):
 * the library has no link-time dependency on it.  CCST_ESTATE if no NCCL can be found. */
int ccst_allreduce_moments(void* nccl_comm, double* d_moments, int64_t count, void* stream);

/* ------------------------------------------------------------------------
 * AdaIN
 * ---------------------------------------------------------------------- */

/* adaIN_StyleStat_ContentFeat(content_feat, style_stat) followed by the alpha blend of
 * style_transfer  [function.py:26-33, CCST_OverallStyleTransfer.py:44-45]:
 *   out = alpha * ((x - mu_c)/sigma_c * sigma_s + mu_s) + (1 - alpha) * x
 * with (mu_c, sigma_c) = calc_mean_std(x) computed in the same pass.  d_mu_s/d_sigma_s are
 * [C] when stat_batch_stride == 0 (the reference's [1,C,1,1] broadcast) or [N,C] with
 * stat_batch_stride == C.  alpha = 1 gives the bare operator. */
int ccst_adain_stat_nchw_f32(const float* d_x, int N, int C, int64_t hw, const float* d_mu_s,
                             const float* d_sigma_s, int64_t stat_batch_stride, float alpha,
                             float eps, float* d_out, void* stream);

/* adaptive_instance_normalization(content_feat, style_feat)  [function.py:16-24] (+ alpha blend).
 * style is [N,C,hw_s]; d_scratch must hold 2*N*C floats (style mean/std). */
int ccst_adain_feat_nchw_f32(const float* d_content, const float* d_style, int N, int C,
                             int64_t hw_c, int64_t hw_s, float alpha, float eps, float* d_out,
                             float* d_scratch, void* stream);

/* ------------------------------------------------------------------------
 * Encoder / decoder / style_transfer
 * ---------------------------------------------------------------------- */

ccst_handle* ccst_create(int device);
void ccst_destroy(ccst_handle* h);

/* vgg[:31] weights  [net.py:38-69]: 10 convolutions in order (the 1x1 colour conv first),
 * HOST pointers, fp32, OIHW contiguous, exactly as in the state_dict.  The library folds
 * the 1x1 conv into conv1_1, packs K-major bf16 + fp32 copies and uploads them. Synchronous. */
int ccst_set_encoder_weights(ccst_handle* h, const float* const* h_weights,
                             const float* const* h_biases);
/* decoder weights  [net.py:6-36]: 9 convolutions in order. Synchronous. */
int ccst_set_decoder_weights(ccst_handle* h, const float* const* h_weights,
                             const float* const* h_biases);

/* vgg(content)  [CCST_OverallStyleTransfer.py:35]: [N,3,H,W] -> relu4_1 [N,512,h,w] with
 * h = ceil(ceil(ceil(H/2)/2)/2).  H, W >= 8. */
int ccst_encoder_fwd(ccst_handle* h, const float* d_img, int N, int H, int W, float* d_feat,
                     int precision, void* stream);
/* decoder(feat)  [CCST_OverallStyleTransfer.py:46]: [N,512,fh,fw] -> [N,3,8fh,8fw]. */
int ccst_decoder_fwd(ccst_handle* h, const float* d_feat, int N, int fh, int fw, float* d_img,
                     int precision, void* stream);

/* style_transfer(vgg, decoder, content, style_stat, alpha)  [CCST_OverallStyleTransfer.py:32-46]
 * d_out is [N,3,8h,8w] (== [N,3,H,W] when H and W are multiples of 8). Activations stay on the
 * device in the handle's arena between the layers (NHWC with a reflection halo). */
int ccst_style_transfer(ccst_handle* h, const float* d_img, int N, int H, int W,
                        const float* d_mu_s, const float* d_sigma_s, int64_t stat_batch_stride,
                        float alpha, float* d_out, int precision, void* stream);

/* The same call with the image I/O of the batch loop fused around it:
 *   d_img  [N,H,W,3] uint8 HWC -- the loader's PIL image before `ToTensor`
 *          (cjm_util/data_helper.py:45; torchvision to_tensor: float(u) / 255);
 *   d_out  [N,8h,8w,3] uint8 HWC -- what `save_image(out_img, out_name)` hands to the encoder
 *          (CCST_OverallStyleTransfer.py:167; torchvision utils.save_image:
 *          mul(255).add(0.5).clamp(0,255).to(uint8)), quantised in the last conv's store.
 * 4x fewer bytes each way across PCIe than the fp32 tensors. */
int ccst_style_transfer_u8(ccst_handle* h, const uint8_t* d_img, int N, int H, int W,
                           const float* d_mu_s, const float* d_sigma_s, int64_t stat_batch_stride,
                           float alpha, uint8_t* d_out, int precision, void* stream);

/* The two conversions on their own: ToTensor ([N,H,W,C] uint8 -> [N,C,H,W] fp32 in [0,1]) and
 * save_image's quantisation ([N,C,H,W] fp32 -> [N,H,W,C] uint8). */
int ccst_u8_to_tensor(const uint8_t* d_img_nhwc, int N, int C, int H, int W, float* d_out_nchw,
                      void* stream);
int ccst_quantize_u8(const float* d_img_nchw, int N, int C, int H, int W, uint8_t* d_out_nhwc,
                     void* stream);

/* The loader's `transforms.Resize((S, S))` on the PIL image  [cjm_util/data_helper.py:45-49]: torchvision
 * calls Image.resize(BILINEAR), i.e. Pillow's 8-bit two-pass resample (Pillow is an un-vendored dependency of the
 * reference; restated from src/libImaging/Resample.c of Pillow 12.2.0: 22-bit fixed-point coefficients, horizontal
 * pass into an 8-bit intermediate, then the vertical pass) -- bit-exact.  d_in [N,H,W,C] uint8 -> d_out
 * [N,OH,OW,C] uint8, so that a batch can be uploaded at its ORIGINAL size (PACS: 227x227, 5x fewer bytes than
 * 512x512) and resized on the GPU.  d_scratch: ccst_resize_pil_scratch_bytes(...) bytes of device memory. */
int64_t ccst_resize_pil_scratch_bytes(int N, int H, int W, int C, int OH, int OW);
int ccst_resize_pil_bilinear_u8(const uint8_t* d_in, int N, int H, int W, int C, int OH, int OW, uint8_t* d_out,
                                void* d_scratch, void* stream);

/* `resize = transforms.Resize(args.output_size); output = resize(output)`
 * [CCST_OverallStyleTransfer.py:134-135,154-155] on the device tensor: torchvision's Resize of a float
 * tensor = torch's anti-aliased bilinear interpolation (align_corners = false).  d_in is `planes`
 * contiguous H x W fp32 planes (planes = N*C), d_out `planes` OH x OW planes. */
int ccst_resize_bilinear_aa_f32(const float* d_in, int64_t planes, int H, int W, int OH, int OW,
                                float* d_out, void* stream);

/* Net.encode_with_intermediate + the statistics of calc_style_loss  [net.py:112-136], forward only:
 * encodes d_img and returns calc_mean_std (unbiased variance, sqrt(var + eps)) of relu1_1, relu2_1, relu3_1
 * and relu4_1 -- d_mean[l] / d_std[l] hold N*C_l floats, C_l = 64, 128, 256, 512 (the arrays of 4 pointers
 * are HOST arrays of device pointers) -- taken from the arena while encoding; d_feat (may be NULL) receives
 * relu4_1 as [N,512,h,w] fp32. */
int ccst_encoder_levels(ccst_handle* h, const float* d_img, int N, int H, int W, float* d_feat,
                        float* const* d_mean, float* const* d_std, float eps, int precision, void* stream);

/* nn.MSELoss()(a, b)  [net.py:104, calc_content_loss / calc_style_loss]: d_out[0] = mean((a - b)^2) over n
 * floats; squares in fp32, sums in fp64 in a fixed order (bit-reproducible).  d_scratch: 1024 doubles. */
int ccst_mse_f32(const float* d_a, const float* d_b, int64_t n, double* d_scratch, float* d_out, void* stream);

/* one iteration of the overall-statistics loop: vgg(data) + calc_sum + accumulate
 * [mean_std_computation_effcientMem.py:121-131], encoder output never leaves the arena. */
int ccst_encoder_accumulate(ccst_handle* h, const float* d_img, int N, int H, int W,
                            double* d_state, int precision, void* stream);

/* the same on the loader's uint8 HWC batch [N,H,W,3] (ToTensor on the GPU, a quarter of the upload) */
int ccst_encoder_accumulate_u8(ccst_handle* h, const uint8_t* d_img, int N, int H, int W,
                               double* d_state, int precision, void* stream);

/* f16 operands (CCST_PREC_FP16) store activations with saturation at +-65504 instead of overflowing to
 * inf.  Every epilogue counts the threads whose stores hit the clamp in a device counter owned by the
 * handle; weights whose activations exceed the f16 range (the synthetic and the published VGG weights do
 * not) show up here instead of silently producing a wrong image.  `snapshot` enqueues an asynchronous
 * copy of the counter to *h_count (pinned host memory recommended; valid once `stream` has reached this
 * point), `reset` zeroes it in stream order. */
int ccst_saturation_snapshot(ccst_handle* h, uint32_t* h_count, void* stream);
int ccst_saturation_reset(ccst_handle* h, void* stream);

/* shape helper: relu4_1 spatial size for an input of H x W */
void ccst_feature_hw(int H, int W, int* fh, int* fw);

/* number of kernels this library has launched in this process (monotonic). */
int64_t ccst_launch_count(void);

/* ------------------------------------------------------------------------
 * Diagnostics (used by tests/ and bench.py only)
 * ---------------------------------------------------------------------- */

/* One reflect-pad 3x3 convolution through the selected engine on caller data:
 * d_in  [N,H,W,Cin]  NHWC fp32 (unpadded), h_weight OIHW fp32 [Cout,Cin,3,3], h_bias [Cout].
 * mode 0: d_out [N,H,W,Cout]; mode 1: nearest x2 fused, d_out [N,2H,2W,Cout]; mode 2: 2x2 ceil-mode
 * max-pool fused (requires relu), d_out [N,ceil(H/2),ceil(W/2),Cout]; mode 3: Cout <= 3 (fp32 engine: <= 16),
 * d_out is NCHW [N,Cout,H,W] (the last decoder conv's store); mode 4 (16-bit precisions only): the input is
 * nearest-x2 upsampled BEFORE the reflect-pad conv (net.py:10-11), computed by the phase-decomposed
 * kernel straight from the low-resolution map, d_out [N,2H,2W,Cout].  d_out is fp32. Synchronous. */
int ccst_debug_conv3x3(ccst_handle* h, const float* d_in, int N, int H, int W, int Cin, int Cout,
                       const float* h_weight, const float* h_bias, int relu, int mode,
                       float* d_out, int precision, void* stream);

/* Fusions of the tcgen05 path, all on by default; the tests switch them off one at a time to compare
 * each fused kernel with its un-fused form (bit-exact for the pool, within one 16-bit rounding else). */
#define CCST_FUSE_POOL 1      /* ceil-mode 2x2 max-pool in the producing conv's epilogue            */
#define CCST_FUSE_UPSAMPLE 2  /* nearest x2 upsample folded into the NEXT conv (4 phase convolutions) */
#define CCST_FUSE_STATS 4     /* relu4_1 statistics taken in conv4_1's epilogue                      */
#define CCST_FUSE_ADAIN 8     /* AdaIN folded into dec1's per-image weights / bias (maps >= 2048 px)  */
#define CCST_FUSE_TOTENSOR 16 /* uint8 entry points: conv1_1 reads the uint8 HWC batch itself (W % 16 == 0)      */
#define CCST_FUSE_ALL 31
int ccst_set_fusion(ccst_handle* h, int mask);

/* Per-launch device timing of the encoder/decoder entry points (CUDA events recorded on the
 * caller's stream around every kernel the entry point enqueues).  Off by default. */
int ccst_profile_enable(ccst_handle* h, int on);
/* Synchronises the recorded events of the LAST encoder/decoder/style_transfer call and returns
 * the number of launches; for launch i: ms[i] device time, flops[i] algorithmic FLOPs (convs,
 * else 0), bytes[i] algorithmic HBM bytes, kind[i]: 0 conv_first, 1 tcgen05 conv, 2 ffma conv,
 * 3 pool, 4 adain/stats, 5 layout convert, 6 AdaIN fold (coefficients + per-image dec1 weights).  Arrays hold
 * `max` entries. */
int ccst_profile_read(ccst_handle* h, int max, float* ms, double* flops, double* bytes, int* kind);

#ifdef __cplusplus
}
#endif
#endif /* CCST_B200_H_ */
