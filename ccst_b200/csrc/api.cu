// C ABI of libccst_b200.so: handle, weight packing and the encoder / AdaIN / decoder pipelines.
//
// Pipelines follow vgg[:31] (net.py:38-69), decoder (net.py:6-36) and style_transfer
// (CCST_OverallStyleTransfer.py:32-46).  Activations never leave the device arena between layers;
// the only NCHW fp32 tensors are the caller's image / feature / output tensors.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <type_traits>
#include <utility>
#include <vector>

#include "layers.h"

namespace ccst {

// ------------------------------------------------------------------ globals
static thread_local char g_err[512] = "";
int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_sm_count[64];
static int g_cc_major[64];
static bool g_dev_known[64];

static int query_device(int dev) {
  if (dev < 0 || dev >= 64) return CCST_EINVAL;
  if (!g_dev_known[dev]) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
      cudaGetLastError();
      set_error("no usable CUDA device %d (libccst_b200 has no CPU fallback)", dev);
      return CCST_ECUDA;
    }
    g_sm_count[dev] = prop.multiProcessorCount;
    g_cc_major[dev] = prop.major;
    g_dev_known[dev] = true;
  }
  return CCST_OK;
}

int require_sm100() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available (libccst_b200 has no CPU fallback)");
    return CCST_ECUDA;
  }
  if (int e = query_device(dev)) return e;
  if (g_cc_major[dev] != 10) {
    set_error("device %d has compute capability %d.x; libccst_b200 is built for sm_100a only", dev,
              g_cc_major[dev]);
    return CCST_EARCH;
  }
  return CCST_OK;
}

cudaError_t ensure_dyn_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({kernel, dev})) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.insert({kernel, dev});
  return e;
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && g_dev_known[dev]) return g_sm_count[dev];
  return 148;
}

}  // namespace ccst

using namespace ccst;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ handle
namespace {

struct ConvLayer {
  int cin = 0, cout = 0;
  int pad64 = 0;     // Cout rounded up to 64 (ffma weights / bias)
  int pad_umma = 0;  // Cout, or 16 for the 3-channel last conv
  float* w_ffma = nullptr;  // [9][cin][pad64]
  bf16* w_umma = nullptr;   // [pad_umma][9*cin] bf16
  __half* w_umma_h = nullptr;  // same, fp16
  bf16* w_sm = nullptr;     // cout == 64: [192 = s*64 + co][3*cin = r*cin + c] (s-merged kernel)
  __half* w_sm_h = nullptr;
  bf16* w_up = nullptr;     // layer after a nearest-x2 upsample: phase weights [4][cout][4*cin] (EPI_UPS)
  __half* w_up_h = nullptr;
  float* bias = nullptr;    // [pad64]
  // first decoder conv only (AdaIN folded into it): fp32 K-major master weights [cout][9*cin] and the
  // per-(cout, cin) sums over the 9 taps
  float* w_k32 = nullptr;
  float* w_tapsum = nullptr;
  // f16x3 engine (encoder layers): [cout/64][hi | lo][64][9*cin] f16 halves of w * 2^e; the epilogue multiplies
  // the accumulator by x3_scale = 2^-e
  __half* w_x3 = nullptr;
  bf16* w_x3_b = nullptr;  // the same split with bf16 halves (bf16x3 engine)
  // x3 engines, layer after an upsample: phase weights (w_up) split the same way, [4 phases][cout/64][hi | lo][64][4*cin]
  __half* w_up_x3 = nullptr;
  bf16* w_up_x3_b = nullptr;
  float x3_scale = 1.f;
};

constexpr int kEncLayers = 8;  // conv1_2 .. conv4_1 (conv1_1 is the fused first conv)
constexpr int kDecLayers = 9;
constexpr int kMaxProf = 48;

struct ProfSlot {
  cudaEvent_t a = nullptr, b = nullptr;
  double flops = 0, bytes = 0;
  int kind = 0;
};

}  // namespace

struct ccst_handle {
  int device = 0;
  bool enc_ready = false, dec_ready = false;
  float* first_w27 = nullptr;   // folded conv1_1 weights [27][64] fp32 (FFMA path)
  float* first_b64 = nullptr;
  bf16* first_wk_b = nullptr;   // same, [64][32] K-major bf16 / f16 (tcgen05 path)
  __half* first_wk_h = nullptr;
  bf16* first_x3_b = nullptr;   // x3 engines: [128 = hi | lo][64 = 2 x 32 k] halves of w * 2^e (conv_first_x3_ws_kernel)
  __half* first_x3_h = nullptr;
  float first_x3_scale = 1.f;   // 2^-e
  ConvLayer enc[kEncLayers];
  ConvLayer dec[kDecLayers];
  void* arena[2] = {nullptr, nullptr};
  size_t arena_bytes[2] = {0, 0};
  float2* raw = nullptr;
  size_t raw_elems = 0;
  float* io_f32 = nullptr;  // uint8 I/O: ToTensor'd input batch / fp32-engine output before quantisation
  size_t io_elems = 0;
  bool fuse_pool = true;
  bool fuse_up = true;  // nearest-x2 upsample folded into the NEXT conv (EPI_UPS) instead of the store
  bool fuse_stats = true;  // relu4_1 statistics taken in conv4_1's epilogue (EPI_ACT_STATS)
  bool fuse_adain = true;  // AdaIN folded into dec1's weights / bias per image (maps >= kFoldMinHW pixels)
  bool fuse_totensor = true;  // uint8 entry points: conv1_1's window loader reads the uint8 HWC batch itself
  void* w_fold = nullptr;  // per-image dec1 weights [N][256][9*512] (16-bit) of the folded AdaIN
  size_t w_fold_bytes = 0;
  float* b_fold = nullptr; // per-image dec1 bias [N][256]
  size_t b_fold_elems = 0;
  unsigned int* sat_count = nullptr;  // f16 stores that hit the +-65504 clamp since the last reset
  bool split = false;  // the call in flight runs the f16x3 engine (set by check_common from the precision code)
  bool profiling = false;
  int prof_n = 0;
  ProfSlot prof[kMaxProf];
};

namespace {

struct ProfScope {
  ccst_handle* h;
  cudaStream_t st;
  int slot = -1;
  ProfScope(ccst_handle* h_, cudaStream_t st_, int kind, double flops, double bytes) : h(h_), st(st_) {
    if (!h->profiling || h->prof_n >= kMaxProf) return;
    slot = h->prof_n++;
    ProfSlot& s = h->prof[slot];
    if (!s.a) cudaEventCreate(&s.a), cudaEventCreate(&s.b);
    s.kind = kind, s.flops = flops, s.bytes = bytes;
    cudaEventRecord(s.a, st);
  }
  ~ProfScope() {
    if (slot >= 0) cudaEventRecord(h->prof[slot].b, st);
  }
};

void free_layer(ConvLayer& L) {
  cudaFree(L.w_ffma);
  cudaFree(L.w_umma);
  cudaFree(L.w_umma_h);
  cudaFree(L.w_sm);
  cudaFree(L.w_sm_h);
  cudaFree(L.w_up);
  cudaFree(L.w_up_h);
  cudaFree(L.bias);
  cudaFree(L.w_k32);
  cudaFree(L.w_tapsum);
  cudaFree(L.w_x3);
  cudaFree(L.w_x3_b);
  cudaFree(L.w_up_x3);
  cudaFree(L.w_up_x3_b);
  L = ConvLayer();
}

// OIHW fp32 host weights -> device packs
int pack_layer(ConvLayer& L, int cin, int cout, const float* w, const float* b, bool up_before = false,
               bool fold_src = false, bool split_pack = false) {
  free_layer(L);
  L.cin = cin, L.cout = cout;
  L.pad64 = (cout + 63) / 64 * 64;
  L.pad_umma = cout % 64 == 0 ? cout : 16;
  CCST_CHECK_ARG(cout % 64 == 0 || cout <= 16, "unsupported Cout=%d", cout);
  const int K = 9 * cin;
  std::vector<float> wf((size_t)K * L.pad64, 0.f);
  std::vector<bf16> wu((size_t)L.pad_umma * K);
  std::vector<__half> wh((size_t)L.pad_umma * K);
  std::vector<float> bp(L.pad64, 0.f);
  for (size_t i = 0; i < wu.size(); ++i) wu[i] = __float2bfloat16(0.f), wh[i] = __float2half(0.f);
  for (int o = 0; o < cout; ++o) {
    bp[o] = b[o];
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < 9; ++t) {
        const float v = w[((size_t)o * cin + c) * 9 + t];
        wf[((size_t)t * cin + c) * L.pad64 + o] = v;
        wu[(size_t)o * K + (size_t)t * cin + c] = __float2bfloat16(v);
        wh[(size_t)o * K + (size_t)t * cin + c] = __float2half(v);
      }
  }
  CCST_CUDA(cudaMalloc(&L.w_ffma, wf.size() * sizeof(float)));
  CCST_CUDA(cudaMalloc(&L.w_umma, wu.size() * sizeof(bf16)));
  CCST_CUDA(cudaMalloc(&L.w_umma_h, wh.size() * sizeof(__half)));
  CCST_CUDA(cudaMalloc(&L.bias, bp.size() * sizeof(float)));
  CCST_CUDA(cudaMemcpy(L.w_umma_h, wh.data(), wh.size() * sizeof(__half), cudaMemcpyHostToDevice));
  CCST_CUDA(cudaMemcpy(L.w_ffma, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice));
  CCST_CUDA(cudaMemcpy(L.w_umma, wu.data(), wu.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  CCST_CUDA(cudaMemcpy(L.bias, bp.data(), bp.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (split_pack) {
    // split weights for the x3 engines: w * 2^e = hi + lo with hi = T16(w * 2^e), lo = T16(w * 2^e - hi); e puts
    // the largest |w| just below 2^10, so the low parts (2^-11 relative) of all but vanishing weights are normal
    // f16 numbers and the product of the scale with any activation stays far inside the fp32 range.
    // Rows [tile of 64 output channels][hi | lo][co], K = 9 * cin: the two halves of a tile are the N = 128
    // rows of one MMA (conv_x3.cuh); a layer with fewer than 64 output channels is zero-padded to one tile.
    float wmax = 0.f;
    for (size_t i = 0; i < (size_t)cout * cin * 9; ++i) wmax = fmaxf(wmax, fabsf(w[i]));
    int e = 0;
    if (wmax > 0.f && std::isfinite(wmax)) {
      int ex;
      frexpf(wmax, &ex);  // wmax = m * 2^ex, m in [0.5, 1)
      e = 10 - ex;
    }
    const int tiles = (cout + 63) / 64;
    std::vector<__half> wx((size_t)tiles * 128 * K, __float2half(0.f));
    std::vector<bf16> wxb((size_t)tiles * 128 * K, __float2bfloat16(0.f));
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int t = 0; t < 9; ++t) {
          const float v = ldexpf(w[((size_t)o * cin + c) * 9 + t], e);
          const size_t row_hi = (size_t)(o / 64) * 128 + (o % 64), row_lo = row_hi + 64;
          const size_t k = (size_t)t * cin + c;
          const __half hi = __float2half(v);
          wx[row_hi * K + k] = hi;
          wx[row_lo * K + k] = __float2half(v - __half2float(hi));
          const bf16 hb = __float2bfloat16(v);
          wxb[row_hi * K + k] = hb;
          wxb[row_lo * K + k] = __float2bfloat16(v - __bfloat162float(hb));
        }
    CCST_CUDA(cudaMalloc(&L.w_x3, wx.size() * sizeof(__half)));
    CCST_CUDA(cudaMemcpy(L.w_x3, wx.data(), wx.size() * sizeof(__half), cudaMemcpyHostToDevice));
    CCST_CUDA(cudaMalloc(&L.w_x3_b, wxb.size() * sizeof(bf16)));
    CCST_CUDA(cudaMemcpy(L.w_x3_b, wxb.data(), wxb.size() * sizeof(bf16), cudaMemcpyHostToDevice));
    L.x3_scale = ldexpf(1.f, -e);
  }
  if (fold_src) {
    std::vector<float> wk((size_t)cout * K), ws((size_t)cout * cin);
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c) {
        double acc = 0;
        for (int t = 0; t < 9; ++t) {
          const float v = w[((size_t)o * cin + c) * 9 + t];
          wk[(size_t)o * K + (size_t)t * cin + c] = v;
          acc += (double)v;
        }
        ws[(size_t)o * cin + c] = (float)acc;
      }
    CCST_CUDA(cudaMalloc(&L.w_k32, wk.size() * sizeof(float)));
    CCST_CUDA(cudaMalloc(&L.w_tapsum, ws.size() * sizeof(float)));
    CCST_CUDA(cudaMemcpy(L.w_k32, wk.data(), wk.size() * sizeof(float), cudaMemcpyHostToDevice));
    CCST_CUDA(cudaMemcpy(L.w_tapsum, ws.data(), ws.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (cout == 64) {
    const int K3 = 3 * cin;
    std::vector<bf16> sb((size_t)192 * K3);
    std::vector<__half> sh((size_t)192 * K3);
    for (int sc = 0; sc < 3; ++sc)
      for (int o = 0; o < 64; ++o)
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < cin; ++c) {
            const float v = w[((size_t)o * cin + c) * 9 + r * 3 + sc];
            const size_t idx = (size_t)(sc * 64 + o) * K3 + (size_t)r * cin + c;
            sb[idx] = __float2bfloat16(v);
            sh[idx] = __float2half(v);
          }
    CCST_CUDA(cudaMalloc(&L.w_sm, sb.size() * sizeof(bf16)));
    CCST_CUDA(cudaMalloc(&L.w_sm_h, sh.size() * sizeof(__half)));
    CCST_CUDA(cudaMemcpy(L.w_sm, sb.data(), sb.size() * sizeof(bf16), cudaMemcpyHostToDevice));
    CCST_CUDA(cudaMemcpy(L.w_sm_h, sh.data(), sh.size() * sizeof(__half), cudaMemcpyHostToDevice));
  }
  if (up_before && cout % 64 == 0) {
    // conv3x3(reflect_pad(nearest_up2(S))) as four 2x2 phase convolutions over S (net.py:10-11,
    // 23-24, 30-31).  Output pixel (2y + a, 2x + b) reads U[2y + a + r - 1] = S[(2y + a + r - 1) >> 1]:
    //   a = 0: r = 0 -> S[y - 1], r = 1, 2 -> S[y];   a = 1: r = 0, 1 -> S[y], r = 2 -> S[y + 1]
    // (same in x), so the taps that land on the same source pixel are summed once here, in double,
    // and rounded once to the operand type.  Reflection of U at the border = clamping S.
    const int K4 = 4 * cin;
    std::vector<bf16> ub((size_t)4 * cout * K4);
    std::vector<__half> uh((size_t)4 * cout * K4);
    static const int lo[2][2] = {{0, 1}, {0, 2}}, hi[2][2] = {{0, 2}, {1, 2}};  // [phase][d] -> taps lo..hi
    for (int a = 0; a < 2; ++a)
      for (int bb = 0; bb < 2; ++bb)
        for (int o = 0; o < cout; ++o)
          for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx)
              for (int c = 0; c < cin; ++c) {
                double acc = 0;
                for (int r = lo[a][dy]; r <= hi[a][dy]; ++r)
                  for (int sc = lo[bb][dx]; sc <= hi[bb][dx]; ++sc)
                    acc += (double)w[((size_t)o * cin + c) * 9 + r * 3 + sc];
                const size_t idx = ((size_t)(a * 2 + bb) * cout + o) * K4 + (size_t)(dy * 2 + dx) * cin + c;
                ub[idx] = __float2bfloat16((float)acc);
                uh[idx] = __float2half((float)acc);
              }
    CCST_CUDA(cudaMalloc(&L.w_up, ub.size() * sizeof(bf16)));
    CCST_CUDA(cudaMalloc(&L.w_up_h, uh.size() * sizeof(__half)));
    CCST_CUDA(cudaMemcpy(L.w_up, ub.data(), ub.size() * sizeof(bf16), cudaMemcpyHostToDevice));
    CCST_CUDA(cudaMemcpy(L.w_up_h, uh.data(), uh.size() * sizeof(__half), cudaMemcpyHostToDevice));
    if (split_pack) {
      // x3 engines: the same phase sums (in double) times the layer's 2^e, split into hi + lo; rows
      // [phase][tile of 64 output channels][hi | lo][co].  A phase weight sums at most 4 taps: |w| 2^e < 2^12.
      int ex;
      frexpf(L.x3_scale, &ex);  // x3_scale = 2^-e = 0.5 * 2^ex
      const int e = 1 - ex;
      const int tiles = cout / 64;
      std::vector<__half> xh((size_t)4 * tiles * 128 * K4);
      std::vector<bf16> xb((size_t)4 * tiles * 128 * K4);
      for (int a = 0; a < 2; ++a)
        for (int bb = 0; bb < 2; ++bb)
          for (int o = 0; o < cout; ++o)
            for (int dy = 0; dy < 2; ++dy)
              for (int dx = 0; dx < 2; ++dx)
                for (int c = 0; c < cin; ++c) {
                  double acc = 0;
                  for (int r = lo[a][dy]; r <= hi[a][dy]; ++r)
                    for (int sc = lo[bb][dx]; sc <= hi[bb][dx]; ++sc)
                      acc += (double)w[((size_t)o * cin + c) * 9 + r * 3 + sc];
                  const float v = (float)ldexp(acc, e);
                  const size_t row_hi = ((size_t)(a * 2 + bb) * tiles + o / 64) * 128 + (o % 64), row_lo = row_hi + 64;
                  const size_t k = (size_t)(dy * 2 + dx) * cin + c;
                  const __half hh = __float2half(v);
                  xh[row_hi * K4 + k] = hh;
                  xh[row_lo * K4 + k] = __float2half(v - __half2float(hh));
                  const bf16 hb = __float2bfloat16(v);
                  xb[row_hi * K4 + k] = hb;
                  xb[row_lo * K4 + k] = __float2bfloat16(v - __bfloat162float(hb));
                }
      CCST_CUDA(cudaMalloc(&L.w_up_x3, xh.size() * sizeof(__half)));
      CCST_CUDA(cudaMalloc(&L.w_up_x3_b, xb.size() * sizeof(bf16)));
      CCST_CUDA(cudaMemcpy(L.w_up_x3, xh.data(), xh.size() * sizeof(__half), cudaMemcpyHostToDevice));
      CCST_CUDA(cudaMemcpy(L.w_up_x3_b, xb.data(), xb.size() * sizeof(bf16), cudaMemcpyHostToDevice));
    }
  }
  return CCST_OK;
}

const int kEncCh[kEncLayers][2] = {{64, 64},   {64, 128},  {128, 128}, {128, 256},
                                   {256, 256}, {256, 256}, {256, 256}, {256, 512}};
const bool kEncPoolAfter[kEncLayers] = {true, false, true, false, false, false, true, false};
const int kDecCh[kDecLayers][2] = {{512, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 128},
                                   {128, 128}, {128, 64},  {64, 64},   {64, 3}};
const bool kDecUpAfter[kDecLayers] = {true, false, false, false, true, false, true, false, false};

size_t act_bytes(int N, int H, int W, int C, size_t esz) {
  return (size_t)N * (H + 2) * (W + 2) * C * esz;
}

int ensure_arena(ccst_handle* h, size_t bytes) {
  for (int i = 0; i < 2; ++i) {
    if (h->arena_bytes[i] >= bytes) continue;
    if (h->arena[i]) CCST_CUDA(cudaFree(h->arena[i]));
    h->arena[i] = nullptr, h->arena_bytes[i] = 0;
    CCST_CUDA(cudaMalloc(&h->arena[i], bytes));
    h->arena_bytes[i] = bytes;
  }
  return CCST_OK;
}

int ensure_raw(ccst_handle* h, size_t elems) {
  if (h->raw_elems >= elems) return CCST_OK;
  if (h->raw) CCST_CUDA(cudaFree(h->raw));
  h->raw = nullptr, h->raw_elems = 0;
  CCST_CUDA(cudaMalloc(&h->raw, elems * sizeof(float2)));
  h->raw_elems = elems;
  return CCST_OK;
}

// largest activation of encoder (from an HxW image) and decoder (from an fh x fw feature map)
size_t plan_bytes(int N, int H, int W, int fh, int fw, size_t esz, bool enc, bool dec) {
  size_t m = 0;
  auto upd = [&](int hh, int ww, int c) {
    size_t b = act_bytes(N, hh, ww, c, esz);
    if (b > m) m = b;
  };
  if (enc) {
    int hh = H, ww = W;
    upd(hh, ww, 64);
    for (int i = 0; i < kEncLayers; ++i) {
      upd(hh, ww, kEncCh[i][1]);
      if (kEncPoolAfter[i]) {
        hh = (hh + 1) / 2, ww = (ww + 1) / 2;
        upd(hh, ww, kEncCh[i][1]);
      }
    }
  }
  if (dec) {
    int hh = fh, ww = fw;
    upd(hh, ww, 512);
    for (int i = 0; i < kDecLayers - 1; ++i) {
      if (kDecUpAfter[i]) hh *= 2, ww *= 2;
      upd(hh, ww, kDecCh[i][1]);
    }
  }
  return m;
}

// maps with at least this many pixels fold AdaIN into dec1 (per-image weights: 2.4 MB written per
// image); below it the affine pass over the feature map (2 KiB per pixel) is cheaper
constexpr int kFoldMinHW = 2048;

int ensure_fold(ccst_handle* h, size_t w_bytes, size_t b_elems) {
  if (h->w_fold_bytes < w_bytes) {
    if (h->w_fold) CCST_CUDA(cudaFree(h->w_fold));
    h->w_fold = nullptr, h->w_fold_bytes = 0;
    CCST_CUDA(cudaMalloc(&h->w_fold, w_bytes));
    h->w_fold_bytes = w_bytes;
  }
  if (h->b_fold_elems < b_elems) {
    if (h->b_fold) CCST_CUDA(cudaFree(h->b_fold));
    h->b_fold = nullptr, h->b_fold_elems = 0;
    CCST_CUDA(cudaMalloc(&h->b_fold, b_elems * sizeof(float)));
    h->b_fold_elems = b_elems;
  }
  return CCST_OK;
}

template <typename T>
struct Weights16;  // the operand-type view of a layer's 16-bit packs
template <>
struct Weights16<bf16> {
  static const bf16* wk(const ConvLayer& L) { return L.w_umma; }
  static const bf16* sm(const ConvLayer& L) { return L.w_sm; }
  static const bf16* up(const ConvLayer& L) { return L.w_up; }
  static const bf16* first(const ccst_handle* h) { return h->first_wk_b; }
  static const bf16* first_x3(const ccst_handle* h) { return h->first_x3_b; }
  static const bf16* x3(const ConvLayer& L) { return L.w_x3_b; }
  static const bf16* x3_up(const ConvLayer& L) { return L.w_up_x3_b; }
};
template <>
struct Weights16<__half> {
  static const __half* wk(const ConvLayer& L) { return L.w_umma_h; }
  static const __half* sm(const ConvLayer& L) { return L.w_sm_h; }
  static const __half* up(const ConvLayer& L) { return L.w_up_h; }
  static const __half* first(const ccst_handle* h) { return h->first_wk_h; }
  static const __half* first_x3(const ccst_handle* h) { return h->first_x3_h; }
  static const __half* x3(const ConvLayer& L) { return L.w_x3; }
  static const __half* x3_up(const ConvLayer& L) { return L.w_up_x3; }
};

template <typename T>
struct Pipe {
  ccst_handle* h;
  cudaStream_t st;
  int cur_slot = 0;
  ActView<T> cur;
  bool up_pending = false;  // `cur` is a low-resolution map (replicate halo) awaiting its x2 upsample
  uint8_t* out_u8 = nullptr;  // decoder(): store NHWC uint8 (save_image quantisation) instead of NCHW fp32
  bool stats_in_tiles = false;  // the last encoder conv left tile statistics of `cur` in h->raw
  bool fold_pending = false;  // AdaIN of `cur` lives in h->w_fold / h->b_fold: the next conv applies it
  // per-(n, c) mean / std (calc_mean_std) of relu1_1, relu2_1, relu3_1, relu4_1 while encoding
  // (Net.encode_with_intermediate + calc_style_loss, net.py:112-136); NULL = not wanted
  float* const* lvl_mean = nullptr;
  float* const* lvl_std = nullptr;
  float lvl_eps = 1e-5f;
  // the batch as the loader holds it (uint8 HWC): conv1_1 of the tcgen05 engines reads it directly, `img` is unused
  const uint8_t* img_u8 = nullptr;

  // x3 engines: every map holds the [hi | lo] halves of its C logical channels (v.C = 2 * C)
  bool split() const { return sizeof(T) == 2 && h->split; }
  ActView<T> view(int slot, int N, int H, int W, int C) {
    ActView<T> v;
    v.p = reinterpret_cast<T*>(h->arena[slot]);
    v.N = N, v.H = H, v.W = W, v.C = split() ? 2 * C : C;
    return v;
  }

  int conv(const ConvLayer& L, int relu, int epi, ActView<T> out, float* out_nchw, int halo_edge = 1,
           float2* tile_stats = nullptr, bool per_sample = false);

  int first_launch(const float* img, int N, int H, int W);
  int first(const float* img, int N, int H, int W) {
    cur_slot = 0;
    cur = view(0, N, H, W, 64);
    ProfScope ps(h, st, 0, 2.0 * 27 * 64 * (double)N * H * W,
                 (double)N * H * W * (12.0 + 64.0 * sizeof(T)));
    return first_launch(img, N, H, W);
  }

  // conv (+ fused or separate pool / fused upsample); result becomes `cur`
  int step(const ConvLayer& L, bool pool_after, bool up_after, bool want_stats = false) {
    const int N = cur.N, H = cur.H, W = cur.W;
    const bool fused_pool = pool_after && (h->fuse_pool || split()) && sizeof(T) == 2;
    // tcgen05 path: a layer followed by `Upsample` stores its low-resolution output with a replicate
    // halo and the NEXT conv consumes it through the phase-decomposed kernel (EPI_UPS): 16 instead of
    // 36 tap-GEMMs per source pixel and no 4x-replicated activation in HBM.
    const bool defer_up = up_after && h->fuse_up && sizeof(T) == 2 && !pool_after;
    const bool ups = up_pending;
    CCST_CHECK_ARG(!ups || (!pool_after && !up_after && L.w_up != nullptr),
                   "upsample-fused conv cannot pool/upsample itself and needs phase weights");
    int oh = H, ow = W, epi = EPI_ACT, halo_edge = 1;
    if (ups) oh = 2 * H, ow = 2 * W, epi = EPI_UPS;
    else if (defer_up) halo_edge = 0;
    else if (up_after) oh = 2 * H, ow = 2 * W, epi = EPI_ACT_UP2;
    if (fused_pool) oh = (H + 1) / 2, ow = (W + 1) / 2, epi = EPI_ACT_POOL;
    // relu4_1 statistics in the epilogue of the conv that produces it (tcgen05 path, N = 256 tiles)
    float2* tile_stats = nullptr;
    if (want_stats && h->fuse_stats && sizeof(T) == 2 && !split() && epi == EPI_ACT && halo_edge == 1 &&
        L.cout % 256 == 0 && N <= 65535) {
      if (int e = ensure_raw(h, nhwc_tile_scratch_elems(N, L.cout, H, W))) return e;
      tile_stats = h->raw + 2 * (size_t)N * L.cout;
      epi = EPI_ACT_STATS;
    }
    const bool per_sample = fold_pending;
    CCST_CHECK_ARG(!per_sample || epi == EPI_ACT, "folded AdaIN needs the plain epilogue");
    ActView<T> out = view(cur_slot ^ 1, N, oh, ow, L.cout);
    {
      // executed FLOPs: the phase form runs 4 phases x 4 taps per SOURCE pixel (= 4 taps per output
      // pixel) where the plain form runs 9 taps per output pixel
      const double flops = (ups ? 2.0 * 16 * L.cin * L.cout * (double)N * H * W
                                : 2.0 * 9 * L.cin * L.cout * (double)N * H * W) * (split() ? 4.0 : 1.0);
      const double bytes = (double)cur.elems() * sizeof(T) + (double)out.elems() * sizeof(T);
      ProfScope ps(h, st, sizeof(T) == 2 ? 1 : 2, flops, bytes);
      if (int e = conv(L, 1, epi, out, nullptr, halo_edge, tile_stats, per_sample)) return e;
    }
    cur = out, cur_slot ^= 1;
    up_pending = defer_up;
    fold_pending = false;
    stats_in_tiles = tile_stats != nullptr;
    if (pool_after && !fused_pool) {
      ActView<T> po = view(cur_slot ^ 1, N, (H + 1) / 2, (W + 1) / 2, L.cout);
      ProfScope ps(h, st, 3, 0, (double)(cur.elems() + po.elems()) * sizeof(T));
      if (int e = launch_pool<T>(cur, po, st)) return e;
      cur = po, cur_slot ^= 1;
    }
    return CCST_OK;
  }

  // statistics of `cur` (one of the four relu*_1 maps) for the style losses
  int level_stats(int level) {
    if (!lvl_mean) return CCST_OK;
    const int C = split() ? cur.C / 2 : cur.C;
    const int NC = cur.N * C, HW = cur.H * cur.W;
    ProfScope ps(h, st, 4, 0, (double)cur.elems() * sizeof(T));
    if (int e = ensure_raw(h, nhwc_scratch_elems(cur.N, C, HW))) return e;
    if (int e = stats_launch()) return e;
    return launch_raw_to_mean_std(h->raw, NC, HW, lvl_eps, 1, lvl_mean[level], lvl_std[level], st);
  }

  int encoder(const float* img, int N, int H, int W) {
    if (int e = first(img, N, H, W)) return e;
    if (int e = level_stats(0)) return e;  // relu1_1
    for (int i = 0; i < kEncLayers; ++i) {
      // with level statistics on, relu4_1's come from the same pass as the other levels
      if (int e = step(h->enc[i], kEncPoolAfter[i], false, i == kEncLayers - 1 && !lvl_mean)) return e;
      if (i == 1) if (int e = level_stats(1)) return e;  // relu2_1 = conv2_1
      if (i == 3) if (int e = level_stats(2)) return e;  // relu3_1 = conv3_1
      if (i == 7) if (int e = level_stats(3)) return e;  // relu4_1 = conv4_1
    }
    return CCST_OK;
  }

  // AdaIN (+ alpha blend) of `cur` = relu4_1.  tcgen05 path with tile statistics and a map of at least
  // kFoldMinHW pixels: nothing touches the feature map -- the per-(n, c) affine goes into per-image
  // copies of dec1's weights and bias (conv(W, x*A + B') = conv(W*diag(A), x) + tapsum(W) . B', exact
  // under reflection padding) and the decoder's first conv reads relu4_1 as conv4_1 left it.
  int adain(const float* mu_s, const float* sigma_s, int64_t stride, float alpha) {
    if (try_fold(mu_s, sigma_s, stride, alpha)) return fold_rc;
    ActView<T> out = view(cur_slot ^ 1, cur.N, cur.H, cur.W, split() ? cur.C / 2 : cur.C);
    ProfScope ps(h, st, 4, 0, 2.0 * (double)cur.N * cur.H * cur.W * cur.C * sizeof(T));
    if (int e = adain_launch(out, mu_s, sigma_s, stride, alpha)) return e;
    cur = out, cur_slot ^= 1;
    stats_in_tiles = false;
    return CCST_OK;
  }

  // {mean, M2} per (n, c) of `cur` -> h->raw[0 .. N*C) (scratch already sized)
  int stats_launch();
  int to_nchw(float* d_feat);
  int from_nchw(const float* d_feat) {
    if constexpr (sizeof(T) == 2) {
      if (split()) return launch_nchw_to_act_split<T>(d_feat, cur, st);
    }
    return launch_nchw_to_act<T>(d_feat, cur, st);
  }

  int fold_rc = CCST_OK;
  bool try_fold(const float* mu_s, const float* sigma_s, int64_t stride, float alpha);
  int adain_launch(ActView<T> out, const float* mu_s, const float* sigma_s, int64_t stride, float alpha);

  int decoder(float* out_nchw) {
    for (int i = 0; i < kDecLayers - 1; ++i)
      if (int e = step(h->dec[i], false, kDecUpAfter[i])) return e;
    const ConvLayer& L = h->dec[kDecLayers - 1];
    const double flops = 2.0 * 9 * L.cin * L.cout * (double)cur.N * cur.H * cur.W;
    const double bytes =
        (double)cur.elems() * sizeof(T) + (double)cur.N * cur.H * cur.W * L.cout * 4.0;
    ProfScope ps(h, st, sizeof(T) == 2 ? 1 : 2, flops, bytes);
    return conv(L, 0, EPI_NCHW_F32, cur /*unused*/, out_nchw);
  }
};

template <>
bool Pipe<float>::try_fold(const float*, const float*, int64_t, float) {
  return false;
}
template <typename T>
bool Pipe<T>::try_fold(const float* mu_s, const float* sigma_s, int64_t stride, float alpha) {
  const ConvLayer& L = h->dec[0];
  const size_t w_bytes = (size_t)cur.N * L.cout * 9 * L.cin * sizeof(T);
  if (!(stats_in_tiles && !split() && h->fuse_adain && h->fuse_up && L.w_k32 && cur.C == L.cin && L.cout == 256 &&
        cur.H * cur.W >= kFoldMinHW && w_bytes <= ((size_t)1 << 30)))
    return false;
  fold_rc = CCST_OK;
  ProfScope ps(h, st, 6, 0, (double)w_bytes + (double)L.cout * 9 * L.cin * 4.0);
  if ((fold_rc = ensure_fold(h, w_bytes, (size_t)cur.N * L.cout))) return true;
  fold_rc = launch_adain_fold<T>(cur.N, cur.C, cur.H, cur.W, L.cout, h->raw, mu_s, sigma_s, stride, alpha, 1e-5f,
                                 L.w_k32, L.w_tapsum, L.bias, reinterpret_cast<T*>(h->w_fold), h->b_fold,
                                 h->sat_count, st);
  if (fold_rc == CCST_OK) fold_pending = true, stats_in_tiles = false;
  return true;
}

template <>
int Pipe<float>::adain_launch(ActView<float> out, const float* mu_s, const float* sigma_s, int64_t stride,
                              float alpha) {
  if (int e = ensure_raw(h, nhwc_scratch_elems(cur.N, cur.C, cur.H * cur.W))) return e;
  return launch_adain_nhwc<float>(cur, out, mu_s, sigma_s, stride, alpha, 1e-5f, h->raw, st);
}
template <typename T>
int Pipe<T>::adain_launch(ActView<T> out, const float* mu_s, const float* sigma_s, int64_t stride,
                          float alpha) {
  if (split()) {
    if (int e = ensure_raw(h, nhwc_scratch_elems(cur.N, cur.C / 2, cur.H * cur.W))) return e;
    return launch_adain_nhwc_split<T>(cur, out, mu_s, sigma_s, stride, alpha, 1e-5f, h->raw, st);
  }
  if (stats_in_tiles)  // statistics already taken by the producing conv's epilogue
    return launch_adain_nhwc_tiles<T>(cur, out, mu_s, sigma_s, stride, alpha, 1e-5f, h->raw, st);
  if (int e = ensure_raw(h, nhwc_scratch_elems(cur.N, cur.C, cur.H * cur.W))) return e;
  return launch_adain_nhwc<T>(cur, out, mu_s, sigma_s, stride, alpha, 1e-5f, h->raw, st);
}

template <typename T>
int Pipe<T>::stats_launch() {
  if constexpr (sizeof(T) == 2) {
    if (split()) return launch_stats_nhwc_split<T>(cur, h->raw, st);
  }
  return launch_stats_nhwc<T>(cur, h->raw, st);
}
template <typename T>
int Pipe<T>::to_nchw(float* d_feat) {
  if constexpr (sizeof(T) == 2) {
    if (split()) return launch_act_to_nchw_split<T>(cur, d_feat, st);
  }
  return launch_act_to_nchw<T>(cur, d_feat, st);
}

template <>
int Pipe<float>::first_launch(const float* img, int N, int H, int W) {
  return launch_conv_first<float>(img, N, H, W, h->first_w27, h->first_b64, cur, st);
}
template <typename T>
int Pipe<T>::first_launch(const float* img, int N, int H, int W) {
  if (split()) {
    // x3 engines: split im2col rows on the tensor pipe when TMA can fetch the image rows, else the fp32
    // CUDA-core kernel; both store [hi | lo]
    if (img_u8 || (W % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0))
      return launch_conv_first_x3<T>(img, N, H, W, Weights16<T>::first_x3(h), h->first_x3_scale, h->first_b64, cur, st,
                                     h->sat_count, img_u8);
    return launch_conv_first_split<T>(img, N, H, W, h->first_w27, h->first_b64, cur, st, h->sat_count);
  }
  return launch_conv_first_umma<T>(img, N, H, W, Weights16<T>::first(h), h->first_b64, cur, st, h->sat_count, img_u8);
}
template <>
int Pipe<float>::conv(const ConvLayer& L, int relu, int epi, ActView<float> out, float* out_nchw, int,
                      float2*, bool) {
  return launch_conv_ffma(cur, L.w_ffma, L.bias, L.cout, L.pad64, relu, epi, out, out_nchw, st);
}
template <typename T>
int Pipe<T>::conv(const ConvLayer& L, int relu, int epi, ActView<T> out, float* out_nchw, int halo_edge,
                  float2* tile_stats, bool per_sample) {
  UmmaConvArgs<T> a;
  a.in = cur, a.out = out;
  a.wk = per_sample ? reinterpret_cast<const T*>(h->w_fold) : Weights16<T>::wk(L);
  a.wk_sm = Weights16<T>::sm(L), a.wk_up = Weights16<T>::up(L);
  a.bias = per_sample ? h->b_fold : L.bias;
  a.Cout = L.cout, a.CoutPad = L.pad_umma;
  a.relu = relu, a.epi = epi, a.halo_edge = halo_edge;
  a.out_nchw = out_nchw, a.out_u8 = epi == EPI_NCHW_F32 ? out_u8 : nullptr;
  a.tile_stats = tile_stats;
  a.sat_count = h->sat_count;
  a.per_sample = per_sample;
  if (split()) {
    CCST_CHECK_ARG(Weights16<T>::x3(L) != nullptr, "this layer has no split weights for the x3 engines");
    a.split = true, a.wk_x3 = Weights16<T>::x3(L), a.wk_x3_up = Weights16<T>::x3_up(L), a.out_scale = L.x3_scale;
  }
  return launch_conv_umma<T>(a, st);
}

int check_common(ccst_handle* h, int precision, bool allow_x3 = true) {
  CCST_CHECK_ARG(h != nullptr, "null handle");
  CCST_CHECK_ARG(precision == CCST_PREC_FP32 || precision == CCST_PREC_BF16 ||
                     precision == CCST_PREC_FP16 ||
                     ((precision == CCST_PREC_FP16X3 || precision == CCST_PREC_BF16X3) && allow_x3),
                 (precision == CCST_PREC_FP16X3 || precision == CCST_PREC_BF16X3)
                     ? "precision %d (an x3 engine) is not available for this entry point"
                     : "bad precision %d",
                 precision);
  CCST_CUDA(cudaSetDevice(h->device));
  if (int e = require_sm100()) return e;
  h->prof_n = 0;
  h->split = precision == CCST_PREC_FP16X3 || precision == CCST_PREC_BF16X3;
  return CCST_OK;
}

int ensure_io(ccst_handle* h, size_t elems) {
  if (h->io_elems >= elems) return CCST_OK;
  if (h->io_f32) CCST_CUDA(cudaFree(h->io_f32));
  h->io_f32 = nullptr, h->io_elems = 0;
  CCST_CUDA(cudaMalloc(&h->io_f32, elems * sizeof(float)));
  h->io_elems = elems;
  return CCST_OK;
}

template <typename T>
int run_style_transfer(ccst_handle* h, const float* d_img, int N, int H, int W, const float* mu,
                       const float* sg, int64_t stride, float alpha, float* d_out, cudaStream_t st) {
  int fh, fw;
  ccst_feature_hw(H, W, &fh, &fw);
  if (int e = ensure_arena(h, plan_bytes(N, H, W, fh, fw, sizeof(T) * (h->split ? 2 : 1), true, true))) return e;
  Pipe<T> p{h, st};
  if (int e = p.encoder(d_img, N, H, W)) return e;
  if (int e = p.adain(mu, sg, stride, alpha)) return e;
  return p.decoder(d_out);
}

// uint8 HWC in (the loader's PIL image before ToTensor) -> uint8 HWC out (what save_image encodes)
template <typename T>
int run_style_transfer_u8(ccst_handle* h, const uint8_t* d_img, int N, int H, int W, const float* mu,
                          const float* sg, int64_t stride, float alpha, uint8_t* d_out,
                          cudaStream_t st) {
  int fh, fw;
  ccst_feature_hw(H, W, &fh, &fw);
  const size_t in_elems = (size_t)N * 3 * H * W, out_elems = (size_t)N * 3 * (8 * fh) * (8 * fw);
  const bool fused_store = sizeof(T) == 2;  // tcgen05 path: quantisation fused into the last conv
  // ... and ToTensor into conv1_1's window loader when TMA can fetch the uint8 rows
  const bool fused_load = sizeof(T) == 2 && h->fuse_totensor && first_u8_ok(d_img, W);
  if (int e = ensure_io(h, (fused_load ? 0 : in_elems) + (fused_store ? 0 : out_elems))) return e;
  if (int e = ensure_arena(h, plan_bytes(N, H, W, fh, fw, sizeof(T) * (h->split ? 2 : 1), true, true))) return e;
  Pipe<T> p{h, st};
  if (fused_load) {
    p.img_u8 = d_img;
  } else {
    ProfScope ps(h, st, 5, 0, (double)in_elems * 5.0);
    if (int e = launch_u8_nhwc_to_f32_nchw(d_img, N, 3, H, W, h->io_f32, st)) return e;
  }
  if (int e = p.encoder(h->io_f32, N, H, W)) return e;
  if (int e = p.adain(mu, sg, stride, alpha)) return e;
  if (fused_store) {
    p.out_u8 = d_out;
    return p.decoder(nullptr);
  }
  float* tmp = h->io_f32 + in_elems;
  if (int e = p.decoder(tmp)) return e;
  ProfScope ps(h, st, 5, 0, (double)out_elems * 5.0);
  return launch_quantize_nchw_to_u8_nhwc(tmp, N, 3, 8 * fh, 8 * fw, d_out, st);
}

template <typename T>
int run_encoder(ccst_handle* h, const float* d_img, int N, int H, int W, float* d_feat,
                double* d_state, cudaStream_t st, float* const* lvl_mean = nullptr, float* const* lvl_std = nullptr,
                float lvl_eps = 1e-5f, const uint8_t* d_img_u8 = nullptr) {
  int fh, fw;
  ccst_feature_hw(H, W, &fh, &fw);
  if (int e = ensure_arena(h, plan_bytes(N, H, W, fh, fw, sizeof(T) * (h->split ? 2 : 1), true, false))) return e;
  Pipe<T> p{h, st};
  p.lvl_mean = lvl_mean, p.lvl_std = lvl_std, p.lvl_eps = lvl_eps;
  p.img_u8 = d_img_u8;
  if (int e = p.encoder(d_img, N, H, W)) return e;
  if (d_feat) {
    ProfScope ps(h, st, 5, 0, (double)N * fh * fw * 512 * (4.0 + sizeof(T)));
    if (int e = p.to_nchw(d_feat)) return e;
  }
  if (d_state) {
    ProfScope ps(h, st, 4, 0, (double)N * fh * fw * 512 * sizeof(T));
    if (p.stats_in_tiles) {
      if (int e = launch_stats_from_tiles(N, 512, fh, fw, h->raw, st)) return e;
    } else {
      if (int e = ensure_raw(h, nhwc_scratch_elems(N, 512, fh * fw))) return e;
      if (int e = p.stats_launch()) return e;
    }
    if (int e = merge_raw_into_state(h->raw, N, 512, (int64_t)fh * fw, d_state, st)) return e;
  }
  return CCST_OK;
}

template <typename T>
int run_decoder(ccst_handle* h, const float* d_feat, int N, int fh, int fw, float* d_img,
                cudaStream_t st) {
  if (int e = ensure_arena(h, plan_bytes(N, 0, 0, fh, fw, sizeof(T) * (h->split ? 2 : 1), false, true))) return e;
  Pipe<T> p{h, st};
  p.cur_slot = 0;
  p.cur = p.view(0, N, fh, fw, 512);
  {
    ProfScope ps(h, st, 5, 0, (double)N * fh * fw * 512 * (4.0 + sizeof(T)));
    if (int e = p.from_nchw(d_feat)) return e;
  }
  return p.decoder(d_img);
}

}  // namespace

// ------------------------------------------------------------------ exported functions
extern "C" int ccst_abi_version(void) { return CCST_ABI_VERSION; }
extern "C" const char* ccst_last_error(void) { return g_err; }
extern "C" int64_t ccst_launch_count(void) { return g_launches; }

extern "C" int ccst_check_device(int device) {
  if (int e = query_device(device)) return e;
  if (g_cc_major[device] != 10) {
    set_error("device %d has compute capability %d.x; libccst_b200 is built for sm_100a only",
              device, g_cc_major[device]);
    return CCST_EARCH;
  }
  return CCST_OK;
}

extern "C" int ccst_set_device(int device) {
  if (int e = ccst_check_device(device)) return e;
  CCST_CUDA(cudaSetDevice(device));
  return CCST_OK;
}

extern "C" void ccst_feature_hw(int H, int W, int* fh, int* fw) {
  int hh = H, ww = W;
  for (int i = 0; i < 3; ++i) hh = (hh + 1) / 2, ww = (ww + 1) / 2;
  if (fh) *fh = hh;
  if (fw) *fw = ww;
}

extern "C" ccst_handle* ccst_create(int device) {
  if (ccst_check_device(device) != CCST_OK) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) {
    set_error("cudaSetDevice(%d) failed", device);
    return nullptr;
  }
  ccst_handle* h = new ccst_handle();
  h->device = device;
  if (cudaMalloc(&h->sat_count, sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(h->sat_count, 0, sizeof(unsigned int)) != cudaSuccess) {
    set_error("ccst_create: cannot allocate the saturation counter: %s", cudaGetErrorString(cudaGetLastError()));
    delete h;
    return nullptr;
  }
  return h;
}

extern "C" void ccst_destroy(ccst_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->first_w27);
  cudaFree(h->first_b64);
  cudaFree(h->first_wk_b);
  cudaFree(h->first_wk_h);
  cudaFree(h->first_x3_b);
  cudaFree(h->first_x3_h);
  for (auto& L : h->enc) free_layer(L);
  for (auto& L : h->dec) free_layer(L);
  cudaFree(h->arena[0]);
  cudaFree(h->arena[1]);
  cudaFree(h->raw);
  cudaFree(h->io_f32);
  cudaFree(h->w_fold);
  cudaFree(h->b_fold);
  cudaFree(h->sat_count);
  for (auto& s : h->prof) {
    if (s.a) cudaEventDestroy(s.a);
    if (s.b) cudaEventDestroy(s.b);
  }
  delete h;
}

extern "C" int ccst_set_encoder_weights(ccst_handle* h, const float* const* w,
                                        const float* const* b) {
  CCST_CHECK_ARG(h && w && b, "ccst_set_encoder_weights: null argument");
  CCST_CUDA(cudaSetDevice(h->device));
  for (int i = 0; i < 10; ++i)
    CCST_CHECK_ARG(w[i] && b[i], "ccst_set_encoder_weights: null tensor %d", i);
  // fold the 1x1 colour conv (net.py:39) into conv1_1 (net.py:41): a pointwise op commutes with
  // reflection padding, so  conv3x3(pad(W1 x + b1)) = conv3x3'(pad(x)) with
  //   W'[o][i][t] = sum_c W2[o][c][t] W1[c][i],   b'[o] = b2[o] + sum_{c,t} W2[o][c][t] b1[c]
  const float *W1 = w[0], *b1 = b[0], *W2 = w[1], *b2 = b[1];
  std::vector<float> w27(27 * 64), b64(64);
  for (int o = 0; o < 64; ++o) {
    double bacc = b2[o];
    for (int t = 0; t < 9; ++t) {
      for (int i = 0; i < 3; ++i) {
        double acc = 0;
        for (int c = 0; c < 3; ++c) acc += (double)W2[((size_t)o * 3 + c) * 9 + t] * W1[c * 3 + i];
        w27[(size_t)(t * 3 + i) * 64 + o] = (float)acc;
      }
      for (int c = 0; c < 3; ++c) bacc += (double)W2[((size_t)o * 3 + c) * 9 + t] * b1[c];
    }
    b64[o] = (float)bacc;
  }
  if (!h->first_w27) CCST_CUDA(cudaMalloc(&h->first_w27, w27.size() * sizeof(float)));
  if (!h->first_b64) CCST_CUDA(cudaMalloc(&h->first_b64, b64.size() * sizeof(float)));
  CCST_CUDA(cudaMemcpy(h->first_w27, w27.data(), w27.size() * 4, cudaMemcpyHostToDevice));
  CCST_CUDA(cudaMemcpy(h->first_b64, b64.data(), b64.size() * 4, cudaMemcpyHostToDevice));
  std::vector<bf16> wkb(64 * 32);
  std::vector<__half> wkh(64 * 32);
  for (int o = 0; o < 64; ++o)
    for (int k = 0; k < 32; ++k) {
      const float v = k < 27 ? w27[(size_t)k * 64 + o] : 0.f;
      wkb[o * 32 + k] = __float2bfloat16(v);
      wkh[o * 32 + k] = __float2half(v);
    }
  if (!h->first_wk_b) CCST_CUDA(cudaMalloc(&h->first_wk_b, wkb.size() * sizeof(bf16)));
  if (!h->first_wk_h) CCST_CUDA(cudaMalloc(&h->first_wk_h, wkh.size() * sizeof(__half)));
  CCST_CUDA(cudaMemcpy(h->first_wk_b, wkb.data(), wkb.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  CCST_CUDA(cudaMemcpy(h->first_wk_h, wkh.data(), wkh.size() * sizeof(__half), cudaMemcpyHostToDevice));
  {
    // x3 engines: w * 2^e = hi + lo (e puts the largest |w| just below 2^10, see pack_layer); row o holds the 27
    // hi halves of channel o under k = 0..26 AND k = 32..58 (the [a_hi | a_lo] halves of the im2col row meet the
    // same weights), row 64 + o the lo halves
    float wmax = 0.f;
    for (float v : w27) wmax = fmaxf(wmax, fabsf(v));
    int e = 0;
    if (wmax > 0.f) {
      int ex;
      frexpf(wmax, &ex);
      e = 10 - ex;
    }
    std::vector<bf16> xb(128 * 64, __float2bfloat16(0.f));
    std::vector<__half> xh(128 * 64, __float2half(0.f));
    for (int o = 0; o < 64; ++o)
      for (int k = 0; k < 27; ++k) {
        const float v = ldexpf(w27[(size_t)k * 64 + o], e);
        const __half hh = __float2half(v), hl = __float2half(v - __half2float(hh));
        const bf16 bh = __float2bfloat16(v), bl = __float2bfloat16(v - __bfloat162float(bh));
        for (int half = 0; half < 2; ++half) {
          xh[(size_t)o * 64 + half * 32 + k] = hh, xh[(size_t)(64 + o) * 64 + half * 32 + k] = hl;
          xb[(size_t)o * 64 + half * 32 + k] = bh, xb[(size_t)(64 + o) * 64 + half * 32 + k] = bl;
        }
      }
    if (!h->first_x3_b) CCST_CUDA(cudaMalloc(&h->first_x3_b, xb.size() * sizeof(bf16)));
    if (!h->first_x3_h) CCST_CUDA(cudaMalloc(&h->first_x3_h, xh.size() * sizeof(__half)));
    CCST_CUDA(cudaMemcpy(h->first_x3_b, xb.data(), xb.size() * sizeof(bf16), cudaMemcpyHostToDevice));
    CCST_CUDA(cudaMemcpy(h->first_x3_h, xh.data(), xh.size() * sizeof(__half), cudaMemcpyHostToDevice));
    h->first_x3_scale = ldexpf(1.f, -e);
  }
  for (int i = 0; i < kEncLayers; ++i)
    if (int e = pack_layer(h->enc[i], kEncCh[i][0], kEncCh[i][1], w[2 + i], b[2 + i], false, false, /*split_pack=*/true))
      return e;
  h->enc_ready = true;
  return CCST_OK;
}

extern "C" int ccst_set_decoder_weights(ccst_handle* h, const float* const* w,
                                        const float* const* b) {
  CCST_CHECK_ARG(h && w && b, "ccst_set_decoder_weights: null argument");
  CCST_CUDA(cudaSetDevice(h->device));
  for (int i = 0; i < kDecLayers; ++i) {
    CCST_CHECK_ARG(w[i] && b[i], "ccst_set_decoder_weights: null tensor %d", i);
    if (int e = pack_layer(h->dec[i], kDecCh[i][0], kDecCh[i][1], w[i], b[i], i > 0 && kDecUpAfter[i - 1],
                           /*fold_src=*/i == 0, /*split_pack=*/true))
      return e;
  }
  h->dec_ready = true;
  return CCST_OK;
}

// expands `call` for the activation type selected by `precision`
#define CCST_DISPATCH(precision, call_T)                       \
  do {                                                         \
    if ((precision) == CCST_PREC_BF16 || (precision) == CCST_PREC_BF16X3) { \
      typedef bf16 T;                                          \
      return call_T;                                           \
    } else if ((precision) == CCST_PREC_FP16 || (precision) == CCST_PREC_FP16X3) { \
      typedef __half T;                                        \
      return call_T;                                           \
    } else {                                                   \
      typedef float T;                                         \
      return call_T;                                           \
    }                                                          \
  } while (0)

#define CCST_REQUIRE_STATE(cond, msg) \
  do {                                \
    if (!(cond)) {                    \
      set_error(msg);                 \
      return CCST_ESTATE;             \
    }                                 \
  } while (0)

extern "C" int ccst_encoder_fwd(ccst_handle* h, const float* d_img, int N, int H, int W,
                                float* d_feat, int precision, void* stream) {
  if (int e = check_common(h, precision)) return e;
  CCST_REQUIRE_STATE(h->enc_ready, "ccst_encoder_fwd: encoder weights not set");
  CCST_CHECK_ARG(d_img && d_feat && N >= 1 && H >= 8 && W >= 8, "ccst_encoder_fwd: bad argument");
  CCST_DISPATCH(precision, run_encoder<T>(h, d_img, N, H, W, d_feat, nullptr, (cudaStream_t)stream));
}

extern "C" int ccst_encoder_levels(ccst_handle* h, const float* d_img, int N, int H, int W, float* d_feat,
                                   float* const* d_mean, float* const* d_std, float eps, int precision,
                                   void* stream) {
  if (int e = check_common(h, precision)) return e;
  CCST_REQUIRE_STATE(h->enc_ready, "ccst_encoder_levels: encoder weights not set");
  CCST_CHECK_ARG(d_img && d_mean && d_std && N >= 1 && H >= 8 && W >= 8, "ccst_encoder_levels: bad argument");
  for (int l = 0; l < 4; ++l)
    CCST_CHECK_ARG(d_mean[l] && d_std[l], "ccst_encoder_levels: null output for level %d", l);
  CCST_DISPATCH(precision, run_encoder<T>(h, d_img, N, H, W, d_feat, nullptr, (cudaStream_t)stream, d_mean, d_std, eps));
}

extern "C" int ccst_mse_f32(const float* d_a, const float* d_b, int64_t n, double* d_scratch, float* d_out,
                            void* stream) {
  if (int e = require_sm100()) return e;
  CCST_CHECK_ARG(d_a && d_b && d_scratch && d_out && n >= 1, "ccst_mse_f32: bad argument");
  return launch_mse(d_a, d_b, n, d_scratch, d_out, (cudaStream_t)stream);
}

extern "C" int ccst_encoder_accumulate(ccst_handle* h, const float* d_img, int N, int H, int W,
                                       double* d_state, int precision, void* stream) {
  if (int e = check_common(h, precision)) return e;
  CCST_REQUIRE_STATE(h->enc_ready, "ccst_encoder_accumulate: encoder weights not set");
  CCST_CHECK_ARG(d_img && d_state && N >= 1 && H >= 8 && W >= 8,
                 "ccst_encoder_accumulate: bad argument");
  CCST_DISPATCH(precision, run_encoder<T>(h, d_img, N, H, W, nullptr, d_state, (cudaStream_t)stream));
}

namespace {
template <typename T>
int run_encoder_u8(ccst_handle* h, const uint8_t* d_img, int N, int H, int W, double* d_state, cudaStream_t st) {
  if (sizeof(T) == 2 && h->fuse_totensor && first_u8_ok(d_img, W))  // conv1_1 reads the uint8 batch itself
    return run_encoder<T>(h, nullptr, N, H, W, nullptr, d_state, st, nullptr, nullptr, 1e-5f, d_img);
  if (int e = ensure_io(h, (size_t)N * 3 * H * W)) return e;
  {
    ProfScope ps(h, st, 5, 0, (double)N * 3 * H * W * 5.0);
    if (int e = launch_u8_nhwc_to_f32_nchw(d_img, N, 3, H, W, h->io_f32, st)) return e;
  }
  return run_encoder<T>(h, h->io_f32, N, H, W, nullptr, d_state, st);
}
}  // namespace

extern "C" int ccst_encoder_accumulate_u8(ccst_handle* h, const uint8_t* d_img, int N, int H, int W,
                                          double* d_state, int precision, void* stream) {
  if (int e = check_common(h, precision)) return e;
  CCST_REQUIRE_STATE(h->enc_ready, "ccst_encoder_accumulate_u8: encoder weights not set");
  CCST_CHECK_ARG(d_img && d_state && N >= 1 && H >= 8 && W >= 8,
                 "ccst_encoder_accumulate_u8: bad argument");
  CCST_DISPATCH(precision, run_encoder_u8<T>(h, d_img, N, H, W, d_state, (cudaStream_t)stream));
}

extern "C" int ccst_decoder_fwd(ccst_handle* h, const float* d_feat, int N, int fh, int fw,
                                float* d_img, int precision, void* stream) {
  if (int e = check_common(h, precision)) return e;
  CCST_REQUIRE_STATE(h->dec_ready, "ccst_decoder_fwd: decoder weights not set");
  CCST_CHECK_ARG(d_feat && d_img && N >= 1 && fh >= 2 && fw >= 2, "ccst_decoder_fwd: bad argument");
  CCST_DISPATCH(precision, run_decoder<T>(h, d_feat, N, fh, fw, d_img, (cudaStream_t)stream));
}

extern "C" int ccst_style_transfer(ccst_handle* h, const float* d_img, int N, int H, int W,
                                   const float* d_mu_s, const float* d_sigma_s,
                                   int64_t stat_batch_stride, float alpha, float* d_out,
                                   int precision, void* stream) {
  if (int e = check_common(h, precision)) return e;
  CCST_REQUIRE_STATE(h->enc_ready && h->dec_ready, "ccst_style_transfer: weights not set");
  CCST_CHECK_ARG(d_img && d_out && d_mu_s && d_sigma_s && N >= 1 && H >= 16 && W >= 16,
                 "ccst_style_transfer: bad argument");
  CCST_CHECK_ARG(stat_batch_stride == 0 || stat_batch_stride == 512,
                 "ccst_style_transfer: stat_batch_stride must be 0 or 512");
  CCST_CHECK_ARG(alpha >= 0.f && alpha <= 1.f, "ccst_style_transfer: alpha outside [0,1]");
  CCST_DISPATCH(precision, run_style_transfer<T>(h, d_img, N, H, W, d_mu_s, d_sigma_s,
                                                 stat_batch_stride, alpha, d_out,
                                                 (cudaStream_t)stream));
}

extern "C" int ccst_style_transfer_u8(ccst_handle* h, const uint8_t* d_img, int N, int H, int W,
                                      const float* d_mu_s, const float* d_sigma_s,
                                      int64_t stat_batch_stride, float alpha, uint8_t* d_out,
                                      int precision, void* stream) {
  if (int e = check_common(h, precision)) return e;
  CCST_REQUIRE_STATE(h->enc_ready && h->dec_ready, "ccst_style_transfer_u8: weights not set");
  CCST_CHECK_ARG(d_img && d_out && d_mu_s && d_sigma_s && N >= 1 && H >= 16 && W >= 16,
                 "ccst_style_transfer_u8: bad argument");
  CCST_CHECK_ARG(stat_batch_stride == 0 || stat_batch_stride == 512,
                 "ccst_style_transfer_u8: stat_batch_stride must be 0 or 512");
  CCST_CHECK_ARG(alpha >= 0.f && alpha <= 1.f, "ccst_style_transfer_u8: alpha outside [0,1]");
  CCST_DISPATCH(precision, run_style_transfer_u8<T>(h, d_img, N, H, W, d_mu_s, d_sigma_s,
                                                    stat_batch_stride, alpha, d_out,
                                                    (cudaStream_t)stream));
}

extern "C" int ccst_u8_to_tensor(const uint8_t* d_img_nhwc, int N, int C, int H, int W,
                                 float* d_out_nchw, void* stream) {
  if (int e = require_sm100()) return e;
  CCST_CHECK_ARG(d_img_nhwc && d_out_nchw && N >= 1 && C >= 1 && H >= 1 && W >= 1,
                 "ccst_u8_to_tensor: bad argument");
  return launch_u8_nhwc_to_f32_nchw(d_img_nhwc, N, C, H, W, d_out_nchw, (cudaStream_t)stream);
}

extern "C" int ccst_quantize_u8(const float* d_img_nchw, int N, int C, int H, int W,
                                uint8_t* d_out_nhwc, void* stream) {
  if (int e = require_sm100()) return e;
  CCST_CHECK_ARG(d_img_nchw && d_out_nhwc && N >= 1 && C >= 1 && H >= 1 && W >= 1,
                 "ccst_quantize_u8: bad argument");
  return launch_quantize_nchw_to_u8_nhwc(d_img_nchw, N, C, H, W, d_out_nhwc, (cudaStream_t)stream);
}

extern "C" int ccst_resize_bilinear_aa_f32(const float* d_in, int64_t planes, int H, int W, int OH, int OW,
                                           float* d_out, void* stream) {
  if (int e = require_sm100()) return e;
  CCST_CHECK_ARG(d_in && d_out && planes >= 1 && H >= 1 && W >= 1 && OH >= 1 && OW >= 1,
                 "ccst_resize_bilinear_aa_f32: bad argument");
  return launch_resize_aa(d_in, planes, H, W, OH, OW, d_out, (cudaStream_t)stream);
}

extern "C" int ccst_set_fusion(ccst_handle* h, int mask) {
  CCST_CHECK_ARG(h != nullptr, "null handle");
  CCST_CHECK_ARG(mask >= 0 && mask <= CCST_FUSE_ALL, "ccst_set_fusion: bad mask %d", mask);
  h->fuse_pool = (mask & CCST_FUSE_POOL) != 0;
  h->fuse_up = (mask & CCST_FUSE_UPSAMPLE) != 0;
  h->fuse_stats = (mask & CCST_FUSE_STATS) != 0;
  h->fuse_adain = (mask & CCST_FUSE_ADAIN) != 0;
  h->fuse_totensor = (mask & CCST_FUSE_TOTENSOR) != 0;
  return CCST_OK;
}

extern "C" int ccst_saturation_snapshot(ccst_handle* h, uint32_t* h_count, void* stream) {
  CCST_CHECK_ARG(h != nullptr && h_count != nullptr, "ccst_saturation_snapshot: null argument");
  CCST_CUDA(cudaSetDevice(h->device));
  CCST_CUDA(cudaMemcpyAsync(h_count, h->sat_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return CCST_OK;
}

extern "C" int ccst_saturation_reset(ccst_handle* h, void* stream) {
  CCST_CHECK_ARG(h != nullptr, "null handle");
  CCST_CUDA(cudaSetDevice(h->device));
  CCST_CUDA(cudaMemsetAsync(h->sat_count, 0, sizeof(unsigned int), (cudaStream_t)stream));
  return CCST_OK;
}

namespace {
// Pillow's precompute_coeffs + normalize_coeffs_8bpc (Resample.c) for the bilinear filter (support 1),
// in the same double arithmetic: per output index the window {first, count} and its fixed-point weights.
int pil_bilinear_coeffs(int in_size, int out_size, std::vector<int>& kk, std::vector<int>& bounds) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  kk.assign((size_t)out_size * ksize, 0);
  bounds.assign((size_t)out_size * 2, 0);
  std::vector<double> w(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double v = (x + xmin - center + 0.5) * ss;
      if (v < 0.0) v = -v;
      w[x] = v < 1.0 ? 1.0 - v : 0.0;
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) w[x] /= ww;
      kk[(size_t)xx * ksize + x] = w[x] < 0 ? (int)(-0.5 + w[x] * (1 << 22)) : (int)(0.5 + w[x] * (1 << 22));
    }
    bounds[2 * xx] = xmin, bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}
size_t align256(size_t v) { return (v + 255) / 256 * 256; }
}  // namespace

extern "C" int64_t ccst_resize_pil_scratch_bytes(int N, int H, int W, int C, int OH, int OW) {
  if (N < 1 || H < 1 || W < 1 || C < 1 || OH < 1 || OW < 1) return 0;
  const int ksx = (int)ceil(W > OW ? (double)W / OW : 1.0) * 2 + 1, ksy = (int)ceil(H > OH ? (double)H / OH : 1.0) * 2 + 1;
  return (int64_t)(align256((size_t)OW * ksx * 4) + align256((size_t)OW * 8) + align256((size_t)OH * ksy * 4) +
                   align256((size_t)OH * 8) + align256((size_t)N * H * OW * C));
}

extern "C" int ccst_resize_pil_bilinear_u8(const uint8_t* d_in, int N, int H, int W, int C, int OH, int OW,
                                           uint8_t* d_out, void* d_scratch, void* stream) {
  if (int e = require_sm100()) return e;
  CCST_CHECK_ARG(d_in && d_out && d_scratch && N >= 1 && H >= 1 && W >= 1 && C >= 1 && OH >= 1 && OW >= 1,
                 "ccst_resize_pil_bilinear_u8: bad argument");
  std::vector<int> kx, bx, ky, by;
  const int ksx = pil_bilinear_coeffs(W, OW, kx, bx), ksy = pil_bilinear_coeffs(H, OH, ky, by);
  uint8_t* sp = static_cast<uint8_t*>(d_scratch);
  int* d_kx = reinterpret_cast<int*>(sp);
  sp += align256(kx.size() * 4);
  int* d_bx = reinterpret_cast<int*>(sp);
  sp += align256(bx.size() * 4);
  int* d_ky = reinterpret_cast<int*>(sp);
  sp += align256(ky.size() * 4);
  int* d_by = reinterpret_cast<int*>(sp);
  sp += align256(by.size() * 4);
  cudaStream_t st = (cudaStream_t)stream;
  // (pageable sources: the runtime stages them before returning, so the vectors may die at the end of the call)
  CCST_CUDA(cudaMemcpyAsync(d_kx, kx.data(), kx.size() * 4, cudaMemcpyHostToDevice, st));
  CCST_CUDA(cudaMemcpyAsync(d_bx, bx.data(), bx.size() * 4, cudaMemcpyHostToDevice, st));
  CCST_CUDA(cudaMemcpyAsync(d_ky, ky.data(), ky.size() * 4, cudaMemcpyHostToDevice, st));
  CCST_CUDA(cudaMemcpyAsync(d_by, by.data(), by.size() * 4, cudaMemcpyHostToDevice, st));
  return launch_resize_pil_u8(d_in, N, H, W, C, OH, OW, d_kx, d_bx, ksx, d_ky, d_by, ksy, sp, d_out, st);
}

extern "C" int ccst_profile_enable(ccst_handle* h, int on) {
  CCST_CHECK_ARG(h != nullptr, "null handle");
  h->profiling = on != 0;
  h->prof_n = 0;
  return CCST_OK;
}

extern "C" int ccst_profile_read(ccst_handle* h, int max, float* ms, double* flops, double* bytes,
                                 int* kind) {
  CCST_CHECK_ARG(h != nullptr, "null handle");
  int n = h->prof_n < max ? h->prof_n : max;
  for (int i = 0; i < n; ++i) {
    ProfSlot& s = h->prof[i];
    if (cudaEventSynchronize(s.b) != cudaSuccess) {
      set_error("ccst_profile_read: event sync failed: %s", cudaGetErrorString(cudaGetLastError()));
      return CCST_ECUDA;
    }
    float t = 0.f;
    cudaEventElapsedTime(&t, s.a, s.b);
    if (ms) ms[i] = t;
    if (flops) flops[i] = s.flops;
    if (bytes) bytes[i] = s.bytes;
    if (kind) kind[i] = s.kind;
  }
  return n;
}

namespace {
template <typename T>
int debug_conv(ccst_handle* h, const float* d_in, int N, int H, int W, int Cin, int Cout,
               const ConvLayer& L, int relu, int mode, float* d_out, cudaStream_t st) {
  int oh = H, ow = W, epi = EPI_ACT;
  if (mode == 1) oh = 2 * H, ow = 2 * W, epi = EPI_ACT_UP2;
  if (mode == 2) oh = (H + 1) / 2, ow = (W + 1) / 2, epi = EPI_ACT_POOL;
  if (mode == 3) epi = EPI_NCHW_F32;
  if (mode == 4) oh = 2 * H, ow = 2 * W, epi = EPI_UPS;  // upsample BEFORE the conv, fused
  const int in_edge = mode == 4 ? 0 : 1;
  T *bin = nullptr, *bout = nullptr, *btmp = nullptr;
  CCST_CUDA(cudaMalloc(&bin, act_bytes(N, H, W, Cin, sizeof(T))));
  ActView<T> vin{bin, N, H, W, Cin};
  int rc = launch_nhwc_to_act<T>(d_in, vin, st, in_edge);
  ActView<T> vout{nullptr, N, oh, ow, L.cout};
  if (rc == CCST_OK && mode != 3) {
    if (cudaMalloc(&bout, act_bytes(N, oh, ow, L.cout, sizeof(T))) != cudaSuccess) rc = CCST_ECUDA;
    vout.p = bout;
  }
  if (rc == CCST_OK) {
    Pipe<T> p{h, st};
    p.cur = vin;
    if (mode == 2 && sizeof(T) == 4) {
      // fp32 engine: conv then the stand-alone pool kernel
      ActView<T> full{nullptr, N, H, W, L.cout};
      if (cudaMalloc(&btmp, act_bytes(N, H, W, L.cout, sizeof(T))) != cudaSuccess) rc = CCST_ECUDA;
      full.p = btmp;
      if (rc == CCST_OK) rc = p.conv(L, relu, EPI_ACT, full, nullptr);
      if (rc == CCST_OK) rc = launch_pool<T>(full, vout, st);
    } else {
      rc = p.conv(L, relu, epi, vout, d_out);
    }
  }
  if (rc == CCST_OK && mode != 3) rc = launch_act_to_nhwc<T>(vout, d_out, st);
  cudaError_t se = cudaStreamSynchronize(st);
  if (rc == CCST_OK && se != cudaSuccess) {
    set_error("ccst_debug_conv3x3: %s", cudaGetErrorString(se));
    rc = CCST_ECUDA;
  }
  cudaFree(bin);
  cudaFree(bout);
  cudaFree(btmp);
  return rc;
}
}  // namespace

extern "C" int ccst_debug_conv3x3(ccst_handle* h, const float* d_in, int N, int H, int W, int Cin,
                                  int Cout, const float* h_weight, const float* h_bias, int relu,
                                  int mode, float* d_out, int precision, void* stream) {
  if (int e = check_common(h, precision, /*allow_x3=*/false)) return e;
  CCST_CHECK_ARG(d_in && d_out && h_weight && h_bias, "ccst_debug_conv3x3: null pointer");
  CCST_CHECK_ARG(N >= 1 && H >= 2 && W >= 2 && Cin % 64 == 0 && mode >= 0 && mode <= 4,
                 "ccst_debug_conv3x3: bad shape/mode");
  CCST_CHECK_ARG(mode != 4 || precision != CCST_PREC_FP32,
                 "ccst_debug_conv3x3: the upsample-fused conv exists on the tcgen05 path only");
  CCST_CHECK_ARG(mode == 3 ? Cout <= (precision == CCST_PREC_FP32 ? 16 : 3) : Cout % 64 == 0,
                 "ccst_debug_conv3x3: bad Cout");
  CCST_CHECK_ARG(mode != 2 || relu, "ccst_debug_conv3x3: pool mode requires relu");
  ConvLayer L;
  int rc = pack_layer(L, Cin, Cout, h_weight, h_bias, mode == 4);
  if (rc == CCST_OK) {
    if (precision == CCST_PREC_BF16)
      rc = debug_conv<bf16>(h, d_in, N, H, W, Cin, Cout, L, relu, mode, d_out, (cudaStream_t)stream);
    else if (precision == CCST_PREC_FP16)
      rc = debug_conv<__half>(h, d_in, N, H, W, Cin, Cout, L, relu, mode, d_out,
                              (cudaStream_t)stream);
    else
      rc = debug_conv<float>(h, d_in, N, H, W, Cin, Cout, L, relu, mode, d_out,
                             (cudaStream_t)stream);
  }
  free_layer(L);
  return rc;
}
