// conv1_1 (+ folded 1x1 colour conv) on tcgen05: thread-built im2col rows (K = 27 padded to 32).
#pragma once
#include "umma_common.cuh"

namespace ccst {
namespace {

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
  unsigned int* sat_count;  // see SatTracker
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
  SatTracker<T16> sat;

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
      sat.track_nonneg(pk[j]);  // (packed with ReLU)
      sat.track_nonneg(pk[16 + j]);
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
  sat.flush(p.sat_count);
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
constexpr int kF2OffLut = kF2OffBias + 256;  // 256 table words (U8 only)
constexpr int kF2OffBar = kF2OffLut + 1024;
constexpr int kF2NumBars = 2 * kF2WinStages + 8;
constexpr int kF2Smem = 1024 + kF2OffBar + 8 * kF2NumBars + 16;

// (n, y, x tile) of the tiles blockIdx.x, + stride, + 2 stride ... kept by additions: the plain decode costs two
// integer divisions and two remainders (~80 instructions) per tile in EVERY warp, a third of the instructions the
// warp-specialised first-layer kernels issued per tile (ncu: issue slots 59 % busy at 3.8 TB/s).
struct FirstTileIter {
  int tx, y, n;
  int step_tx, step_y, step_n, tiles_x, H;
  __device__ __forceinline__ void init(int tiles_x_, int H_, int tile0, int stride) {
    tiles_x = tiles_x_, H = H_;
    tx = tile0 % tiles_x;
    const int row0 = tile0 / tiles_x;
    y = row0 % H, n = row0 / H;
    step_tx = stride % tiles_x;
    const int rows = stride / tiles_x;
    step_y = rows % H, step_n = rows / H;
  }
  __device__ __forceinline__ void next() {
    tx += step_tx;
    int carry = 0;
    if (tx >= tiles_x) tx -= tiles_x, carry = 1;
    y += step_y + carry;  // < 2 H
    if (y >= H) y -= H, ++n;
    n += step_n;
  }
};

// U8 = true: the window is fetched from the loader's uint8 HWC batch [N,H,W,3] itself (ToTensor fused into the
// builders, cjm_util/data_helper.py:45): TMA box {104 x 4 bytes, 3 rows} = bytes x0*3 - 16 .. x0*3 + 399 of rows
// y-1..y+1 (needs 3 W % 16 == 0), and float(b) / 255 rounded to T16 comes from a 256-entry table (low half: the
// T16 value, high half: the T16 remainder for the x3 engines) -- the same values as ccst_u8_to_tensor + pack.
constexpr int kU8RowBytes = 416;  // 16 + 130 pixels x 3, rounded up to 16
constexpr int kU8WinX0 = 16;      // window byte of pixel x0, channel 0
constexpr int kU8WinTx = 3 * kU8RowBytes;
static_assert(kU8WinTx <= kF2WinBytes, "the uint8 window fits the fp32 window's slot");

template <typename T16>
__device__ __forceinline__ uint32_t u8_lut_word(int b) {
  const float v = __fdiv_rn((float)b, 255.f);  // ToTensor: IEEE division, as torch's div
  const uint32_t hi = pack16x2<T16>(v, 0.f) & 0xffffu;
  const uint32_t lo = pack16x2<T16>(v - unpack16x2<T16>(hi).x, 0.f) & 0xffffu;
  return hi | (lo << 16);
}
// the 27 taps of pixel px as table words (k = (r*3+s)*3 + ci)
__device__ __forceinline__ void u8_gather27(const uint8_t* win, const uint32_t* lut, const int (&ridx)[3], int x0, int px,
                                            int W, uint32_t (&wv)[28]) {
  int cidx[3];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    int xx = x0 + px + s - 1;
    xx = xx < 0 ? -xx : (xx >= W ? 2 * W - 2 - xx : xx);
    const int c = kU8WinX0 + (xx - x0) * 3;
    cidx[s] = c < 0 ? 0 : (c > kU8RowBytes - 3 ? kU8RowBytes - 3 : c);  // only for pixels past W
  }
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const int tap = k / 3, ci = k - 3 * tap;
    const int r = tap / 3, s = tap - 3 * r;
    wv[k] = lut[win[ridx[r] * kU8RowBytes + cidx[s] + ci]];
  }
  wv[27] = 0u;
}

template <typename T16, bool U8 = false>
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
  uint32_t* lut = reinterpret_cast<uint32_t*>(gen + kF2OffLut);
  if (U8 && tid < 256) lut[tid] = u8_lut_word<T16>(tid);
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

  FirstTileIter ti;  // this thread's walk over the CTA's tiles
  ti.init(p.tiles_x, p.H, blockIdx.x, gridDim.x);

  if (warp == 0) {
    // ===================== TMA producer
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it, ti.next()) {
      const int ws = it % kF2WinStages;
      const uint32_t ph = (it / kF2WinStages) & 1;
      const int n = ti.n, y = ti.y, x0 = ti.tx * kFirstPx;
      mbar_wait(win_empty(ws), ph ^ 1, 910);
      if (elect_one()) {
        mbar_expect_tx(win_full(ws), U8 ? kU8WinTx : kF2WinTx);
        tma_load_4d(base + kF2OffWin + ws * kF2WinBytes, &tmap_img, win_full(ws),
                    U8 ? (x0 * 3 - kU8WinX0) >> 2 : x0 - kF2WinX0, y - 1, 0, n);
      }
      __syncwarp();
    }
  } else if (warp <= 4) {
    // ===================== builders: im2col row of pixel px of the tile
    const int px = tid - 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it, ti.next()) {
      const int ws = it % kF2WinStages, as = it & 1;
      const int y = ti.y, x0 = ti.tx * kFirstPx;
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
      uint32_t pk[16];
      if (U8) {
        uint32_t wv[28];
        u8_gather27(gen + kF2OffWin + ws * kF2WinBytes, lut, ridx, x0, px, p.W, wv);
        fence_async_smem();  // (as below)
        mbar_arrive(win_empty(ws));
#pragma unroll
        for (int k2 = 0; k2 < 14; ++k2) pk[k2] = __byte_perm(wv[2 * k2], wv[2 * k2 + 1], 0x5410);
      } else {
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
#pragma unroll
        for (int k2 = 0; k2 < 14; ++k2) pk[k2] = pack16x2<T16>(v[2 * k2], v[2 * k2 + 1]);
      }
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
    SatTracker<T16> sat;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it, ti.next()) {
      const int as = it & 1;
      const int n = ti.n, y = ti.y, x0 = ti.tx * kFirstPx;
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
        // packed fp32 adds (each half rounds like a scalar add)
        const float2 a0 = add2_f32(make_float2(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1])),
                                   *reinterpret_cast<const float2*>(&sbias[2 * j]));
        const float2 a1 = add2_f32(make_float2(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1])),
                                   *reinterpret_cast<const float2*>(&sbias[32 + 2 * j]));
        pk[j] = pack16x2_relu<T16>(a0.x, a0.y);
        pk[16 + j] = pack16x2_relu<T16>(a1.x, a1.y);
        sat.track_nonneg(pk[j]);
        sat.track_nonneg(pk[16 + j]);
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
    sat.flush(p.sat_count);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<128>(tmem_base);
}

// =====================================================================================
// conv1_1 of the x3 engines (split operands, conv_x3.cuh) on the tensor pipe: the warp-specialised kernel
// above with 64-wide im2col rows [a_hi (27, padded to 32) | a_lo (27, padded to 32)] and a 128-row weight
// tile [w_hi | w_lo] whose rows repeat the same 27 weights under both halves of K, so that
//   D[:, 0..63] = (a_hi + a_lo) * w_hi,   D[:, 64..127] = (a_hi + a_lo) * w_lo      (4 MMAs of K = 16, N = 128).
// Four MMAs per accumulator: the truncating TMEM accumulation that forces the promoted partial sums of the
// deeper layers is negligible here.  Epilogue: v = (D_hi + D_lo) * 2^-e + bias, ReLU, split into hi / lo, two
// TMA stores per warp (channels 0..63 and 64..127 of the [hi | lo] map).  HBM-bound on its 256 B per pixel.
// One CTA per SM (two staging tiles per accumulator stage).
// =====================================================================================
constexpr int kF3OffA = 0;                                     // 2 A tiles (128 rows x 128 B, all 64 k used)
constexpr int kF3OffOut = 2 * kF2ABytes;                       // 2 stages x {hi, lo} staging tiles
constexpr int kF3OffB = 6 * kF2ABytes;                         // weights 128 x 128 B
constexpr int kF3OffWin = kF3OffB + 128 * 128;
constexpr int kF3OffBias = kF3OffWin + kF2WinStages * kF2WinBytes;
constexpr int kF3OffLut = kF3OffBias + 256;
constexpr int kF3OffBar = kF3OffLut + 1024;
constexpr int kF3Smem = 1024 + kF3OffBar + 8 * kF2NumBars + 16;

constexpr int kF3Threads = 448;  // producer, 4 builder warps, MMA issuer, 2 epilogue groups of 4 warps

template <typename T16, bool U8 = false>
__global__ void __launch_bounds__(kF3Threads, 1)
    conv_first_x3_ws_kernel(const __grid_constant__ CUtensorMap tmap_img,
                            const __grid_constant__ CUtensorMap tmap_out, FirstParams<T16> p, float out_scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kF3OffBar;
  auto win_full = [&](int s) { return bar0 + 8u * s; };
  auto win_empty = [&](int s) { return bar0 + 8u * (kF2WinStages + s); };
  auto a_full = [&](int s) { return bar0 + 8u * (2 * kF2WinStages + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 * kF2WinStages + 2 + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * kF2WinStages + 4 + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * kF2WinStages + 6 + s); };
  const uint32_t tmem_slot = bar0 + 8u * kF2NumBars;
  float* sbias = reinterpret_cast<float*>(gen + kF3OffBias);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // one-time: weights [128 rows = hi | lo][64 k] -> swizzled K-major B tile, bias, barriers, TMEM (2 x 128 columns)
  for (int i = tid; i < 128 * 8; i += kF3Threads) {
    const int o = i >> 3, j = i & 7;
    const uint4 v = reinterpret_cast<const uint4*>(p.wk)[o * 8 + j];
    *reinterpret_cast<uint4*>(gen + kF3OffB + o * 128 + ((j ^ (o & 7)) << 4)) = v;
  }
  if (tid < 64) sbias[tid] = p.bias[tid];
  uint32_t* lut = reinterpret_cast<uint32_t*>(gen + kF3OffLut);
  if (U8 && tid < 256) lut[tid] = u8_lut_word<T16>(tid);
  if (tid == 0) {
    for (int s = 0; s < kF2WinStages; ++s) {
      mbar_init(win_full(s), 1);
      mbar_init(win_empty(s), 128);
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
  if (warp == 5) tmem_alloc<256>(tmem_slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + kF3OffBar + 8 * kF2NumBars);

  FirstTileIter ti;  // this thread's walk over the CTA's tiles
  ti.init(p.tiles_x, p.H, blockIdx.x, gridDim.x);

  if (warp == 0) {
    // ===================== TMA producer
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it, ti.next()) {
      const int ws = it % kF2WinStages;
      const uint32_t ph = (it / kF2WinStages) & 1;
      const int n = ti.n, y = ti.y, x0 = ti.tx * kFirstPx;
      mbar_wait(win_empty(ws), ph ^ 1, 960);
      if (elect_one()) {
        mbar_expect_tx(win_full(ws), U8 ? kU8WinTx : kF2WinTx);
        tma_load_4d(base + kF3OffWin + ws * kF2WinBytes, &tmap_img, win_full(ws),
                    U8 ? (x0 * 3 - kU8WinX0) >> 2 : x0 - kF2WinX0, y - 1, 0, n);
      }
      __syncwarp();
    }
  } else if (warp <= 4) {
    // ===================== builders: split im2col row of pixel px
    const int px = tid - 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it, ti.next()) {
      const int ws = it % kF2WinStages, as = it & 1;
      const int y = ti.y, x0 = ti.tx * kFirstPx;
      int ridx[3] = {0, 1, 2};
      if (y == 0) ridx[0] = 2;
      if (y == p.H - 1) ridx[2] = 0;
      int cidx[3];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int xx = x0 + px + s - 1;
        xx = xx < 0 ? -xx : (xx >= p.W ? 2 * p.W - 2 - xx : xx);
        int c = xx - (x0 - kF2WinX0);
        cidx[s] = c < 0 ? 0 : (c > kF2WinCols - 1 ? kF2WinCols - 1 : c);
      }
      mbar_wait(win_full(ws), (it / kF2WinStages) & 1, 961);
      uint32_t pk[32];
      if (U8) {
        uint32_t wv[28];
        u8_gather27(gen + kF3OffWin + ws * kF2WinBytes, lut, ridx, x0, px, p.W, wv);
        fence_async_smem();
        mbar_arrive(win_empty(ws));
#pragma unroll
        for (int k2 = 0; k2 < 14; ++k2) {
          pk[k2] = __byte_perm(wv[2 * k2], wv[2 * k2 + 1], 0x5410);       // hi halves
          pk[16 + k2] = __byte_perm(wv[2 * k2], wv[2 * k2 + 1], 0x7632);  // lo halves
        }
      } else {
        const float* win = reinterpret_cast<const float*>(gen + kF3OffWin + ws * kF2WinBytes);
        float v[28];
#pragma unroll
        for (int k = 0; k < 27; ++k) {
          const int tap = k / 3, ci = k - 3 * tap;
          const int r = tap / 3, s = tap - 3 * r;
          v[k] = win[(ci * 3 + ridx[r]) * kF2WinCols + cidx[s]];
        }
        v[27] = 0.f;
        fence_async_smem();  // (see conv_first_umma_ws_kernel: generic reads before the async-proxy rewrite)
        mbar_arrive(win_empty(ws));
#pragma unroll
        for (int k2 = 0; k2 < 14; ++k2) {
          pk[k2] = pack16x2<T16>(v[2 * k2], v[2 * k2 + 1]);
          const float2 hf = unpack16x2<T16>(pk[k2]);
          pk[16 + k2] = pack16x2<T16>(v[2 * k2] - hf.x, v[2 * k2 + 1] - hf.y);
        }
      }
      pk[14] = 0u, pk[15] = 0u, pk[30] = 0u, pk[31] = 0u;
      MBAR_WAIT_RELAXED(a_empty(as), ((it >> 1) & 1) ^ 1, 962);
      const uint32_t sA = base + kF3OffA + as * kF2ABytes;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t dst = sA + px * 128 + ((j ^ (px & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]),
                     "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                     : "memory");
      }
      fence_async_smem();
      mbar_arrive(a_full(as));
    }
  } else if (warp == 5) {
    // ===================== MMA issuer: K = 64 = [hi | lo] in four steps, N = 128 = [w_hi | w_lo]
    constexpr uint32_t idesc = make_idesc<T16, 128>();
    const uint64_t bdesc = make_kmajor_sw128_desc(base + kF3OffB);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(t_empty(as), ph ^ 1, 963);
      mbar_wait(a_full(as), ph, 964);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = make_kmajor_sw128_desc(base + kF3OffA + as * kF2ABytes);
        const uint32_t d = tmem_base + (uint32_t)(as * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
        umma_commit(a_empty(as));
        umma_commit(t_full(as));
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue: group g = tiles of parity g, accumulator stage g, staging pair g; warp q of a
    // group owns TMEM lanes 32q..32q+31 = pixels 32q.. of the tile
    const int grp = (warp - 6) >> 2;
    const int q = warp & 3;
    const int px = q * 32 + lane;
    SatTracker<T16> sat;
    const uint32_t sHi = base + kF3OffOut + (2 * grp) * kF2ABytes, sLo = sHi + kF2ABytes;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * 128);
    int it = grp;
    if ((long long)blockIdx.x + (long long)grp * gridDim.x < p.total_tiles)
      ti.init(p.tiles_x, p.H, blockIdx.x + grp * gridDim.x, 2 * gridDim.x);
    for (long long tile = (long long)blockIdx.x + (long long)grp * gridDim.x; tile < p.total_tiles;
         tile += 2ll * gridDim.x, it += 2, ti.next()) {
      const int n = ti.n, y = ti.y, x0 = ti.tx * kFirstPx;
      const int x = x0 + px;
      MBAR_WAIT_RELAXED(t_full(grp), (it >> 1) & 1, 965);
      tc_fence_after();
      // the group's staging pair: the stores of its previous tile must have read it
      bulk_wait_read<0>();
      __syncwarp();
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t u[32], w[32];
        tmem_ld32(taddr + h2 * 32, u);       // a * w_hi, channels 32 h2 ..
        tmem_ld32(taddr + 64 + h2 * 32, w);  // a * w_lo
        tmem_ld_wait();
        if (h2 == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty(grp));
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v0 = fmaxf(fmaf(__uint_as_float(u[2 * j]) + __uint_as_float(w[2 * j]), out_scale,
                                      sbias[h2 * 32 + 2 * j]), 0.f);
          const float v1 = fmaxf(fmaf(__uint_as_float(u[2 * j + 1]) + __uint_as_float(w[2 * j + 1]), out_scale,
                                      sbias[h2 * 32 + 2 * j + 1]), 0.f);
          hi[j] = pack16x2<T16>(v0, v1);
          const float2 hf = unpack16x2<T16>(hi[j]);
          lo[j] = pack16x2<T16>(v0 - hf.x, v1 - hf.y);
          sat.track_nonneg(hi[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = px * 128 + (((h2 * 4 + j) ^ (px & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sHi + off), "r"(hi[4 * j]),
                       "r"(hi[4 * j + 1]), "r"(hi[4 * j + 2]), "r"(hi[4 * j + 3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sLo + off), "r"(lo[4 * j]),
                       "r"(lo[4 * j + 1]), "r"(lo[4 * j + 2]), "r"(lo[4 * j + 3])
                       : "memory");
        }
        // reflection-halo aliases of a border pixel (the pixel itself goes out through the TMA store)
        if (x < p.W && (y == 1 || y == p.H - 2 || x == 1 || x == p.W - 2)) {
          for_each_halo_alias(y, x, p.H, p.W, [&](int yy, int xx) {
            if (yy == y && xx == x) return;
            uint4* dh = reinterpret_cast<uint4*>(p.out.px(n, yy, xx) + h2 * 32);
            uint4* dl = reinterpret_cast<uint4*>(p.out.px(n, yy, xx) + 64 + h2 * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
              dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
          });
        }
      }
      fence_async_smem();
      __syncwarp();
      if (elect_one()) {
        tma_store_4d(&tmap_out, sHi + q * (32 * 128), 0, x0 + q * 32, y, n);  // clipped at W
        tma_store_4d(&tmap_out, sLo + q * (32 * 128), 64, x0 + q * 32, y, n);
        bulk_commit();
      }
      __syncwarp();
    }
    bulk_wait_all();
    sat.flush(p.sat_count);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<256>(tmem_base);
}

}  // namespace
}  // namespace ccst
