// Probe: tcgen05.ld throughput (TMEM -> registers).  W warps (4 or 8: 8 = two warps per lane quadrant)
// each read COLS columns of their 32-lane quadrant ITERS times.  Prints bytes per SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tmem_ld_bw tools/probes/tmem_ld_bw.cu && /tmp/tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

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

template <int COLS>
__global__ void probe(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < COLS; c += 32) {
      uint32_t r[32];
      tmem_ld32(base + ((c + (warp >> 2) * COLS) & 511), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

int main() {
  long long* d_cyc;
  uint32_t* d_sink;
  cudaMalloc(&d_cyc, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 148 * 256 * 4);
  const int iters = 2000;
  for (int warps : {4, 8}) {
    probe<256><<<148, warps * 32, 0>>>(iters, d_cyc, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double bytes = (double)warps * 32 * 256 * 4 * iters;  // lanes x cols x 4 B per iteration
    printf("warps=%d  %s  cycles=%lld  -> %.1f B/clk per SM (%.1f per warp)\n", warps, cudaGetErrorString(e), c,
           bytes / c, bytes / c / warps);
  }
  return 0;
}
