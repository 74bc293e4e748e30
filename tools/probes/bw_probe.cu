// Pure-read / pure-write / copy HBM bandwidth of this GPU with plain vectorised kernels and cudaMemset:
// the ceilings a write-dominated (conv1_1, dec8) or read-dominated (dec9, conv1_2) kernel is to be read against.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probes/bw_probe tools/probes/bw_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_write(uint4* __restrict__ p, size_t n, uint4 v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_read(const uint4* __restrict__ p, size_t n, unsigned* out) {
  unsigned acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(p + i);
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *out = acc;
}
__global__ void k_copy(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = __ldg(a + i);
}
// 1 write : 8 read and 8 write : 1 read mixes (dec9-like, conv1_1-like)
__global__ void k_mix(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n, int rd, int wr, unsigned* out) {
  unsigned acc = 0;
  const int per = rd + wr;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int slot = (int)((i / 32) % per);  // warp-granular choice keeps accesses coalesced
    if (slot < rd) {
      const uint4 v = __ldg(a + i);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    } else {
      b[i] = make_uint4(1, 2, 3, 4);
    }
  }
  if (acc == 0x12345678u) *out = acc;
}

template <typename F>
float time_ms(F f, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

int main() {
  const size_t bytes = (size_t)1 << 30;  // 1 GiB per buffer, far above the 126 MB L2
  uint4 *a, *b;
  unsigned* out;
  cudaMalloc(&a, bytes), cudaMalloc(&b, bytes), cudaMalloc(&out, 4);
  cudaMemset(a, 1, bytes), cudaMemset(b, 2, bytes);
  const size_t n = bytes / 16;
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int cta_per_sm : {4, 8, 16}) {
    const int grid = sms * cta_per_sm, blk = 256;
    const float w = time_ms([&] { k_write<<<grid, blk>>>(a, n, make_uint4(1, 2, 3, 4)); }, 10);
    const float r = time_ms([&] { k_read<<<grid, blk>>>(a, n, out); }, 10);
    const float c = time_ms([&] { k_copy<<<grid, blk>>>(a, b, n); }, 10);
    const float m81 = time_ms([&] { k_mix<<<grid, blk>>>(a, b, n, 8, 1, out); }, 10);
    const float m18 = time_ms([&] { k_mix<<<grid, blk>>>(a, b, n, 1, 8, out); }, 10);
    const float m14 = time_ms([&] { k_mix<<<grid, blk>>>(a, b, n, 1, 4, out); }, 10);
    printf("grid %4d x %d: write %.0f GB/s, read %.0f GB/s, copy %.0f GB/s (r+w), mix 8r:1w %.0f GB/s, 1r:8w %.0f GB/s, 1r:4w %.0f GB/s\n",
           grid, blk, bytes / w / 1e6, bytes / r / 1e6, 2.0 * bytes / c / 1e6, bytes / m81 / 1e6, bytes / m18 / 1e6, bytes / m14 / 1e6);
  }
  const float ms = time_ms([&] { cudaMemsetAsync(a, 0, bytes); }, 10);
  const float mc = time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 10);
  printf("cudaMemset %.0f GB/s, cudaMemcpy D2D %.0f GB/s (r+w)\n", bytes / ms / 1e6, 2.0 * bytes / mc / 1e6);
  return 0;
}
