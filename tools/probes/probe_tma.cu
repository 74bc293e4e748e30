// Probe: which fp32 non-swizzled TMA boxes load correctly (debugging aid for conv1_1's window load).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o probe_tma probe_tma.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int box_elems, int cx, int cy, int cz, int cn) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  unsigned sb = (unsigned)__cvta_generic_to_shared(&bar);
  unsigned sd = (unsigned)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(box_elems * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(sd),
        "l"(&map), "r"(sb), "r"(cx), "r"(cy), "r"(cz), "r"(cn)
        : "memory");
  }
  unsigned ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(sb) : "memory");
    if (clock64() - t0 > 200000000LL) { if (threadIdx.x == 0) printf("timeout\n"); break; }
  }
  for (int i = threadIdx.x; i < box_elems; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1;
  int idx = -1;
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN enc = (PFN)fp;
  struct Case { int W, H, N, bx, by, bc, cx, cy; };
  Case cases[] = {{40, 40, 1, 132, 3, 3, -1, 5}, {512, 512, 2, 132, 3, 3, -1, 5}, {512, 512, 2, 132, 3, 3, 127, -1},
                  {512, 512, 2, 128, 3, 3, 0, 5},  {512, 512, 2, 64, 3, 3, -1, 5},  {512, 512, 2, 132, 1, 1, -1, 5},
                  {512, 512, 2, 132, 3, 1, -1, 5}, {512, 512, 2, 132, 1, 3, -1, 5}, {40, 40, 1, 44, 3, 3, -1, 5},
                  {512, 512, 2, 136, 3, 3, -4, 5}};
  for (auto c : cases) {
    if (++idx != only && only >= 0) continue;
    size_t n = (size_t)c.N * 3 * c.H * c.W;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 100003);
    float *d, *o;
    cudaMalloc(&d, n * 4);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    int be = c.bx * c.by * c.bc;
    cudaMalloc(&o, be * 4);
    cudaMemset(o, 0xff, be * 4);
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)c.W, (cuuint64_t)c.H, 3, (cuuint64_t)c.N};
    cuuint64_t st[3] = {(cuuint64_t)c.W * 4, (cuuint64_t)c.H * c.W * 4, (cuuint64_t)3 * c.H * c.W * 4};
    cuuint32_t box[4] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, (cuuint32_t)c.bc, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("case W=%d box={%d,%d,%d} at (%d,%d): encode=%d ", c.W, c.bx, c.by, c.bc, c.cx, c.cy, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); continue; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<<<1, 128, 65536>>>(m, o, be, c.cx, c.cy, 0, c.N - 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("launch error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> got(be);
    cudaMemcpy(got.data(), o, be * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int ch = 0; ch < c.bc; ++ch)
      for (int y = 0; y < c.by; ++y)
        for (int x = 0; x < c.bx; ++x) {
          int gx = c.cx + x, gy = c.cy + y;
          float want = 0.f;
          if (gx >= 0 && gx < c.W && gy >= 0 && gy < c.H)
            want = h[(((size_t)(c.N - 1) * 3 + ch) * c.H + gy) * c.W + gx];
          if (got[(ch * c.by + y) * c.bx + x] != want) ++bad;
        }
    printf("mismatches=%d\n", bad);
    cudaFree(d), cudaFree(o);
  }
  return 0;
}
