// Probe 2: do back-to-back launches with DIFFERENT tensor maps passed as __grid_constant__ parameters see their own map?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, float expect_base, int* bad) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned int mb = (unsigned int)__cvta_generic_to_shared(&bar);
  const unsigned int dst = (unsigned int)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(mb), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  for (int it = 0; it < 8; ++it) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mb), "r"(64 * 32 * 4) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(dst), "l"(&map), "r"(64 * (int)(blockIdx.x & 1)), "r"(32 * it), "r"(0), "r"(mb) : "memory");
    }
    unsigned int ok, spins = 0;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(mb), "r"(it & 1) : "memory");
      if (!ok && ++spins > (1u << 20)) { atomicAdd(bad + 1, 1); break; }
    } while (!ok);
    const float* s = reinterpret_cast<const float*>(smem);
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
      const int x = 64 * (blockIdx.x & 1) + (i & 63), y = 32 * it + (i >> 6);
      if (s[i] != expect_base + (float)(y * 128 + x)) atomicAdd(bad, 1);
    }
    __syncthreads();
  }
}
int main() {
  const int W = 128, H = 256;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  float* buf[4];
  std::vector<float> h((size_t)W * H);
  for (int k = 0; k < 4; ++k) {
    cudaMalloc(&buf[k], h.size() * 4);
    for (size_t i = 0; i < h.size(); ++i) h[i] = 100000.f * k + (float)i;
    cudaMemcpy(buf[k], h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  }
  int* bad;
  cudaMalloc(&bad, 8);
  cudaMemset(bad, 0, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int it = 0; it < 2000; ++it) {
    const int k = it & 3;
    CUtensorMap m;
    const cuuint64_t dims[3] = {W, H, 1};
    const cuuint64_t strides[2] = {W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {64, 32, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf[k], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    probe<<<592, 128, 64 * 32 * 4>>>(m, 100000.f * k, bad);
  }
  cudaError_t e = cudaDeviceSynchronize();
  int hb[2];
  cudaMemcpy(hb, bad, 8, cudaMemcpyDeviceToHost);
  printf("run: %s, mismatching elements %d, stalled waits %d\n", cudaGetErrorString(e), hb[0], hb[1]);
  return 0;
}
