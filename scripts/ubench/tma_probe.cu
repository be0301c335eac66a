// Probe: one cp.async.bulk.tensor.3d box load (with out-of-bounds start) exactly as k3t_pass2 issues it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, const CUtensorMap* gmap, int mode, float* out, int cols, int rows, int c0, int c1, int c2) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned int mb = (unsigned int)__cvta_generic_to_shared(&bar);
  const unsigned int dst = (unsigned int)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(mb), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (mode == 0) {
      asm volatile("mbarrier.arrive.shared.b64 _, [%0];" ::"r"(mb) : "memory");
    } else {
      const CUtensorMap* mp = mode == 1 ? &map : gmap;
      asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mb), "r"(cols * rows * 4) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(dst), "l"(mp), "r"(c0), "r"(c1), "r"(c2), "r"(mb) : "memory");
    }
  }
  unsigned int ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(mb), "r"(0) : "memory");
  } while (!ok);
  const float* s = reinterpret_cast<const float*>(smem);
  for (int i = threadIdx.x; i < cols * rows; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv) {
  const int W = 128, H = 64, P = 3;
  const int cols = argc > 1 ? atoi(argv[1]) : 68, rows = argc > 2 ? atoi(argv[2]) : 36;
  std::vector<float> h((size_t)W * H * P);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, cols * rows * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap m;
  const cuuint64_t dims[3] = {W, H, P};
  const cuuint64_t strides[2] = {W * 4, (cuuint64_t)W * H * 4};
  const cuuint32_t box[3] = {(cuuint32_t)cols, (cuuint32_t)rows, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d (cols %d rows %d)\n", (int)r, cols, rows);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int mode = argc > 3 ? atoi(argv[3]) : 1;
  CUtensorMap* gm;
  cudaMalloc(&gm, sizeof(CUtensorMap));
  cudaMemcpy(gm, &m, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  const int c0 = argc > 4 ? atoi(argv[4]) : 62, c1 = argc > 5 ? atoi(argv[5]) : -2;
  printf("mode %d coords %d %d\n", mode, c0, c1);
  probe<<<1, 128, cols * rows * 4 + 128>>>(m, gm, mode, o, cols, rows, c0, c1, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s\n", cudaGetErrorString(e));
  std::vector<float> res(cols * rows);
  cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int rr = 0; rr < rows; ++rr)
    for (int cc = 0; cc < cols; ++cc) {
      const int x = c0 + cc, y = c1 + rr;
      const float want = (x >= 0 && x < W && y >= 0 && y < H) ? (float)((size_t)1 * W * H + (size_t)y * W + x) : 0.f;
      if (res[rr * cols + cc] != want) ++bad;
    }
  printf("mismatches %d of %d\n", bad, cols * rows);
  return 0;
}
