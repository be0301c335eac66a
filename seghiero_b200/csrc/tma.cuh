// TMA (cp.async.bulk.tensor) helpers shared by the kernels that move whole tiles with the copy engine (sm_100a):
// host side = tensor-map construction through the driver entry point (the library links cudart only), device side =
// box load + mbarrier transaction hand-over.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace sh {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;           // pure function pointer, resolved once
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// planes of [H][W] elements, box = boxw x boxh of one plane, out-of-bounds elements read as zero
inline bool make_plane_map(CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* base, int W, int H, long planes,
                           int boxw, int boxh) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr || ((uintptr_t)base & 15) || ((long)W * esize) % 16 || (boxw * esize) % 16) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)W * esize, (cuuint64_t)W * H * esize};
  const cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, 1u};
  const cuuint32_t es[3] = {1u, 1u, 1u};
  return enc(m, dt, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template <typename T> struct TmaType;
template <> struct TmaType<float> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };
template <> struct TmaType<__nv_bfloat16> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; };
template <> struct TmaType<__half> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; };

__device__ __forceinline__ void mbar_expect_tx(unsigned int addr, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned int dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned int mbar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(mbar) : "memory");
}
// acquire-wait on the box barrier; a transaction-count mistake must surface as a launch error, not as a hung GPU
__device__ __forceinline__ void mbar_wait_or_trap(unsigned int addr, unsigned int parity) {
  unsigned int ok, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();
  } while (!ok);
}

}  // namespace sh
