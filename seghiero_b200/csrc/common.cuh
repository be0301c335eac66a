// Shared device helpers for the seghiero_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define SH_OK 0
#define SH_ERR_BAD_ARG (-1)
#define SH_ERR_UNSUPPORTED (-2)

#define SH_DT_F32 0
#define SH_DT_BF16 1
#define SH_DT_F16 2

// label element types accepted at the C ABI (the reference uses int64; uint8 / int32 cut the label traffic 8x / 2x)
#define SH_LAB_I64 0
#define SH_LAB_I32 1
#define SH_LAB_U8 2

#define SH_IGNORE 255
// SM count of the current device (148 on a B200), queried once: grids of the persistent kernels are sized from it.
// All devices of a box are the same part; the cached integer is the library's only other process-wide state besides
// the cuTensorMapEncodeTiled entry point.
inline int sh_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        v > 0)
      n = v;
    else
      n = 148;
  }
  return n;
}
#define SH_NUM_SMS sh_num_sms()

#define SH_CHECK_LAUNCH()                          \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

// run `...` with the label element type bound to the name L; returns SH_ERR_BAD_ARG for an unknown code
#define SH_LABEL_SWITCH(code, L, ...)                                   \
  switch (code) {                                                       \
    case SH_LAB_I64: { using L = long long; __VA_ARGS__; } break;       \
    case SH_LAB_I32: { using L = int; __VA_ARGS__; } break;             \
    case SH_LAB_U8: { using L = unsigned char; __VA_ARGS__; } break;    \
    default: return SH_ERR_BAD_ARG;                                     \
  }

namespace sh {

// labels of any accepted element type as int64 values (uint8 255 stays 255 = ignore; int32 sign-extends)
template <typename L>
__device__ __forceinline__ long long lab_ld(const L* p, long i) { return (long long)p[i]; }
// two consecutive labels; `vec` = the pair is naturally aligned (one load)
template <typename L>
__device__ __forceinline__ void lab_ld2(const L* p, bool vec, long long& a, long long& b) {
  if (sizeof(L) == 8) {
    if (vec) { const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(p)); a = v.x; b = v.y; return; }
  } else if (sizeof(L) == 4) {
    if (vec) { const int2 v = __ldg(reinterpret_cast<const int2*>(p)); a = v.x; b = v.y; return; }
  } else {
    if (vec) { const unsigned short v = __ldg(reinterpret_cast<const unsigned short*>(p)); a = v & 0xffu; b = v >> 8; return; }
  }
  a = (long long)p[0]; b = (long long)p[1];
}
// 16 bytes of labels = 16 / sizeof(L) values
template <typename L>
__device__ __forceinline__ void lab_ld16(const L* p, long long (&o)[16 / sizeof(L)]) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const unsigned int w[4] = {r.x, r.y, r.z, r.w};
  if (sizeof(L) == 8) {
    o[0] = (long long)((unsigned long long)w[0] | ((unsigned long long)w[1] << 32));
    o[1 % (16 / sizeof(L))] = (long long)((unsigned long long)w[2] | ((unsigned long long)w[3] << 32));
  } else if (sizeof(L) == 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k % (16 / sizeof(L))] = (long long)(int)w[k];
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) o[k % (16 / sizeof(L))] = (long long)((w[k >> 2] >> (8 * (k & 3))) & 0xffu);
  }
}
template <typename L>
__device__ __forceinline__ void lab_st16(L* p, const long long (&o)[16 / sizeof(L)]) {
  unsigned int w[4] = {0u, 0u, 0u, 0u};
  if (sizeof(L) == 8) {
    w[0] = (unsigned int)o[0]; w[1] = (unsigned int)((unsigned long long)o[0] >> 32);
    w[2] = (unsigned int)o[1 % (16 / sizeof(L))]; w[3] = (unsigned int)((unsigned long long)o[1 % (16 / sizeof(L))] >> 32);
  } else if (sizeof(L) == 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = (unsigned int)o[k % (16 / sizeof(L))];
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k >> 2] |= ((unsigned int)o[k % (16 / sizeof(L))] & 0xffu) << (8 * (k & 3));
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// natural log through the MUFU lg2 (abs error ~1e-7 for arguments in [1e-8, 2])
__device__ __forceinline__ float fast_log(float x) { return lg2(x) * kLn2; }

// e^x and sigmoid(x) from three MUFU ops (see sig_exp3 below, which every kernel uses through this wrapper or
// directly): the sigmoid follows torch's own formula 1/(1+e^-x) INCLUDING its fp32 rounding of (1 + e^-x) -- the
// reference takes log(1 - s + eps) of that quantised value and breaks ties between equal sigmoids by index, so all
// kernels must round the same way.
struct SigExp {
  float v;  // e^x
  float s;  // sigmoid(x)
};
__device__ __forceinline__ void sig_exp3(float x, float& s, float& ex);
__device__ __forceinline__ SigExp sig_exp(float x) {
  SigExp r;
  sig_exp3(x, r.s, r.v);
  return r;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// N consecutive elements -> fp32.  Vector path needs (ptr) aligned to N*sizeof(T).
template <typename T, int N>
struct VecIO;
template <typename T>
struct VecIO<T, 1> {       // scalar form (kernels that fall back to one pixel per thread for very wide hierarchies)
  static __device__ __forceinline__ void load(const T* p, float (&o)[1]) { o[0] = to_f32<T>(__ldg(p)); }
  static __device__ __forceinline__ void store(T* p, const float (&o)[1]) { *p = from_f32<T>(o[0]); }
};

template <>
struct VecIO<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&o)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(o[0], o[1], o[2], o[3]));
  }
};
template <>
struct VecIO<float, 2> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[2]) {
    float2 v = __ldg(reinterpret_cast<const float2*>(p));
    o[0] = v.x; o[1] = v.y;
  }
  static __device__ __forceinline__ void store(float* p, const float (&o)[2]) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(o[0], o[1]));
  }
};
template <>
struct VecIO<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[4]) {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
    o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&o)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(o[2], o[3]);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&a);
    v.y = *reinterpret_cast<uint32_t*>(&b);
    __stcs(reinterpret_cast<uint2*>(p), v);
  }
};
template <>
struct VecIO<__nv_bfloat16, 2> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[2]) {
    uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
    o[0] = __uint_as_float(v << 16); o[1] = __uint_as_float(v & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&o)[2]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
    __stcs(reinterpret_cast<uint32_t*>(p), *reinterpret_cast<uint32_t*>(&a));
  }
};
template <>
struct VecIO<__half, 4> {
  static __device__ __forceinline__ void load(const __half* p, float (&o)[4]) {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    __half2 a = *reinterpret_cast<__half2*>(&v.x), b = *reinterpret_cast<__half2*>(&v.y);
    float2 fa = __half22float2(a), fb = __half22float2(b);
    o[0] = fa.x; o[1] = fa.y; o[2] = fb.x; o[3] = fb.y;
  }
  static __device__ __forceinline__ void store(__half* p, const float (&o)[4]) {
    __half2 a = __floats2half2_rn(o[0], o[1]), b = __floats2half2_rn(o[2], o[3]);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&a);
    v.y = *reinterpret_cast<uint32_t*>(&b);
    __stcs(reinterpret_cast<uint2*>(p), v);
  }
};
template <>
struct VecIO<__half, 2> {
  static __device__ __forceinline__ void load(const __half* p, float (&o)[2]) {
    uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
    float2 f = __half22float2(*reinterpret_cast<__half2*>(&v));
    o[0] = f.x; o[1] = f.y;
  }
  static __device__ __forceinline__ void store(__half* p, const float (&o)[2]) {
    __half2 a = __floats2half2_rn(o[0], o[1]);
    __stcs(reinterpret_cast<uint32_t*>(p), *reinterpret_cast<uint32_t*>(&a));
  }
};


template <>
struct VecIO<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[2 * k] = __uint_as_float(w[k] << 16);
      o[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&o)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat162 a = __floats2bfloat162_rn(o[2 * k], o[2 * k + 1]);
      w[k] = *reinterpret_cast<uint32_t*>(&a);
    }
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(w[0], w[1], w[2], w[3]));
  }
};
template <>
struct VecIO<__half, 8> {
  static __device__ __forceinline__ void load(const __half* p, float (&o)[8]) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = __half22float2(*reinterpret_cast<__half2*>(&w[k]));
      o[2 * k] = f.x;
      o[2 * k + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__half* p, const float (&o)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __half2 a = __floats2half2_rn(o[2 * k], o[2 * k + 1]);
      w[k] = *reinterpret_cast<uint32_t*>(&a);
    }
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(w[0], w[1], w[2], w[3]));
  }
};

// Guarded N-element load/store: vector when `vec_ok` and fully in range, scalar otherwise.
template <typename T, int N>
__device__ __forceinline__ void load_n(const T* base, long idx, long limit, bool vec_ok, float (&o)[N]) {
  if (vec_ok && idx + N <= limit) {
    VecIO<T, N>::load(base + idx, o);
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) o[k] = (idx + k < limit) ? to_f32<T>(base[idx + k]) : 0.0f;
  }
}
template <typename T, int N>
__device__ __forceinline__ void store_n(T* base, long idx, long limit, bool vec_ok, const float (&o)[N]) {
  if (vec_ok && idx + N <= limit) {
    VecIO<T, N>::store(base + idx, o);
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k)
      if (idx + k < limit) base[idx + k] = from_f32<T>(o[k]);
  }
}

// ---- asynchronous global -> shared copies (LDGSTS): prefetch without holding registers -------------
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 4 consecutive elements of T staged by cp.async into a 16-byte slot -> fp32
template <typename T>
__device__ __forceinline__ void cp_async_vec4(void* slot, const T* g) {
  if (sizeof(T) == 4) cp_async_16(slot, g); else cp_async_8(slot, g);
}
template <typename T>
__device__ __forceinline__ void staged_vec4(const void* slot, float (&o)[4]);
template <>
__device__ __forceinline__ void staged_vec4<float>(const void* slot, float (&o)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(slot);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void staged_vec4<__nv_bfloat16>(const void* slot, float (&o)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(slot);
  o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
  o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
}
template <>
__device__ __forceinline__ void staged_vec4<__half>(const void* slot, float (&o)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(slot);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
// one element of T staged as the aligned 4-byte word that contains it
template <typename T>
__device__ __forceinline__ void cp_async_elem(void* slot, const T* base, long idx) {
  if (sizeof(T) == 4) cp_async_4(slot, base + idx); else cp_async_4(slot, base + (idx & ~1L));
}
template <typename T>
__device__ __forceinline__ float staged_elem(const void* slot, long idx);
template <>
__device__ __forceinline__ float staged_elem<float>(const void* slot, long) { return *reinterpret_cast<const float*>(slot); }
template <>
__device__ __forceinline__ float staged_elem<__nv_bfloat16>(const void* slot, long idx) {
  const unsigned int w = *reinterpret_cast<const unsigned int*>(slot);
  return __uint_as_float((idx & 1) ? (w & 0xffff0000u) : (w << 16));
}
template <>
__device__ __forceinline__ float staged_elem<__half>(const void* slot, long idx) {
  const unsigned int w = *reinterpret_cast<const unsigned int*>(slot);
  const __half2 h2 = *reinterpret_cast<const __half2*>(&w);
  return (idx & 1) ? __high2float(h2) : __low2float(h2);
}

// named barriers (ids 1..15; id 0 is __syncthreads): producer/consumer hand-over of shared-memory buffers.
// `n` = number of threads that take part (arrivers + waiters), a multiple of 32.
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// shared-memory mbarriers: data hand-over between warps without making the waiters rendezvous with each other
// (a named barrier synchronises everyone who takes part; here only the dependency is waited for).
__device__ __forceinline__ void mbar_init(unsigned int addr, int count) {
  asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned int addr) {      // release: the thread's earlier stores are visible to waiters
  asm volatile("mbarrier.arrive.shared.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int addr, unsigned int parity) {   // acquire
  unsigned int ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// clamp that lets NaN through (FMNMX.NAN): NaN logits must poison the loss like they do in the reference
__device__ __forceinline__ float clamp_nan(float v, float lo, float hi) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "f"(lo));
  asm("min.NaN.f32 %0, %0, %1;" : "+f"(r) : "f"(hi));
  return r;
}

// sigmoid(x) with torch's own rounding of (1 + e^-x) (the reference takes log(1 - s + eps) of that
// quantised value) and e^x, from 3 MUFU ops.  |x| is clamped to ~80 so that sums of e^x stay finite.
//   W = e^-x ;  s = RN(1 / (1 + W)) (rcp + one Newton step) ;  e^x = 1 / W
__device__ __forceinline__ void sig_exp3(float x, float& s, float& ex) {
  const float t = clamp_nan(x * (-kLog2e), -115.0f, 115.0f);
  const float w = ex2(t);
  const float y = 1.0f + w;
  const float q0 = rcp(y);
  s = fmaf(q0, fmaf(-y, q0, 1.0f), q0);
  ex = rcp(w);
}
__device__ __forceinline__ float sig_only(float x) {
  const float t = clamp_nan(x * (-kLog2e), -115.0f, 115.0f);
  const float y = 1.0f + ex2(t);
  const float q0 = rcp(y);
  return fmaf(q0, fmaf(-y, q0, 1.0f), q0);
}
// any byte of v equal to zero?
__device__ __forceinline__ bool has_zero_byte(unsigned int v) { return ((v - 0x01010101u) & ~v & 0x80808080u) != 0u; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of K floats per thread; result valid in thread 0..K-1 of warp 0
// (element k in thread k).  `scratch` needs K * (blockDim.x/32) floats.
template <int K>
__device__ __forceinline__ float block_sum_k(float (&v)[K], float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float r = warp_sum(v[k]);
    if (lane == 0) scratch[k * nwarp + warp] = r;
  }
  __syncthreads();
  float out = 0.0f;
  if (threadIdx.x < K) {
    for (int w = 0; w < nwarp; ++w) out += scratch[threadIdx.x * nwarp + w];
  }
  __syncthreads();
  return out;
}

}  // namespace sh
