// Probe: when does ptxas keep a stencil weight in a uniform register (FFMA R, R, UR, R)?  Compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cubin -o ur_probe.cubin ur_probe.cu && cuobjdump -sass ur_probe.cubin
// Result (CUDA 12.9): only the __constant__ table indexed by blockIdx (k_const) gives LDCU + UR operands (100 of 100 FFMA);
// a global load from a block-uniform address (k_glob), a __shfl_sync broadcast (k_shfl) and redux.sync (CREDUX writes a UR,
// then copied back to a vector register) all end as three-register FFMAs.
#include <cuda_runtime.h>
__constant__ float cw[16384];
// variant 1: weights from global via uniform address
__global__ void k_glob(const float* __restrict__ w, const float* __restrict__ p, float* out) {
  float wl[25];
  const float* wb = w + blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 25; ++i) wl[i] = __ldg(wb + i);
  float acc[16]; 
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  float win[8];
#pragma unroll 1
  for (int it = 0; it < 100; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) win[i] = p[(it * 8 + i) * 128 + threadIdx.x];
#pragma unroll
    for (int d = 0; d < 25; ++d)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[(d % 4) * 4 + k] = fmaf(wl[d], win[k + d % 5], acc[(d % 4) * 4 + k]);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * 128 + threadIdx.x] = s;
}
// variant 2: weights from constant memory with uniform (blockIdx) index
__global__ void k_const(const float* __restrict__ p, float* out) {
  const float* wl = cw + blockIdx.x * 32;
  float acc[16]; 
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  float win[8];
#pragma unroll 1
  for (int it = 0; it < 100; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) win[i] = p[(it * 8 + i) * 128 + threadIdx.x];
#pragma unroll
    for (int d = 0; d < 25; ++d)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[(d % 4) * 4 + k] = fmaf(wl[d], win[k + d % 5], acc[(d % 4) * 4 + k]);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * 128 + threadIdx.x] = s;
}
// variant 3: shfl broadcast
__global__ void k_shfl(const float* __restrict__ w, const float* __restrict__ p, float* out) {
  float wl[25];
  const float* wb = w + blockIdx.x * 32;
  float mine = wb[threadIdx.x & 31];
#pragma unroll
  for (int i = 0; i < 25; ++i) wl[i] = __shfl_sync(0xffffffffu, mine, i);
  float acc[16]; 
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  float win[8];
#pragma unroll 1
  for (int it = 0; it < 100; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) win[i] = p[(it * 8 + i) * 128 + threadIdx.x];
#pragma unroll
    for (int d = 0; d < 25; ++d)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[(d % 4) * 4 + k] = fmaf(wl[d], win[k + d % 5], acc[(d % 4) * 4 + k]);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * 128 + threadIdx.x] = s;
}
