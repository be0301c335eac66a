// tcgen05 micro-benchmark (sm_100a): how fast can the 5th-generation tensor cores run the SMALL tf32 tiles that the
// RMI contractions of this repository would need ("Tensor cores for the 9 x 9 Gram / the 5 x 5
// stencil")?  One CTA per SM; one thread issues `iters` back-to-back  tcgen05.mma.cta_group::1.kind::tf32  of shape
// M x N x 8 (A, B in shared memory, K-major, no swizzle; accumulator in TMEM), commits them to an mbarrier and the
// CTA reads the accumulator back with tcgen05.ld (`acc` = number of independent accumulators the MMAs rotate over).  A = B = 1 everywhere, so every accumulator element must equal
// 8 * iters whatever the core-matrix layout: a wrong instruction / descriptor encoding shows up as a wrong sum (or
// a launch error), never as a silently fast number.  Reported: SM cycles per MMA (clock64 around issue + commit +
// wait, minimum and mean over the CTAs) and the dense MAC rate it corresponds to.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE.  Core matrix = 8 rows x 16 bytes, stored as 128 contiguous
// bytes; LBO = distance of the two core matrices of one K = 8 (tf32) step, SBO = distance of consecutive 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;                      // descriptor version of sm_100
  return d;                                    // layout type (bits 61..63) = 0: no swizzle
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, dense
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128) k_probe(int M, int N, int iters, int nacc, float* out, long long* cycles, int* bad) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* A = reinterpret_cast<float*>(smem);                 // 128 rows x 8 tf32 = 4 KB
  float* Bm = reinterpret_cast<float*>(smem + 4096);         // up to 256 rows x 8 tf32 = 8 KB
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (4096 + 8192) / 4; i += 128) A[i] = 1.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of A / B -> tensor-core reads
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;

  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint64_t da = make_desc(smem_u32(A), 128, 256), db = make_desc(smem_u32(Bm), 128, 256);
    const uint32_t idesc = make_idesc(M, N);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t acc = i >= nacc;                       // `nacc` independent accumulators, round robin (N columns apart)
      const uint32_t td = tm + (uint32_t)((i & (nacc - 1)) * N);   // nacc is a power of two
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(td), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  {   // everyone waits for the commit (bounded: a mistake must be a launch error, not a hung GPU)
    uint32_t ok = 0, spins = 0;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
      if (!ok && ++spins > (1u << 22)) __trap();
    } while (!ok);
  }
  if (tid == 0) { t1 = clock64(); cycles[blockIdx.x] = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // accumulator: lanes (= rows of D) 32 * warp .. +31, columns 0 .. 15
  uint32_t v[16];
  const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const float want = 8.0f * (float)(iters / nacc);
  const int rows = M == 64 ? 64 : 128;       // M = 64 fills lanes 0..31 and 64..95 on some layouts: only check what we know
  int wrong = 0;
  if (M == 128 || tid < 32)
    for (int j = 0; j < (N < 16 ? N : 16); ++j) wrong += __uint_as_float(v[j]) != want;
  (void)rows;
  if (wrong) atomicAdd(bad, wrong);
  if (blockIdx.x == 0 && tid == 0) out[0] = __uint_as_float(v[0]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256));
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  const int nsm = prop.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
  float* out; long long* cyc; int* bad;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, nsm * sizeof(long long)); cudaMalloc(&bad, 4);
  const size_t smem = 4096 + 8192 + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int shapes[][3] = {{128, 8, 1}, {128, 16, 1}, {128, 16, 2}, {128, 16, 4}, {128, 16, 8}, {128, 32, 1}, {128, 64, 1}, {128, 64, 4},
                           {128, 128, 1}, {128, 128, 2}, {128, 256, 1}, {64, 16, 1}, {64, 64, 1}};
  const int iters = 2048;
  printf("SMs %d, clock attr %d kHz; tf32 M x N x 8, %d MMAs per CTA, one CTA per SM\n", nsm, clk_khz, iters);
  for (auto& s : shapes) {
    cudaMemset(bad, 0, 4);
    k_probe<<<nsm, 128, smem>>>(s[0], s[1], 16, s[2], out, cyc, bad);          // warm-up
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(bad, 0, 4);
    cudaEventRecord(e0);
    k_probe<<<nsm, 128, smem>>>(s[0], s[1], iters, s[2], out, cyc, bad);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("M=%d N=%d: %s\n", s[0], s[1], cudaGetErrorString(err)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(nsm);
    int hbad; float hout;
    cudaMemcpy(h.data(), cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&hout, out, 4, cudaMemcpyDeviceToHost);
    long long mn = h[0]; double mean = 0;
    for (auto c : h) { mn = c < mn ? c : mn; mean += (double)c / nsm; }
    const double cpm = (double)mn / iters;
    printf("M=%3d N=%3d acc=%d: %7.2f cycles/MMA (min CTA; mean %.2f), %7.1f dense MAC/clk/SM, kernel %.3f ms, D[0][0] = %.0f (want %d), wrong elements %d\n",
           s[0], s[1], s[2], cpm, mean / iters, (double)s[0] * s[1] * 8 / cpm, ms, hout, 8 * iters / s[2], hbad);
  }
  return 0;
}
