// Micro-benchmark of per-SM issue rates on sm_100a (used to budget the streaming kernels).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef unsigned long long u64;
#define NACC 16
#define ITERS 2048

template <int OP>
__global__ void __launch_bounds__(256) k(float* out, float seed, int iters) {
  float a[NACC];
  u64 p[NACC / 2];
  unsigned int u[NACC];
  __shared__ float4 sm[256 * 2];
  sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
  sm[threadIdx.x + 256] = make_float4(seed, seed, seed, seed);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NACC; ++i) { a[i] = seed + i + threadIdx.x; u[i] = __float_as_uint(a[i]); }
#pragma unroll
  for (int i = 0; i < NACC / 2; ++i) p[i] = ((u64)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
  const float c1 = seed * 1.0001f, c2 = seed * 0.5f;
  const u64 pc1 = ((u64)__float_as_uint(c1) << 32) | __float_as_uint(c1);
  const u64 pc2 = ((u64)__float_as_uint(c2) << 32) | __float_as_uint(c2);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
      if (OP == 1) { if (i < NACC / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc1), "l"(pc2)); }
      if (OP == 2) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
      if (OP == 3) { if (i < NACC / 2) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc1)); }
      if (OP == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 6) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 7) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % NACC]));
      if (OP == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) % NACC]));
      if (OP == 9) asm volatile("{.reg .pred q; setp.gt.f32 q, %0, %1; selp.f32 %0, %1, %0, q;}" : "+f"(a[i]) : "f"(a[(i + 3) % NACC]));
      if (OP == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % NACC]), "r"(u[(i + 2) % NACC]));
      if (OP == 11) { // mix: FFMA + IADD alternate
        if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
        else asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 2) % NACC]));
      }
      if (OP == 12) { // mix: FFMA2 + IADD alternate
        if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i / 2]) : "l"(pc1), "l"(pc2));
        else asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 2) % NACC]));
      }
      if (OP == 13) { // mix: 3 FFMA + 1 MUFU
        if ((i & 3) == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
      }
      if (OP == 14) { // LDS.128
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(&sm[(threadIdx.x + i) & 511])));
        a[i] += v.x;
      }
      if (OP == 15) { // LDS.32
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(reinterpret_cast<float*>(sm) + ((threadIdx.x + 33 * i) & 2047))));
        a[i] += v;
      }
      if (OP == 16) asm volatile("mul.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
      if (OP == 17) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) % NACC]), "f"(a[(i + 2) % NACC]));  // 3 distinct regs
      if (OP == 18) { if (i < NACC / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(p[(i + 1) % (NACC / 2)]), "l"(p[(i + 2) % (NACC / 2)])); }
      if (OP == 19) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(u[i]) : "r"(u[(i + 1) % NACC]));
      if (OP == 20) { // mix FFMA + FMNMX (alu pipe)
        if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
        else asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 2) % NACC]));
      }
      if (OP == 21) asm volatile("sub.f32 %0, %1, %0;" : "+f"(a[i]) : "f"(a[(i + 1) % NACC]));
      if (OP == 22) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "+r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) % NACC]));
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += a[i] + __uint_as_float(u[i]);
#pragma unroll
  for (int i = 0; i < NACC / 2; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  if (r == 123.456f) out[0] = r;
}

template <int OP>
void run(const char* name, int ops_per_iter) {
  float* out; cudaMalloc(&out, 4);
  const int blocks = 148 * 8;
  k<OP><<<blocks, 256>>>(out, 1.0f, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<blocks, 256>>>(out, 1.0f, ITERS);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double winst = (double)blocks * 8 * ITERS * ops_per_iter;
  // per SM per ns
  printf("%-28s %8.3f ms  %7.3f warp-inst/ns/SM  (=%.3f /clk @1.965GHz, per SMSP %.3f)\n", name, ms, winst / 148 / (ms * 1e6),
         winst / 148 / (ms * 1e6) / 1.965, winst / 148 / (ms * 1e6) / 1.965 / 4);
  cudaFree(out);
}

int main() {
  run<0>("FFMA (r,c,c)", NACC);
  run<17>("FFMA (3 regs)", NACC);
  run<1>("FFMA2 (r,c,c)", NACC / 2);
  run<18>("FFMA2 (3 regs)", NACC / 2);
  run<2>("FADD", NACC);
  run<21>("FSUB 2reg", NACC);
  run<3>("FADD2", NACC / 2);
  run<16>("FMUL", NACC);
  run<4>("MUFU.EX2", NACC);
  run<5>("MUFU.RCP", NACC);
  run<6>("MUFU.LG2", NACC);
  run<7>("IADD", NACC);
  run<8>("FMNMX", NACC);
  run<9>("FSETP+FSEL", 2 * NACC);
  run<10>("LOP3", NACC);
  run<19>("PRMT", NACC);
  run<22>("F2F.BF16X2", NACC);
  run<11>("mix FFMA+IADD", NACC);
  run<12>("mix FFMA2+IADD", NACC);
  run<20>("mix FFMA+FMNMX", NACC);
  run<13>("mix 3FFMA+1MUFU", NACC);
  run<14>("LDS.128 (+FADD)", 2 * NACC);
  run<15>("LDS.32 (+FADD)", 2 * NACC);
  return 0;
}
