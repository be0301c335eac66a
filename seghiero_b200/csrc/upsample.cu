// Bilinear upsampling (align_corners=False) of the head's logits, its adjoint, and the two consumers that never
// need the full-resolution tensor at all: the aux-head cross entropy (train.py:309-313) and the validation /
// inference decode (train.py:350-385, infer.py:296-312).  SURVEY section 8f rows N1-N3.
//
// Arithmetic of F.interpolate(mode="bilinear", align_corners=False, size=(H, W)) as torch computes it on CUDA
// (aten/src/ATen/native/cuda/UpSample.cuh, area_pixel_compute_source_index):
//   scale = in / out (float);  src = scale * (dst + 0.5) - 0.5, clamped at 0;  i0 = (int)src;
//   i1 = i0 + (i0 < in - 1);  l1 = src - i0;  l0 = 1 - l1
//   val = h0 * (w0 * x[i0][j0] + w1 * x[i0][j1]) + h1 * (w0 * x[i1][j0] + w1 * x[i1][j1])
// and the result is rounded to the tensor's dtype (bf16 / fp16 logits stay 16-bit, like the reference's tensors).
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "tma.cuh"

namespace sh {

struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_src(int dst, float scale, int n_in) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  Lerp r;
  r.i0 = min((int)src, n_in - 1);
  r.i1 = r.i0 + (r.i0 < n_in - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.0f - r.l1;
  return r;
}
// w0 * a + w1 * b with a FIXED contraction (every kernel of this file must produce the same bits for the same pixel: the
// fused decode is tested bit for bit against upsample + decode)
__device__ __forceinline__ float lerp2(float a, float b, float w0, float w1) { return __fmaf_rn(w1, b, __fmul_rn(w0, a)); }
__device__ __forceinline__ float bilerp(float a, float b, float c, float d, float w0, float w1, float h0, float h1) {
  return lerp2(lerp2(a, b, w0, w1), lerp2(c, d, w0, w1), h0, h1);
}
// what the value becomes when torch stores it in a tensor of type T
template <typename T>
__device__ __forceinline__ float round_as(float v) { return to_f32<T>(from_f32<T>(v)); }

// ---------------------------------------------------------------------------------------------
// N1, first form: the full-resolution logits are produced by our own kernel (one write of the tensor the loss
// kernels stream), and the gradient comes back to the head's resolution through the exact adjoint.
// Thread = 4 consecutive output pixels of one row (one 16 / 8 byte store).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_upsample(const T* __restrict__ in, T* __restrict__ out, long planes, int h, int w,
                                                  int H, int W, int vec_ok) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const int W4 = (W + 3) >> 2;
  const long total = planes * H * W4;
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
    const int x4 = (int)(g % W4);
    const long r = g / W4;
    const int y = (int)(r % H);
    const long p = r / H;
    const Lerp ly = lerp_src(y, sy, h);
    const T* r0 = in + (p * h + ly.i0) * w;
    const T* r1 = in + (p * h + ly.i1) * w;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = min(4 * x4 + k, W - 1);
      const Lerp lx = lerp_src(x, sx, w);
      v[k] = bilerp(to_f32<T>(__ldg(r0 + lx.i0)), to_f32<T>(__ldg(r0 + lx.i1)), to_f32<T>(__ldg(r1 + lx.i0)),
                    to_f32<T>(__ldg(r1 + lx.i1)), lx.l0, lx.l1, ly.l0, ly.l1);
    }
    T* o = out + (p * H + y) * (long)W + 4 * x4;
    if (vec_ok) {
      VecIO<T, 4>::store(o, v);
    } else {
      for (int k = 0; k < 4 && 4 * x4 + k < W; ++k) o[k] = from_f32<T>(v[k]);
    }
  }
}

// Adjoint (what autograd does for F.interpolate, but as a gather: deterministic, no atomics): thread = one
// low-resolution pixel; the contributing output rows / columns are those whose (i0, i1) of the forward formula hit it.
template <typename T>
__global__ void __launch_bounds__(256) k_upsample_adjoint(const T* __restrict__ gout, T* __restrict__ gin, long planes,
                                                          int h, int w, int H, int W) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const float ry = (float)H / (float)h, rx = (float)W / (float)w;
  const long total = planes * h * w;
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
    const int X = (int)(g % w);
    const long r = g / w;
    const int Y = (int)(r % h);
    const long p = r / h;
    // candidate ranges (one pixel of slack on each side; the exact test is the forward formula itself)
    const int ylo = max(0, (int)floorf(((float)Y - 0.5f) * ry - 0.5f) - 1);
    const int yhi = min(H - 1, (int)ceilf(((float)Y + 1.5f) * ry - 0.5f) + 1);
    const int xlo = max(0, (int)floorf(((float)X - 0.5f) * rx - 0.5f) - 1);
    const int xhi = min(W - 1, (int)ceilf(((float)X + 1.5f) * rx - 0.5f) + 1);
    const T* gp = gout + p * (long)H * W;
    float acc = 0.f;
    for (int y = ylo; y <= yhi; ++y) {
      const Lerp ly = lerp_src(y, sy, h);
      const float wy = (ly.i0 == Y ? ly.l0 : 0.f) + (ly.i1 == Y ? ly.l1 : 0.f);
      if (wy == 0.f) continue;
      float rowacc = 0.f;
      for (int x = xlo; x <= xhi; ++x) {
        const Lerp lx = lerp_src(x, sx, w);
        const float wx = (lx.i0 == X ? lx.l0 : 0.f) + (lx.i1 == X ? lx.l1 : 0.f);
        if (wx != 0.f) rowacc = fmaf(wx, to_f32<T>(__ldg(gp + (long)y * W + x)), rowacc);
      }
      acc = fmaf(wy, rowacc, acc);
    }
    gin[g] = from_f32<T>(acc);
  }
}

// The head's geometry (out = 4 x in exactly) without index arithmetic: thread = 4 x 4 output block whose sources are
// the 3 x 3 low-resolution neighbourhood of (Y, X); weights are the constants (2j - 3) / 8 of the forward formula
// (scale = 0.25 is exact in fp32, so these are the very numbers lerp_src produces).  Blocks that touch row / column 0
// (clamped source index) take the generic formula.
template <typename T>
__device__ __forceinline__ void up4_load_at(const T* __restrict__ in, long p, int Y, int X, int h, int w, float (&v)[3][3]) {
  const T* base = in + p * (long)h * w;
  const int ym = max(Y - 1, 0), yp = min(Y + 1, h - 1), xm = max(X - 1, 0), xp = min(X + 1, w - 1);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const T* row = base + (long)(a == 0 ? ym : (a == 1 ? Y : yp)) * w;
    v[a][0] = to_f32<T>(__ldg(row + xm));
    v[a][1] = to_f32<T>(__ldg(row + X));
    v[a][2] = to_f32<T>(__ldg(row + xp));
  }
}
template <typename T>
__device__ __forceinline__ void up4_store(const T* __restrict__ in, T* __restrict__ out, const float (&v)[3][3], int Y, int X,
                                          long p, int h, int w) {
  const int H = 4 * h, W = 4 * w;
  T* ob = out + (p * H + 4 * Y) * (long)W + 4 * X;
  if (Y > 0 && X > 0) {
    const float l1[4] = {0.625f, 0.875f, 0.125f, 0.375f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o[k] = bilerp(v[j >> 1][k >> 1], v[j >> 1][(k >> 1) + 1], v[(j >> 1) + 1][k >> 1], v[(j >> 1) + 1][(k >> 1) + 1],
                      1.0f - l1[k], l1[k], 1.0f - l1[j], l1[j]);
      VecIO<T, 4>::store(ob + (long)j * W, o);
    }
  } else {         // row / column 0 of the low-resolution map: the clamped source index pairs (0, 1) with weights (1, 0)
    const T* base = in + p * (long)h * w;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const Lerp ly = lerp_src(4 * Y + j, 0.25f, h);
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const Lerp lx = lerp_src(4 * X + k, 0.25f, w);
        o[k] = bilerp(to_f32<T>(__ldg(base + (long)ly.i0 * w + lx.i0)), to_f32<T>(__ldg(base + (long)ly.i0 * w + lx.i1)),
                      to_f32<T>(__ldg(base + (long)ly.i1 * w + lx.i0)), to_f32<T>(__ldg(base + (long)ly.i1 * w + lx.i1)),
                      lx.l0, lx.l1, ly.l0, ly.l1);
      }
      VecIO<T, 4>::store(ob + (long)j * W, o);
    }
  }
}
// (two blocks per trip with all 18 loads in flight before the first store was measured slower: 0.48 vs 0.41 ms at
// config 3, 76 instead of 61 registers)
// CTA = 8 x 32 low-resolution pixels of one plane (a warp = 32 consecutive X of one row: every store instruction of
// the warp covers 512 contiguous bytes); grid = (tiles of a plane, planes), no 64-bit index arithmetic per thread.
template <typename T>
__global__ void __launch_bounds__(256) k_upsample4(const T* __restrict__ in, T* __restrict__ out, long planes, int h, int w,
                                                  int tiles_x) {
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int Y = ty * 8 + (threadIdx.x >> 5), X = tx * 32 + (threadIdx.x & 31);
  if (Y >= h || X >= w) return;
  for (long p = blockIdx.y; p < planes; p += gridDim.y) {
    float v[3][3];
    up4_load_at<T>(in, p, Y, X, h, w, v);
    up4_store<T>(in, out, v, Y, X, p, h, w);
  }
}

// Adjoint for out = 4 x in: thread = one low-resolution pixel, gathering its 8 x 8 output window with the separable
// constant weights {1,3,5,7,7,5,3,1} / 8; pixels on the low-resolution border (clamped sources) take the generic gather.
template <typename T>
__device__ __forceinline__ float adjoint_generic(const T* __restrict__ gp, int Y, int X, int h, int w, int H, int W) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const int ylo = max(0, 4 * Y - 3), yhi = min(H - 1, 4 * Y + 6), xlo = max(0, 4 * X - 3), xhi = min(W - 1, 4 * X + 6);
  float acc = 0.f;
  for (int y = ylo; y <= yhi; ++y) {
    const Lerp ly = lerp_src(y, sy, h);
    const float wy = (ly.i0 == Y ? ly.l0 : 0.f) + (ly.i1 == Y ? ly.l1 : 0.f);
    if (wy == 0.f) continue;
    float rowacc = 0.f;
    for (int x = xlo; x <= xhi; ++x) {
      const Lerp lx = lerp_src(x, sx, w);
      const float wx = (lx.i0 == X ? lx.l0 : 0.f) + (lx.i1 == X ? lx.l1 : 0.f);
      if (wx != 0.f) rowacc = fmaf(wx, to_f32<T>(__ldg(gp + (long)y * W + x)), rowacc);
    }
    acc = fmaf(wy, rowacc, acc);
  }
  return acc;
}

// CTA = 8 x 32 low-resolution pixels of one plane at a time (persistent over tiles): the 36 x 136 output pixels they
// gather from are staged in shared memory with cp.async (16 bytes per request, zero-filled outside the image; each
// gradient byte crosses HBM once), the tile after next is in flight while this one is gathered (two buffers);
// thread = one low-resolution pixel (3 x LDS.128 per window row, separable weights).
constexpr int ADJ_TH = 8, ADJ_TW = 32, ADJ_SR = 4 * ADJ_TH + 4, ADJ_SC = 4 * ADJ_TW + 8;
template <typename T>
__global__ void __launch_bounds__(ADJ_TH * ADJ_TW) k_upsample4_adjoint(const T* __restrict__ gout, T* __restrict__ gin,
                                                                      long planes, int h, int w, int tiles_x, int tiles_y) {
  __shared__ __align__(16) T tile[2][ADJ_SR][ADJ_SC];         // rows 4Y0-2 .. 4Y0+33, cols 4X0-4 .. 4X0+131
  const int H = 4 * h, W = 4 * w;
  const int tid = threadIdx.x;
  const long ntiles = planes * tiles_x * tiles_y;
  auto issue = [&](long t, int buf) {
    const int tx = (int)(t % tiles_x);
    const long r = t / tiles_x;
    const int ty = (int)(r % tiles_y);
    const long p = r / tiles_y;
    const T* gp = gout + p * (long)H * W;
    for (int e = tid; e < ADJ_SR * (ADJ_SC / 4); e += ADJ_TH * ADJ_TW) {
      const int rr = e / (ADJ_SC / 4), q = e - rr * (ADJ_SC / 4);
      const int y = 4 * ty * ADJ_TH - 2 + rr, x = 4 * tx * ADJ_TW - 4 + 4 * q;
      const bool in = y >= 0 && y < H && x >= 0 && x < W;
      const unsigned int dst = (unsigned int)__cvta_generic_to_shared(&tile[buf][rr][4 * q]);
      const T* src = in ? gp + (long)y * W + x : gout;
      const int nbytes = in ? 4 * (int)sizeof(T) : 0;        // src-size 0: the slot is zero-filled
      if (sizeof(T) == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes));
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(nbytes));
    }
    cp_async_commit();
  };
  long t = blockIdx.x;
  if (t < ntiles) issue(t, 0);
  for (int it = 0; t < ntiles; t += gridDim.x, ++it) {
    const int buf = it & 1;
    const bool more = t + gridDim.x < ntiles;
    if (more) issue(t + gridDim.x, buf ^ 1);
    if (more) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const int tx = (int)(t % tiles_x);
    const long r = t / tiles_x;
    const int ty = (int)(r % tiles_y);
    const long p = r / tiles_y;
    const int ly = tid / ADJ_TW, lx = tid - ly * ADJ_TW;
    const int Y = ty * ADJ_TH + ly, X = tx * ADJ_TW + lx;
    if (Y < h && X < w) {
      float acc = 0.f;
      if (h >= 2 && w >= 2) {
        // weights of the 8 window rows / columns 4Y-2 .. 4Y+5 on low-resolution row Y: the interior pattern, or the clamped
        // patterns of the first / last row (sources outside the map fold onto it; rows outside the image weigh 0)
        float wy[8], wx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float in = i < 4 ? 0.125f + 0.25f * i : 0.875f - 0.25f * (i - 4);
          wy[i] = Y == 0 ? (i < 2 ? 0.f : (i < 4 ? 1.f : in)) : (Y == h - 1 ? (i >= 6 ? 0.f : (i >= 4 ? 1.f : in)) : in);
          wx[i] = X == 0 ? (i < 2 ? 0.f : (i < 4 ? 1.f : in)) : (X == w - 1 ? (i >= 6 ? 0.f : (i >= 4 ? 1.f : in)) : in);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a[4], b4[4], c[4];
          staged_vec4<T>(&tile[buf][4 * ly + i][4 * lx], a);
          staged_vec4<T>(&tile[buf][4 * ly + i][4 * lx + 4], b4);
          staged_vec4<T>(&tile[buf][4 * ly + i][4 * lx + 8], c);
          // window columns 4X-2 .. 4X+5 = tile columns 4lx+2 .. 4lx+9
          float rs = wx[0] * a[2];
          rs = fmaf(wx[1], a[3], rs); rs = fmaf(wx[2], b4[0], rs); rs = fmaf(wx[3], b4[1], rs);
          rs = fmaf(wx[4], b4[2], rs); rs = fmaf(wx[5], b4[3], rs); rs = fmaf(wx[6], c[0], rs); rs = fmaf(wx[7], c[1], rs);
          acc = fmaf(wy[i], rs, acc);
        }
      } else {
        acc = adjoint_generic<T>(gout + p * (long)H * W, Y, X, h, w, H, W);
      }
      gin[(p * h + Y) * (long)w + X] = from_f32<T>(acc);
    }
    __syncthreads();           // everyone is done with this buffer before the next trip refills it
  }
}

// The same gather with the copy engine doing the staging: one thread issues the whole 36-row tile (zero-filled outside
// the image) as ONE cp.async.bulk.tensor box per trip, ADJ_NST - 1 tiles ahead, an mbarrier per stage hands it over;
// thread = TWO vertically adjacent low-resolution pixels (their windows share 4 of 12 rows: 18 instead of 24 LDS.128
// per output).  The cp.async form above spends a seventh of its instructions on staging addresses and runs at 74 %
// issue utilisation; it stays for maps that cannot be tensor-mapped (W * sizeof(T) not a multiple of 16 bytes).
constexpr int ADJ_NST = 3, ADJ_NT2 = ADJ_TH / 2 * ADJ_TW;
template <typename T> struct AdjBox {
  static constexpr int XO = sizeof(T) == 4 ? 4 : 8;                 // the box starts XO columns left of the tile: 16-byte aligned start
  static constexpr int SC = 4 * ADJ_TW + 2 * XO;
  static constexpr int BYTES = ADJ_SR * SC * (int)sizeof(T);
  static constexpr int STAGE = ((BYTES + 127) / 128) * 128;
};
template <typename T>
__global__ void __launch_bounds__(ADJ_NT2) k_upsample4_adjoint_tma(const __grid_constant__ CUtensorMap map_g,
                                                                   const T* __restrict__ gout, T* __restrict__ gin, long planes,
                                                                   int h, int w, int tiles_x, int tiles_y) {
  extern __shared__ __align__(128) unsigned char adj_smem[];
  constexpr int XO = AdjBox<T>::XO, SC = AdjBox<T>::SC;
  __shared__ __align__(8) unsigned long long s_mbar[ADJ_NST];
  const int H = 4 * h, W = 4 * w;
  const int tid = threadIdx.x;
  const long ntiles = planes * tiles_x * tiles_y;
  const unsigned int smem_s = (unsigned int)__cvta_generic_to_shared(adj_smem);
  const unsigned int mbar_s = (unsigned int)__cvta_generic_to_shared(s_mbar);
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < ADJ_NST; ++q) mbar_init(mbar_s + 8 * q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](long t, int stage) {       // tid 0 only
    const int tx = (int)(t % tiles_x);
    const long r = t / tiles_x;
    const int ty = (int)(r % tiles_y);
    const int p = (int)(r / tiles_y);
    mbar_expect_tx(mbar_s + 8 * stage, (unsigned int)AdjBox<T>::BYTES);
    tma_load_3d(smem_s + stage * AdjBox<T>::STAGE, &map_g, 4 * tx * ADJ_TW - XO, 4 * ty * ADJ_TH - 2, p, mbar_s + 8 * stage);
  };
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < ADJ_NST - 1; ++q)
      if (blockIdx.x + (long)q * gridDim.x < ntiles) issue(blockIdx.x + (long)q * gridDim.x, q);
  }
  const int ly2 = tid / ADJ_TW, lx = tid - ly2 * ADJ_TW;
  int it = 0;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int stage = it % ADJ_NST;
    if (tid == 0) {        // the stage consumed in the previous trip is free (barrier at the end of that trip)
      const long tn = t + (long)(ADJ_NST - 1) * gridDim.x;
      if (tn < ntiles) issue(tn, (it + ADJ_NST - 1) % ADJ_NST);
    }
    mbar_wait_or_trap(mbar_s + 8 * stage, (unsigned int)(it / ADJ_NST) & 1u);
    const T* tile = reinterpret_cast<const T*>(adj_smem + stage * AdjBox<T>::STAGE);
    const int tx = (int)(t % tiles_x);
    const long r = t / tiles_x;
    const int ty = (int)(r % tiles_y);
    const long p = r / tiles_y;
    const int Y = ty * ADJ_TH + 2 * ly2, X = tx * ADJ_TW + lx;
    if (Y < h && X < w) {
      float accA = 0.f, accB = 0.f;
      if (h >= 2 && w >= 2) {
        float wyA[8], wyB[8], wx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float in = i < 4 ? 0.125f + 0.25f * i : 0.875f - 0.25f * (i - 4);
          wyA[i] = Y == 0 ? (i < 2 ? 0.f : (i < 4 ? 1.f : in)) : (Y == h - 1 ? (i >= 6 ? 0.f : (i >= 4 ? 1.f : in)) : in);
          wyB[i] = Y + 1 == h - 1 ? (i >= 6 ? 0.f : (i >= 4 ? 1.f : in)) : in;        // Y + 1 >= 1: never the first row
          wx[i] = X == 0 ? (i < 2 ? 0.f : (i < 4 ? 1.f : in)) : (X == w - 1 ? (i >= 6 ? 0.f : (i >= 4 ? 1.f : in)) : in);
        }
        const T* base = tile + (8 * ly2) * SC + 4 * lx + XO - 4;      // window columns 4X-2 .. 4X+5 = base[2] .. base[9]
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          float a[4], b4[4], c[4];
          staged_vec4<T>(base + i * SC, a);
          staged_vec4<T>(base + i * SC + 4, b4);
          staged_vec4<T>(base + i * SC + 8, c);
          float rs = wx[0] * a[2];
          rs = fmaf(wx[1], a[3], rs); rs = fmaf(wx[2], b4[0], rs); rs = fmaf(wx[3], b4[1], rs);
          rs = fmaf(wx[4], b4[2], rs); rs = fmaf(wx[5], b4[3], rs); rs = fmaf(wx[6], c[0], rs); rs = fmaf(wx[7], c[1], rs);
          if (i < 8) accA = fmaf(wyA[i], rs, accA);
          if (i >= 4) accB = fmaf(wyB[i - 4], rs, accB);
        }
      } else {
        accA = adjoint_generic<T>(gout + p * (long)H * W, Y, X, h, w, H, W);
        if (Y + 1 < h) accB = adjoint_generic<T>(gout + p * (long)H * W, Y + 1, X, h, w, H, W);
      }
      gin[(p * h + Y) * (long)w + X] = from_f32<T>(accA);
      if (Y + 1 < h) gin[(p * h + Y + 1) * (long)w + X] = from_f32<T>(accB);
    }
    __syncthreads();           // everyone is done with this stage before it is refilled
  }
}

// ---------------------------------------------------------------------------------------------
// N2: aux-head loss  nn.CrossEntropyLoss(ignore_index=255)(F.interpolate(aux_logits, (H, W)), label)  and its
// gradient w.r.t. the LOW-resolution aux logits, fused: nothing of size [B, C, H, W] is ever written.
// CTA = 32 x 8 output tile; thread = one output pixel, channels in a loop (online log-sum-exp over the interpolated
// logits, then softmax - one-hot scattered to the <= 4 source pixels through shared-memory accumulators that are
// flushed to the low-resolution gradient once per CTA).  `gin` must be zeroed by the caller.
// ---------------------------------------------------------------------------------------------
constexpr int AUX_TW = 32, AUX_TH = 8;
// The gradient is scattered with INTEGER atomics (fixed point, 2^-40): the sum does not depend on the order in which warps
// and CTAs arrive, so the aux gradient is bitwise reproducible.  |softmax - onehot| <= 1 per output pixel: no overflow.
constexpr float kAuxFix = 1099511627776.0f;      // 2^40
__device__ __forceinline__ unsigned long long aux_fix(float v) { return (unsigned long long)__float2ll_rn(v * kAuxFix); }

template <typename T, typename L>
__global__ void __launch_bounds__(AUX_TW * AUX_TH) k_aux_ce(const T* __restrict__ in, const L* __restrict__ label,
                                                           unsigned long long* __restrict__ gin, int B, int C, int h, int w, int H,
                                                           int W, int fh, int fw, double* __restrict__ sums) {
  extern __shared__ unsigned long long s_acc[];       // [C][fh][fw] gradient footprint of the tile at low resolution (fixed point)
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const int tiles_x = (W + AUX_TW - 1) / AUX_TW, tiles_y = (H + AUX_TH - 1) / AUX_TH;
  const int tile = blockIdx.x % (tiles_x * tiles_y), b = blockIdx.x / (tiles_x * tiles_y);
  const int ty0 = (tile / tiles_x) * AUX_TH, tx0 = (tile % tiles_x) * AUX_TW;
  const int tid = threadIdx.x;
  const int fsz = C * fh * fw;
  for (int i = tid; i < fsz; i += AUX_TW * AUX_TH) s_acc[i] = 0ull;
  // low-resolution origin of the tile's footprint
  const int Y0 = lerp_src(ty0, sy, h).i0, X0 = lerp_src(tx0, sx, w).i0;
  __syncthreads();
  const int x = tx0 + (tid & (AUX_TW - 1)), y = ty0 + tid / AUX_TW;
  float loss = 0.f;
  int nvalid = 0;
  if (x < W && y < H) {
    const long long t = lab_ld(label, ((long)b * H + y) * W + x);
    if (t != SH_IGNORE && t >= 0 && t < C) {
      const Lerp ly = lerp_src(y, sy, h), lx = lerp_src(x, sx, w);
      const T* base = in + (long)b * C * h * w;
      const int o00 = ly.i0 * w + lx.i0, o01 = ly.i0 * w + lx.i1, o10 = ly.i1 * w + lx.i0, o11 = ly.i1 * w + lx.i1;
      // pass 1: max and sum of exp (two sweeps over <= a few dozen channels, all L1 hits)
      float mx = -INFINITY, xt = 0.f;
      for (int c = 0; c < C; ++c) {
        const T* pc = base + (long)c * h * w;
        const float v = round_as<T>(bilerp(to_f32<T>(__ldg(pc + o00)), to_f32<T>(__ldg(pc + o01)),
                                           to_f32<T>(__ldg(pc + o10)), to_f32<T>(__ldg(pc + o11)), lx.l0, lx.l1, ly.l0, ly.l1));
        mx = fmaxf(mx, v);
        if (c == (int)t) xt = v;
      }
      float se = 0.f;
      for (int c = 0; c < C; ++c) {
        const T* pc = base + (long)c * h * w;
        const float v = round_as<T>(bilerp(to_f32<T>(__ldg(pc + o00)), to_f32<T>(__ldg(pc + o01)),
                                           to_f32<T>(__ldg(pc + o10)), to_f32<T>(__ldg(pc + o11)), lx.l0, lx.l1, ly.l0, ly.l1));
        se += ex2((v - mx) * kLog2e);
      }
      loss = (mx - xt) + lg2(se) * kLn2;
      nvalid = 1;
      if (gin != nullptr) {
        const float inv = rcp(se);
        const int a00 = ((ly.i0 - Y0) * fw + (lx.i0 - X0)), a01 = ((ly.i0 - Y0) * fw + (lx.i1 - X0));
        const int a10 = ((ly.i1 - Y0) * fw + (lx.i0 - X0)), a11 = ((ly.i1 - Y0) * fw + (lx.i1 - X0));
        const float w00 = ly.l0 * lx.l0, w01 = ly.l0 * lx.l1, w10 = ly.l1 * lx.l0, w11 = ly.l1 * lx.l1;
        for (int c = 0; c < C; ++c) {
          const T* pc = base + (long)c * h * w;
          const float v = round_as<T>(bilerp(to_f32<T>(__ldg(pc + o00)), to_f32<T>(__ldg(pc + o01)),
                                             to_f32<T>(__ldg(pc + o10)), to_f32<T>(__ldg(pc + o11)), lx.l0, lx.l1, ly.l0, ly.l1));
          const float gq = ex2((v - mx) * kLog2e) * inv - (c == (int)t ? 1.f : 0.f);
          unsigned long long* a = s_acc + c * fh * fw;
          atomicAdd(a + a00, aux_fix(w00 * gq));
          if (w01 != 0.f) atomicAdd(a + a01, aux_fix(w01 * gq));
          if (w10 != 0.f) atomicAdd(a + a10, aux_fix(w10 * gq));
          if (w11 != 0.f) atomicAdd(a + a11, aux_fix(w11 * gq));
        }
      }
    } else if (t != SH_IGNORE) {
      nvalid = 0x10000;        // out-of-range label: the reference's cross_entropy asserts on the device
    }
  }
  // CTA totals -> global (fp64 atomics: a few thousand per launch)
  __shared__ float s_l[AUX_TW * AUX_TH / 32];
  __shared__ int s_n[AUX_TW * AUX_TH / 32];
  loss = warp_sum(loss);
  nvalid = __reduce_add_sync(0xffffffffu, nvalid);
  if ((tid & 31) == 0) { s_l[tid >> 5] = loss; s_n[tid >> 5] = nvalid; }
  __syncthreads();
  if (tid == 0) {
    float l = 0.f;
    int n = 0;
    for (int q = 0; q < AUX_TW * AUX_TH / 32; ++q) { l += s_l[q]; n += s_n[q]; }
    if (l != 0.f) atomicAdd(sums, (double)l);
    if (n & 0xffff) atomicAdd(sums + 1, (double)(n & 0xffff));
    if (n >> 16) atomicAdd(sums + 2, 1.0);
  }
  if (gin != nullptr) {
    for (int i = tid; i < fsz; i += AUX_TW * AUX_TH) {
      const unsigned long long v = s_acc[i];
      if (v == 0ull) continue;
      const int c = i / (fh * fw), rem = i - c * fh * fw, fy = rem / fw, fx = rem - fy * fw;
      const int Y = Y0 + fy, X = X0 + fx;
      if (Y < h && X < w) atomicAdd(gin + (((long)b * C + c) * h + Y) * w + X, v);
    }
  }
}

// Integer scale k = 2 * PX (the aux head's x16, and x8 / x4): thread = a strip of PX consecutive output pixels of one row,
// aligned to PX.  Such a strip lies inside one half of a k-block, so all its pixels interpolate between the SAME two
// source columns (and, being one row, the same two source rows): the 4 source values are loaded once per channel and
// strip, interpolated vertically once, and the gradient of the strip is summed in registers before it goes to the
// shared-memory footprint -- 4 atomics per (strip, channel), at most 2 lanes of a warp on one address.  (With a thread
// per output pixel every lane of a warp hits the same 2 - 3 accumulators: 32-way serialised atomics, 10 ms at config 3.)
// CTA = 8 rows x 32 strips.  The interpolation is contracted vertically first here (the generic kernel follows ATen's
// horizontal-first order): same value up to fp32 rounding, parity is to the loss / gradient tolerance, not bitwise.
template <typename T, typename L, int PX>
__global__ void __launch_bounds__(256) k_aux_ce_strip(const T* __restrict__ in, const L* __restrict__ label,
                                                      unsigned long long* __restrict__ gin, int B, int C, int h, int w, int H, int W,
                                                      int fh, int fw, double* __restrict__ sums) {
  extern __shared__ unsigned long long s_acc[];       // [C][fh][fw], fixed point
  constexpr int TWS = 32 * PX, THS = 8;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const int tiles_x = (W + TWS - 1) / TWS, tiles_y = (H + THS - 1) / THS;
  const int tile = blockIdx.x % (tiles_x * tiles_y), b = blockIdx.x / (tiles_x * tiles_y);
  const int ty0 = (tile / tiles_x) * THS, tx0 = (tile % tiles_x) * TWS;
  const int tid = threadIdx.x;
  const int fsz = C * fh * fw;
  for (int i = tid; i < fsz; i += 256) s_acc[i] = 0ull;
  const int Y0 = lerp_src(ty0, sy, h).i0, X0 = lerp_src(tx0, sx, w).i0;
  __syncthreads();
  const int x0 = tx0 + (tid & 31) * PX, y = ty0 + (tid >> 5);
  float loss = 0.f;
  int nvalid = 0;
  if (x0 < W && y < H) {
    const Lerp ly = lerp_src(y, sy, h), lxs = lerp_src(x0, sx, w);
    float l0[PX], l1[PX], mx[PX], xt[PX], se[PX];
    int t[PX];
    bool any = false;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      const long long tt = lab_ld(label, ((long)b * H + y) * W + x0 + p);
      t[p] = -1;
      if (tt != SH_IGNORE) {
        if (tt >= 0 && tt < C) { t[p] = (int)tt; any = true; }
        else nvalid = 0x10000;                     // out-of-range label: the reference's cross_entropy asserts on the device
      }
      l1[p] = lerp_src(x0 + p, sx, w).l1;          // the strip shares i0 / i1; only the weight moves
      l0[p] = 1.0f - l1[p];
      mx[p] = -INFINITY; xt[p] = 0.f; se[p] = 0.f;
    }
    if (any) {
      const T* base = in + (long)b * C * h * w;
      const int o00 = ly.i0 * w + lxs.i0, o01 = ly.i0 * w + lxs.i1, o10 = ly.i1 * w + lxs.i0, o11 = ly.i1 * w + lxs.i1;
      const long hw = (long)h * w;
      // sweep 1: online log-sum-exp (running max m and sum of e^(v - m); ONE ex2 per value: e = 2^-|v - m| serves both
      // the case v <= m (sum += e) and v > m (sum = sum * e + 1)) and the target's logit; sweep 2: gradient.  The 4 source
      // values per channel are L1 / L2 hits.
#pragma unroll 1
      for (int c = 0; c < C; ++c) {
        const T* pc = base + c * hw;
        const float A = lerp2(to_f32<T>(__ldg(pc + o00)), to_f32<T>(__ldg(pc + o10)), ly.l0, ly.l1);
        const float Bv = lerp2(to_f32<T>(__ldg(pc + o01)), to_f32<T>(__ldg(pc + o11)), ly.l0, ly.l1);
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const float v = round_as<T>(lerp2(A, Bv, l0[p], l1[p]));
          const float dlt = v - mx[p];
          const float e = ex2(-fabsf(dlt) * kLog2e);           // mx starts at -inf: e = 0, sum = 0 * 0 + 1
          se[p] = dlt > 0.f ? fmaf(se[p], e, 1.0f) : se[p] + e;
          mx[p] = fmaxf(mx[p], v);
          xt[p] = c == t[p] ? v : xt[p];
        }
      }
      float inv[PX];
#pragma unroll
      for (int p = 0; p < PX; ++p) {
        const bool ok = t[p] >= 0;
        if (ok) { loss += (mx[p] - xt[p]) + lg2(se[p]) * kLn2; ++nvalid; }
        inv[p] = ok ? rcp(se[p]) : 0.f;            // ignored pixels: no gradient
      }
      if (gin != nullptr) {
        const int a00 = (ly.i0 - Y0) * fw + (lxs.i0 - X0), a01 = (ly.i0 - Y0) * fw + (lxs.i1 - X0);
        const int a10 = (ly.i1 - Y0) * fw + (lxs.i0 - X0), a11 = (ly.i1 - Y0) * fw + (lxs.i1 - X0);
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
          const T* pc = base + c * hw;
          const float A = lerp2(to_f32<T>(__ldg(pc + o00)), to_f32<T>(__ldg(pc + o10)), ly.l0, ly.l1);
          const float Bv = lerp2(to_f32<T>(__ldg(pc + o01)), to_f32<T>(__ldg(pc + o11)), ly.l0, ly.l1);
          float gA = 0.f, gB = 0.f;
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            const float v = round_as<T>(lerp2(A, Bv, l0[p], l1[p]));
            const float gq = fmaf(ex2((v - mx[p]) * kLog2e), inv[p], c == t[p] ? -1.f : 0.f);
            gA = fmaf(l0[p], gq, gA);
            gB = fmaf(l1[p], gq, gB);
          }
          unsigned long long* a = s_acc + c * fh * fw;
          if (gA != 0.f) {
            atomicAdd(a + a00, aux_fix(ly.l0 * gA));
            if (ly.l1 != 0.f) atomicAdd(a + a10, aux_fix(ly.l1 * gA));
          }
          if (gB != 0.f) {
            atomicAdd(a + a01, aux_fix(ly.l0 * gB));
            if (ly.l1 != 0.f) atomicAdd(a + a11, aux_fix(ly.l1 * gB));
          }
        }
      }
    }
  }
  __shared__ float s_l[8];
  __shared__ int s_n[8];
  loss = warp_sum(loss);
  nvalid = __reduce_add_sync(0xffffffffu, nvalid);
  if ((tid & 31) == 0) { s_l[tid >> 5] = loss; s_n[tid >> 5] = nvalid; }
  __syncthreads();
  if (tid == 0) {
    float l = 0.f;
    int n = 0;
    for (int q = 0; q < 8; ++q) { l += s_l[q]; n += s_n[q]; }
    if (l != 0.f) atomicAdd(sums, (double)l);
    if (n & 0xffff) atomicAdd(sums + 1, (double)(n & 0xffff));
    if (n >> 16) atomicAdd(sums + 2, 1.0);
  }
  if (gin != nullptr) {
    for (int i = tid; i < fsz; i += 256) {
      const unsigned long long v = s_acc[i];
      if (v == 0ull) continue;
      const int c = i / (fh * fw), rem = i - c * fh * fw, fy = rem / fw, fx = rem - fy * fw;
      const int Y = Y0 + fy, X = X0 + fx;
      if (Y < h && X < w) atomicAdd(gin + (((long)b * C + c) * h + Y) * w + X, v);
    }
  }
}

// loss = sums[0] / sums[1] (NaN when nothing is valid, as torch), gradient scale = gscale / sums[1]
template <typename T>
__global__ void __launch_bounds__(256) k_aux_finish(const unsigned long long* __restrict__ gacc, T* __restrict__ gin, long n,
                                                    const double* __restrict__ sums, const float* __restrict__ gscale,
                                                    float* __restrict__ out) {
  const double nv = sums[1];
  if (blockIdx.x == 0 && threadIdx.x == 0 && out != nullptr) {
    double l = sums[0] / nv;
    if (sums[2] != 0.0) l = __longlong_as_double(0x7ff8000000000000LL);
    out[0] = (float)l;
  }
  if (gin == nullptr) return;
  const double s = (gscale ? (double)*gscale : 1.0) / nv / (double)kAuxFix;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    gin[i] = from_f32<T>((float)((double)(long long)gacc[i] * s));
}

// ---------------------------------------------------------------------------------------------
// N3: decode straight from the head's H/4 logits (out = 4 x in exactly): per-level argmax of the interpolated
// logits + fine pixel-accuracy counts, uint8 or int64 predictions.  Thread = 4 x 4 output block = the 16 pixels whose
// sources are the 3 x 3 low-resolution neighbourhood around (Y, X): 9 loads per channel for 16 outputs.
//   output row 4Y + j, j = 0..3: src = Y + (2j - 3) / 8  ->  rows (Y-1, Y) for j < 2, (Y, Y+1) for j >= 2, clamped
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_args4(long long* dst, const int (&a)[4]) {
  __stcs(reinterpret_cast<longlong2*>(dst), make_longlong2(a[0], a[1]));
  __stcs(reinterpret_cast<longlong2*>(dst) + 1, make_longlong2(a[2], a[3]));
}
__device__ __forceinline__ void store_args4(unsigned char* dst, const int (&a)[4]) {
  *reinterpret_cast<unsigned int*>(dst) = (unsigned)a[0] | ((unsigned)a[1] << 8) | ((unsigned)a[2] << 16) | ((unsigned)a[3] << 24);
}
__device__ __forceinline__ float pick3(const float (&v)[3], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]); }

// first-max argmax update without branches: take when val > best, or val is NaN and best is not (torch.argmax counts NaN
// as the maximum and keeps the first one).  best starts at -inf with arg 0, so the first channel needs no special case.
__device__ __forceinline__ void argmax_step(float val, int idx, float& best, int& arg) {
  const bool take = (best == best) && !(val <= best);
  best = take ? val : best;
  arg = take ? idx : arg;
}

// 16-bit logits: the interpolated values are rounded to the logits' type anyway (as torch's tensors are), so two of them
// travel as one packed word and the first-max update is two HSET2 masks and three bitwise selects per PAIR (see
// targets_decode.cu): take the new channel when the running best is not NaN and NOT (value <= best).
template <typename T> struct Pack2;
template <> struct Pack2<__nv_bfloat16> {
  static constexpr unsigned int NEG_INF = 0xff80ff80u;
  static __device__ __forceinline__ unsigned int pack(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const unsigned int*>(&v);
  }
  static __device__ __forceinline__ unsigned int take(unsigned int v, unsigned int best) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v), b = *reinterpret_cast<const __nv_bfloat162*>(&best);
    return ~__hle2_mask(a, b) & __heq2_mask(b, b);
  }
};
template <> struct Pack2<__half> {
  static constexpr unsigned int NEG_INF = 0xfc00fc00u;
  static __device__ __forceinline__ unsigned int pack(float a, float b) {
    const __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<const unsigned int*>(&v);
  }
  static __device__ __forceinline__ unsigned int take(unsigned int v, unsigned int best) {
    const __half2 a = *reinterpret_cast<const __half2*>(&v), b = *reinterpret_cast<const __half2*>(&best);
    return ~__hle2_mask(a, b) & __heq2_mask(b, b);
  }
};
template <> struct Pack2<float> {
  static constexpr unsigned int NEG_INF = 0u;
  static __device__ __forceinline__ unsigned int pack(float, float) { return 0u; }
  static __device__ __forceinline__ unsigned int take(unsigned int, unsigned int) { return 0u; }
};

template <typename T, typename OutT, typename L>
__global__ void __launch_bounds__(256) k_decode_up4(const T* __restrict__ in, int B, int C, int h, int w, int n0, int n1,
                                                    int n2, OutT* __restrict__ o0, OutT* __restrict__ o1,
                                                    OutT* __restrict__ o2, const L* __restrict__ label,
                                                    unsigned long long* __restrict__ counts) {
  const int H = 4 * h, W = 4 * w;
  const long total = (long)B * h * w;
  const int e0 = n0, e1 = n0 + max(n1, 0), e2 = e1 + max(n2, 0);
  long long correct = 0, valid = 0;
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
    const int X = (int)(g % w);
    const long r = g / w;
    const int Y = (int)(r % h), b = (int)(r / h);
    // Source rows of output row 4Y + j: (Y-1, Y) for j < 2, (Y, Y+1) for j >= 2, clamped into the map; weights from the
    // forward formula (they differ from the constants (2j-3)/8 only in row / column 0, where the clamped source index
    // gives (1, 0): with rows (0, 0) in the neighbourhood that is the same value, but the weights are kept exact so that
    // a non-finite neighbour cannot leak in through a zero weight... it cannot: both taps are row 0 there).
    const int ym = max(Y - 1, 0), yp = min(Y + 1, h - 1), xm = max(X - 1, 0), xp = min(X + 1, w - 1);
    float wy1[4], wx1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { wy1[j] = lerp_src(4 * Y + j, 0.25f, h).l1; wx1[j] = lerp_src(4 * X + j, 0.25f, w).l1; }
    const T* base = in + (long)b * C * h * w;
    const int o_m = ym * w, o_c = Y * w, o_p = yp * w;
    float best[4][4];
    int arg[4][4];
    unsigned int best2[4][2], arg2[4][2];      // 16-bit logits: packed pairs (k, k + 1)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { best[j][k] = -INFINITY; arg[j][k] = 0; }
      best2[j][0] = best2[j][1] = Pack2<T>::NEG_INF;
      arg2[j][0] = arg2[j][1] = 0u;
    }
    int lvl = 0, cbeg = 0;
#pragma unroll 1
    for (int c = 0; c < e2; ++c) {
      const T* pc = base + (long)c * h * w;
      float v[3][3];
      v[0][0] = to_f32<T>(__ldg(pc + o_m + xm)); v[0][1] = to_f32<T>(__ldg(pc + o_m + X)); v[0][2] = to_f32<T>(__ldg(pc + o_m + xp));
      v[1][0] = to_f32<T>(__ldg(pc + o_c + xm)); v[1][1] = to_f32<T>(__ldg(pc + o_c + X)); v[1][2] = to_f32<T>(__ldg(pc + o_c + xp));
      v[2][0] = to_f32<T>(__ldg(pc + o_p + xm)); v[2][1] = to_f32<T>(__ldg(pc + o_p + X)); v[2][2] = to_f32<T>(__ldg(pc + o_p + xp));
      // horizontal lerps of the three source rows (shared by the output rows that use them), then the vertical ones
      float hl[3][4];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < 4; ++k) hl[a][k] = lerp2(v[a][k >> 1], v[a][(k >> 1) + 1], 1.0f - wx1[k], wx1[k]);
      if (sizeof(T) == 2) {
        const unsigned int cc = (unsigned int)(c - cbeg) * 0x00010001u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int kp = 0; kp < 2; ++kp) {
            const unsigned int v2 = Pack2<T>::pack(lerp2(hl[j >> 1][2 * kp], hl[(j >> 1) + 1][2 * kp], 1.0f - wy1[j], wy1[j]),
                                                   lerp2(hl[j >> 1][2 * kp + 1], hl[(j >> 1) + 1][2 * kp + 1], 1.0f - wy1[j], wy1[j]));
            const unsigned int take = Pack2<T>::take(v2, best2[j][kp]);
            best2[j][kp] = (best2[j][kp] & ~take) | (v2 & take);
            arg2[j][kp] = (arg2[j][kp] & ~take) | (cc & take);
          }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float val = round_as<T>(lerp2(hl[j >> 1][k], hl[(j >> 1) + 1][k], 1.0f - wy1[j], wy1[j]));
            argmax_step(val, c - cbeg, best[j][k], arg[j][k]);
          }
      }
      const int lend = lvl == 0 ? e0 : (lvl == 1 ? e1 : e2);
      if (c + 1 == lend) {
        OutT* out = lvl == 0 ? o0 : (lvl == 1 ? o1 : o2);
        if (sizeof(T) == 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {
              arg[j][2 * kp] = (int)(arg2[j][kp] & 0xffffu);
              arg[j][2 * kp + 1] = (int)(arg2[j][kp] >> 16);
              best2[j][kp] = Pack2<T>::NEG_INF;
              arg2[j][kp] = 0u;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const long off = ((long)b * H + 4 * Y + j) * W + 4 * X;
          if (out != nullptr) store_args4(out + off, arg[j]);
          if (lvl == 0 && label != nullptr) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const long long t = lab_ld(label, off + k);
              valid += t != SH_IGNORE;
              correct += (t != SH_IGNORE) & (t == arg[j][k]);
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) { best[j][k] = -INFINITY; arg[j][k] = 0; }
        }
        cbeg = lend;
        ++lvl;
        while (lvl < 3 && (lvl == 1 ? e1 : e2) == cbeg) ++lvl;
      }
    }
  }
  if (label != nullptr && counts != nullptr) {
    correct = warp_sum(correct);
    valid = warp_sum(valid);
    if ((threadIdx.x & 31) == 0 && valid) {
      atomicAdd(counts, (unsigned long long)correct);
      atomicAdd(counts + 1, (unsigned long long)valid);
    }
  }
}

}  // namespace sh

extern "C" {

static long sh_up_blocks(long items, int per) {
  long blocks = (items + per - 1) / per;
  if (blocks > SH_NUM_SMS * 16L) blocks = SH_NUM_SMS * 16L;
  return blocks < 1 ? 1 : blocks;
}

int sh_upsample_bilinear(const void* in, int dtype, void* out, long planes, int h, int w, int H, int W, void* stream) {
  if (planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return SH_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long items = planes * H * ((W + 3) / 4);
#define SH_UP(T)                                                                                          \
  {                                                                                                       \
    const int vec_ok = (W % 4 == 0) && ((uintptr_t)out % (4 * sizeof(T)) == 0);                           \
    if (H == 4 * h && W == 4 * w && vec_ok)                                                               \
    {                                                                                                     \
      const int txn = (w + 31) / 32, tyn = (h + 7) / 8;                                                   \
      const dim3 grid((unsigned)(txn * tyn), (unsigned)(planes < 65535 ? planes : 65535));                \
      sh::k_upsample4<T><<<grid, 256, 0, st>>>((const T*)in, (T*)out, planes, h, w, txn);                  \
    }                                                                                                     \
    else                                                                                                  \
      sh::k_upsample<T><<<(unsigned)sh_up_blocks(items, 256), 256, 0, st>>>((const T*)in, (T*)out, planes, h, w, H, W, vec_ok); \
  }                                                                                                       \
  break
  switch (dtype) {
    case SH_DT_F32: SH_UP(float);
    case SH_DT_BF16: SH_UP(__nv_bfloat16);
    case SH_DT_F16: SH_UP(__half);
    default: return SH_ERR_UNSUPPORTED;
  }
#undef SH_UP
  SH_CHECK_LAUNCH();
  return SH_OK;
}

int sh_upsample_bilinear_adjoint(const void* gout, int dtype, void* gin, long planes, int h, int w, int H, int W,
                                 void* stream) {
  if (planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return SH_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)sh_up_blocks(planes * h * w, 256);
#define SH_ADJ(T)                                                                                                   \
  if (H == 4 * h && W == 4 * w && (uintptr_t)gout % (4 * sizeof(T)) == 0) {                                         \
    const int txn = (w + sh::ADJ_TW - 1) / sh::ADJ_TW, tyn = (h + sh::ADJ_TH - 1) / sh::ADJ_TH;                      \
    long nb = planes * txn * tyn;                                                                                   \
    CUtensorMap mg;                                                                                                 \
    const char* adj_e = std::getenv("SEGHIERO_B200_ADJOINT");      /* "legacy": the cp.async form (A/B measurements) */ \
    const bool want_tma = adj_e == nullptr || std::strcmp(adj_e, "legacy") != 0;                                    \
    if (want_tma && planes < (1L << 30) &&                                                                          \
        sh::make_plane_map(&mg, sh::TmaType<T>::v, (int)sizeof(T), gout, W, H, planes, sh::AdjBox<T>::SC, sh::ADJ_SR)) { \
      const int smem = sh::ADJ_NST * sh::AdjBox<T>::STAGE;                                                          \
      if (nb > SH_NUM_SMS * 3L) nb = SH_NUM_SMS * 3L;                                                               \
      auto kern = sh::k_upsample4_adjoint_tma<T>;                                                                   \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                                \
      kern<<<(unsigned)nb, sh::ADJ_NT2, smem, st>>>(mg, (const T*)gout, (T*)gin, planes, h, w, txn, tyn);            \
    } else {                                                                                                        \
      if (nb > SH_NUM_SMS * 32L) nb = SH_NUM_SMS * 32L;                                                             \
      sh::k_upsample4_adjoint<T><<<(unsigned)nb, sh::ADJ_TH * sh::ADJ_TW, 0, st>>>((const T*)gout, (T*)gin, planes, \
                                                                                  h, w, txn, tyn);                  \
    }                                                                                                               \
  }                                                                                                                 \
  else                                                                                                              \
    sh::k_upsample_adjoint<T><<<blocks, 256, 0, st>>>((const T*)gout, (T*)gin, planes, h, w, H, W);                  \
  break
  switch (dtype) {
    case SH_DT_F32: SH_ADJ(float);
    case SH_DT_BF16: SH_ADJ(__nv_bfloat16);
    case SH_DT_F16: SH_ADJ(__half);
    default: return SH_ERR_UNSUPPORTED;
  }
#undef SH_ADJ
  SH_CHECK_LAUNCH();
  return SH_OK;
}

// footprint (low-resolution pixels per axis) of an AUX_T-pixel output span
static int sh_aux_footprint(int span, int n_in, int n_out) {
  const int f = (int)(((long)span * n_in + n_out - 1) / n_out) + 3;
  return f > n_in ? n_in : f;
}

size_t sh_aux_ce_workspace_bytes(int B, int C, int h, int w) { return (size_t)B * C * h * w * 8 + 64; }

int sh_aux_ce_fwdbwd(const void* logits, int dtype, const void* label, int label_dtype, int B, int C, int h, int w, int H,
                     int W, void* grad /* nullable, [B,C,h,w] */, const float* grad_out /* nullable */, float* out_loss,
                     void* workspace, void* stream) {
  if (B <= 0 || C <= 0 || C > 4096 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return SH_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  double* sums = (double*)workspace;                       // [0] loss sum, [1] #valid, [2] error flag
  unsigned long long* gacc = (unsigned long long*)((unsigned char*)workspace + 64);  // fixed-point accumulator of the low-resolution gradient
  const long n = (long)B * C * h * w;
  cudaError_t e = cudaMemsetAsync(workspace, 0, grad ? 64 + (size_t)n * 8 : 64, st);
  if (e != cudaSuccess) return (int)e;
  // integer scale 4 / 8 / 16 (W = k w, H = k h): the strip kernel, PX = k / 2 pixels per thread; anything else: a thread per pixel
  const int k = (w > 0 && W % w == 0 && h > 0 && H % h == 0 && W / w == H / h) ? W / w : 0;
  const int px = (k == 4 || k == 8 || k == 16) ? k / 2 : 0;
  const int tw = px ? 32 * px : sh::AUX_TW, th = sh::AUX_TH;
  const int fh = sh_aux_footprint(th, h, H), fw = sh_aux_footprint(tw, w, W);
  const size_t smem = (size_t)C * fh * fw * 8;
  if (smem > 200 * 1024) return SH_ERR_UNSUPPORTED;
  const int tiles = ((W + tw - 1) / tw) * ((H + th - 1) / th);
#define SH_AUX_LAUNCH(KERN)                                                                                        \
  {                                                                                                                \
    auto kern = KERN;                                                                                              \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                            \
    kern<<<B * tiles, 256, smem, st>>>((const T_*)logits, (const L*)label, grad ? gacc : nullptr, B, C, h, w, H, W, \
                                       fh, fw, sums);                                                              \
  }
#define SH_AUX(T)                                                                                                  \
  SH_LABEL_SWITCH(label_dtype, L, {                                                                                \
    using T_ = T;                                                                                                  \
    if (px == 8) SH_AUX_LAUNCH((sh::k_aux_ce_strip<T, L, 8>))                                                      \
    else if (px == 4) SH_AUX_LAUNCH((sh::k_aux_ce_strip<T, L, 4>))                                                 \
    else if (px == 2) SH_AUX_LAUNCH((sh::k_aux_ce_strip<T, L, 2>))                                                 \
    else SH_AUX_LAUNCH((sh::k_aux_ce<T, L>))                                                                       \
  })                                                                                                               \
  SH_CHECK_LAUNCH();                                                                                               \
  sh::k_aux_finish<T><<<(unsigned)sh_up_blocks(grad ? n : 1, 256), 256, 0, st>>>(gacc, (T*)grad, n, sums, grad_out, out_loss); \
  break
  switch (dtype) {
    case SH_DT_F32: SH_AUX(float);
    case SH_DT_BF16: SH_AUX(__nv_bfloat16);
    case SH_DT_F16: SH_AUX(__half);
    default: return SH_ERR_UNSUPPORTED;
  }
#undef SH_AUX
#undef SH_AUX_LAUNCH
  SH_CHECK_LAUNCH();
  return SH_OK;
}

int sh_decode_upsampled(const void* logits, int dtype, int B, int C, int h, int w, int H, int W, int n0, int n1, int n2,
                        void* out0, void* out1, void* out2, int out_is_u8, const void* label, int label_dtype,
                        unsigned long long* counts, void* stream) {
  if (B <= 0 || h <= 0 || w <= 0) return SH_OK;
  if (n0 <= 0 || n0 + (n1 > 0 ? n1 : 0) + (n2 > 0 ? n2 : 0) > C) return SH_ERR_BAD_ARG;
  if (H != 4 * h || W != 4 * w) return SH_ERR_UNSUPPORTED;    // other sizes: sh_upsample_bilinear + sh_decode
  if (((uintptr_t)out0 | (uintptr_t)out1 | (uintptr_t)out2) % (out_is_u8 ? 4 : 16)) return SH_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)sh_up_blocks((long)B * h * w, 256);
#define SH_DU(T)                                                                                                   \
  SH_LABEL_SWITCH(label_dtype, L, {                                                                                \
    if (out_is_u8)                                                                                                 \
      sh::k_decode_up4<T, unsigned char, L><<<blocks, 256, 0, st>>>((const T*)logits, B, C, h, w, n0, n1, n2,       \
          (unsigned char*)out0, (unsigned char*)out1, (unsigned char*)out2, (const L*)label, counts);              \
    else                                                                                                           \
      sh::k_decode_up4<T, long long, L><<<blocks, 256, 0, st>>>((const T*)logits, B, C, h, w, n0, n1, n2,          \
          (long long*)out0, (long long*)out1, (long long*)out2, (const L*)label, counts);                          \
  })                                                                                                               \
  break
  switch (dtype) {
    case SH_DT_F32: SH_DU(float);
    case SH_DT_BF16: SH_DU(__nv_bfloat16);
    case SH_DT_F16: SH_DU(__half);
    default: return SH_ERR_UNSUPPORTED;
  }
#undef SH_DU
  SH_CHECK_LAUNCH();
  return SH_OK;
}

}  // extern "C"
