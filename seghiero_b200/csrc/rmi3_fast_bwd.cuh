// Fast path of the three-level loss, backward side (pass 2): gradient of tree BCE + CE + RMI w.r.t. the logits.
//
// Applies to tree-shaped maps, W % 4 == 0, 16-byte aligned tensors, C <= 254 (the forward side may be either the
// fast or the generic pass 1: both leave the same summaries).
// Reference arithmetic: models/loss/rmi_hiera_triplet_loss.py:349-526 (autograd of it); analytic RMI backward in
// oracle/rmi_taps.py.  The generic kernel (rmi3_bwd.cu::k3_pass2) computes the same thing for every other case.
//
// One CTA = one 64 x 32 tile of one image, 128 threads, thread = 4 x 4 pixel block (+ one piece of the tile's
// 2-pixel ring).  Per channel, in tree order (fine children, their mid, ..., the high):
//   phase A  sigmoid / e^x (3 MUFU), tree-BCE + CE gradient from the per-pixel summaries of pass 1 (holder
//            bytes, 1/sum e^x) -> 16 registers; P = s*valid + 1e-6 -> shared-memory plane (own block + ring piece)
//   phase B  5x5 stencil over the plane (weights from k3f_finalize), + the one-hot stencil where the block's
//            labels are mixed, combine, 128-bit store
// Planes are double buffered: phase B of channel c and phase A of channel c+1 sit between the same pair of
// barriers.  Everything a thread reads from global memory for channel c+1 (logits, 1/sum e^x of the channel's
// level, holder bytes at group starts) travels with cp.async into the thread's own shared-memory slots while
// phase B of channel c runs; no thread reads another thread's slots, so the only barrier is the plane hand-over.
// Shared memory per CTA is ~57 KB and registers <= 168, so three CTAs share an SM.
#pragma once
#include <type_traits>
#include "rmi3_common.cuh"

namespace sh {
namespace fast2 {

constexpr int TW = 64, TH = 32;
constexpr int NT = (TH / 4) * 16;         // 128 threads, each a 4 x 4 block
constexpr int PW = TW + 4, PR = TH + 4, PLANE = PR * PW;
constexpr int LP = TW + 8;                // label tile pitch (bytes): cols x0-2 .. x0+65 (+4 pad)
constexpr int WS = 64;                    // floats per staged weight record: W1 at 0, W2 at 28, W2full at 56
constexpr int SLOT = NT * 16;             // bytes of one staging row (16 bytes per thread)

struct Hier2 {
  int nf, nm, nh;
  const int* f2m;             // [nf]
  const int* f2h;             // [nf]
  const unsigned int* order;  // [C] kind | class << 8 | flags << 16 | channel << 24 ; flags bit0/bit2 = first of a mid/high group
  const unsigned int* aux;    // [C] mid id | high id << 8 (0xff = none)
};

inline size_t pass2_smem(int C, int nf, int nm, int nh) {
  size_t s = (size_t)2 * PLANE * 4;                 // planes
  s += (size_t)4 * SLOT;                            // logits of the next channel (4 rows per thread)
  s += (size_t)4 * SLOT;                            // 1 / sum e^x of the next channel's level
  s += (size_t)2 * SLOT;                            // ring pieces: strip / pair + corner pair
  s += (size_t)4 * SLOT;                            // holder bytes: mid group, high group, fine target, mid target (4 rows x 4 bytes)
  s += (size_t)3 * PR * LP;                         // label tile
  s += (size_t)3 * WS * 4;                          // stencil weights (2 buffers) + their cp.async staging
  s += (size_t)C * 16 + (size_t)2 * nf * 4 + 64;    // tables
  s += (size_t)256 * 2 + 16;                        // deferred one-hot stencils: work list + counter
  return (s + 15) & ~(size_t)15;
}

__device__ __forceinline__ bool byte_is_zero(unsigned int z, int k) { return ((z >> (8 * k)) & 0xffu) == 0u; }

template <typename T, bool INLINE_OH>
__global__ void __launch_bounds__(NT, 3)
k3f_pass2(const T* __restrict__ x, T* __restrict__ grad, int B, int H, int W, Hier2 hg, Ws3 ws, float eps,
          float loss_weight, const float* __restrict__ gscale_ptr, int tiles_x, int tiles_per_img) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = hg.nf + hg.nm + hg.nh;
  float* planes = reinterpret_cast<float*>(smem_raw);                                  // [2][PLANE]
  unsigned char* xst = reinterpret_cast<unsigned char*>(planes + 2 * PLANE);           // [4][NT] x 16 B
  unsigned char* ist = xst + 4 * SLOT;                                                 // [4][NT] x 16 B
  unsigned char* rst = ist + 4 * SLOT;                                                 // [2][NT] x 16 B
  unsigned char* hst = rst + 2 * SLOT;                                                 // [4 kinds][NT] x (4 rows x 4 B)
  unsigned char* LT = hst + 4 * SLOT;                                                  // [3][PR][LP]
  float* wsm = reinterpret_cast<float*>(LT + 3 * PR * LP);                             // [2][WS]
  float* wst = wsm + 2 * WS;                                                           // [WS] raw weights of the next channel
  long long* s_chb = reinterpret_cast<long long*>(wst + WS);                           // [C]
  unsigned int* s_order = reinterpret_cast<unsigned int*>(s_chb + C);                  // [C]
  unsigned int* s_aux = s_order + C;                                                   // [C]
  int* s_f2m = reinterpret_cast<int*>(s_aux + C);                                      // [nf]
  int* s_f2h = s_f2m + hg.nf;                                                          // [nf]
  int* s_nwork = s_f2h + hg.nf;                                                        // [4] (one used)
  unsigned short* s_work = reinterpret_cast<unsigned short*>(s_nwork + 4);             // [256] block | order index << 7

  const int tid = threadIdx.x;
  {
    // Two instantiations are launched back to back; each image is served by one of them.  Images whose labels are
    // noisy (most 3x8 strips see more than one class, counted by k3f_prep) do their one-hot stencils inline, the
    // others defer them to a work list (see phase B).  Without statistics (generic pass 1) everything defers.
    const int bb = blockIdx.x / tiles_per_img;
    const bool noisy = 2u * ws.strips[2 * bb] > ws.strips[2 * bb + 1];
    if (noisy != INLINE_OH) return;
  }
  if (tid == 0) s_nwork[0] = 0;
  const long HW = (long)H * W, BHW = (long)B * HW;
  const int b = blockIdx.x / tiles_per_img, tile = blockIdx.x - b * tiles_per_img;
  const int tyi = tile / tiles_x, ty0 = tyi * TH, tx0 = (tile - tyi * tiles_x) * TW;
  const unsigned char* lab8 = ws.lab8 + (long)b * HW;
  const char* xbb = reinterpret_cast<const char*>(x + (long)b * C * HW);
  const bool border = ty0 < 2 || ty0 + TH > H - 2 || tx0 < 2 || tx0 + TW > W - 2;

  for (int i = tid; i < C; i += NT) {
    const unsigned int oe = hg.order[i];
    s_order[i] = oe;
    s_aux[i] = hg.aux[i];
    s_chb[i] = (long long)(oe >> 24) * HW * (long long)sizeof(T);
  }
  for (int i = tid; i < hg.nf; i += NT) { s_f2m[i] = hg.f2m[i]; s_f2h[i] = hg.f2h[i]; }
  __syncthreads();                                     // tables are in place

  // ---- label tile (RMI labels of the 3 levels; outside the image 0xff) ----
  for (int e = tid; e < PR * (PW / 2); e += NT) {
    const int r = e / (PW / 2), j = (e - r * (PW / 2)) * 2;
    const int yy = ty0 - 2 + r, xx = tx0 - 2 + j;
    unsigned int f2 = 0xffffu, m2 = 0xffffu, g2 = 0xffffu;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const unsigned int t2 = *reinterpret_cast<const unsigned short*>(lab8 + (long)yy * W + xx);
      f2 = m2 = g2 = 0u;    // void pixels are one-hot of class 0 at every level inside RMI (rmi...py:360-370)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const unsigned int t = (t2 >> (8 * k)) & 0xffu;
        if (t != SH_IGNORE) {
          f2 |= t << (8 * k);
          m2 |= (unsigned int)s_f2m[t] << (8 * k);
          g2 |= (unsigned int)s_f2h[t] << (8 * k);
        }
      }
    }
    *reinterpret_cast<unsigned short*>(LT + (0 * PR + r) * LP + j) = (unsigned short)f2;
    *reinterpret_cast<unsigned short*>(LT + (1 * PR + r) * LP + j) = (unsigned short)m2;
    *reinterpret_cast<unsigned short*>(LT + (2 * PR + r) * LP + j) = (unsigned short)g2;
  }

  const float gscale = *gscale_ptr;
  const int rq = tid >> 4, st = tid & 15;
  const int xg = tx0 + 4 * st;
  const bool colok = xg < W;
  const float nv = fmaxf((float)ws.counts[0], 1.0f);
  const float wF = 2.5f * loss_weight * gscale / (nv * (float)hg.nf);
  const float wM = 2.5f * loss_weight * gscale / (nv * (float)hg.nm);
  const float wH = 2.5f * loss_weight * gscale / (nv * (float)hg.nh);
  const float wCE = loss_weight * gscale / ((float)B * (float)HW);

  bool rowok[4];
  long roff[4];                       // pixel offset of the strip inside one channel plane (clamped into the image)
  unsigned int tc0[4], tc1[4], tc2[4], hmN[4], hhN[4];
  // classes (mod 64, per level) that are the target of some pixel of the block: a set bit sends the channel through
  // the positive-term path; a collision only costs that detour
  unsigned long long presF = 0ull, presM = 0ull, presH = 0ull;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = ty0 + 4 * rq + j;
    rowok[j] = y < H && colok;
    roff[j] = (long)min(y, H - 1) * W + (colok ? xg : 0);
    const unsigned int t4 = rowok[j] ? *reinterpret_cast<const unsigned int*>(lab8 + roff[j]) : 0xffffffffu;
    tc0[j] = t4; tc1[j] = 0xffffffffu; tc2[j] = 0xffffffffu;
    hmN[j] = hhN[j] = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned int t = (t4 >> (8 * k)) & 0xffu;
      if (t != SH_IGNORE) {
        const unsigned int cm = (unsigned int)(hg.nf + s_f2m[t]), chh = (unsigned int)(hg.nf + hg.nm + s_f2h[t]);
        tc1[j] = (tc1[j] & ~(0xffu << (8 * k))) | (cm << (8 * k));
        tc2[j] = (tc2[j] & ~(0xffu << (8 * k))) | (chh << (8 * k));
        presF |= 1ull << (t & 63u);
        presM |= 1ull << (s_f2m[t] & 63);
        presH |= 1ull << (s_f2h[t] & 63);
      }
    }
  }
  // interior masks (border tiles only): bit 4*j + k
  unsigned int imask = 0xffffu;
  if (border) {
    imask = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int y = ty0 + 4 * rq + j;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (y >= 2 && y < H - 2 && xg + k >= 2 && xg + k < W - 2) imask |= 1u << (4 * j + k);
    }
  }
  // ---- this thread's piece of the ring: tid < 64: 4-pixel strip of plane rows 0,1,TH+2,TH+3 ; tid >= 64: 2-pixel pair
  //      left / right of a body row ; tid < 8 also a 2-pixel corner pair ----
  int pidx0, pidx1;
  long goff0, goff1;
  bool in0, in1;
  float hv[6];
  const bool is_strip = tid < 64;
  if (is_strip) {
    const int hrow = tid >> 4, pr = hrow < 2 ? hrow : TH + hrow, strip = tid & 15;
    const int yy = ty0 - 2 + pr, xx = tx0 + 4 * strip;
    pidx0 = pr * PW + 2 + 4 * strip;
    in0 = yy >= 0 && yy < H && xx < W;
    goff0 = in0 ? (long)yy * W + xx : 0;
  } else {
    const int hl = tid - 64, srow = hl >> 1, side = hl & 1;
    const int yy = ty0 + srow, xx = side ? tx0 + TW : tx0 - 2;
    pidx0 = (2 + srow) * PW + (side ? TW + 2 : 0);
    in0 = yy < H && xx >= 0 && xx < W;
    goff0 = in0 ? (long)yy * W + xx : 0;
  }
  {
    const int crow = (tid >> 1) & 3, pr = crow < 2 ? crow : TH + crow, side = tid & 1;
    const int yy = ty0 - 2 + pr, xx = side ? tx0 + TW : tx0 - 2;
    pidx1 = pr * PW + (side ? TW + 2 : 0);
    in1 = tid < 8 && yy >= 0 && yy < H && xx >= 0 && xx < W;
    goff1 = in1 ? (long)yy * W + xx : 0;
  }
  {
    unsigned int t4 = 0xffffffffu;
    if (in0) t4 = is_strip ? *reinterpret_cast<const unsigned int*>(lab8 + goff0)
                           : (0xffff0000u | *reinterpret_cast<const unsigned short*>(lab8 + goff0));
#pragma unroll
    for (int k = 0; k < 4; ++k) hv[k] = ((t4 >> (8 * k)) & 0xffu) != SH_IGNORE ? 1.f : 0.f;
    const unsigned int t2 = in1 ? *reinterpret_cast<const unsigned short*>(lab8 + goff1) : 0xffffu;
    hv[4] = (t2 & 0xffu) != SH_IGNORE ? 1.f : 0.f;
    hv[5] = (t2 >> 8) != SH_IGNORE ? 1.f : 0.f;
  }

  // ---- staging slots of this thread ----
  const unsigned int xs_base = (unsigned int)__cvta_generic_to_shared(xst + tid * 16);
  const unsigned int is_base = (unsigned int)__cvta_generic_to_shared(ist + tid * 16);
  const unsigned int rs_base = (unsigned int)__cvta_generic_to_shared(rst + tid * 16);
  const unsigned int hs_base = (unsigned int)__cvta_generic_to_shared(hst + tid * 16);
  const unsigned char* xs_gen = xst + tid * 16;
  const unsigned char* is_gen = ist + tid * 16;
  const unsigned char* rs_gen = rst + tid * 16;
  const unsigned char* hs_gen = hst + tid * 16;
  const unsigned char* holdb = ws.hold + (long)b * HW;
  const float* invb = ws.inv + (long)b * HW;

  // everything phase A of channel ci needs from global memory -> this thread's slots (one commit group)
  auto prefetch = [&](int ci) {
    const unsigned int oe = s_order[ci], ax = s_aux[ci];
    const int kind = oe & 3;
    const unsigned int fl = (oe >> 16) & 0xffu;
    const char* g = xbb + s_chb[ci];
    const float* ivl = invb + (long)kind * BHW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const char* gp = g + roff[j] * (long)sizeof(T);
      if (sizeof(T) == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xs_base + j * SLOT), "l"(gp));
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(xs_base + j * SLOT), "l"(gp));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(is_base + j * SLOT), "l"(ivl + roff[j]));
    }
    if (is_strip) {
      if (sizeof(T) == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rs_base), "l"(g + goff0 * 4));
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(rs_base), "l"(g + goff0 * 2));
    } else {
      if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(rs_base), "l"(g + goff0 * 4));
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(rs_base), "l"(g + goff0 * 2));
    }
    if (tid < 8) {
      if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(rs_base + SLOT), "l"(g + goff1 * 4));
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(rs_base + SLOT), "l"(g + goff1 * 2));
    }
    if (fl & 1u) {       // first channel of a mid group: the bytes that say which channel holds the group's max
      const unsigned char* hp = holdb + (long)(ax & 0xffu) * BHW;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(hs_base + 4 * j), "l"(hp + roff[j]));
    }
    if ((fl & 4u) && (ax >> 8) != 0xffu) {
      const unsigned char* hp = holdb + (long)(hg.nm + (ax >> 8)) * BHW;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(hs_base + SLOT + 4 * j), "l"(hp + roff[j]));
    }
    if (tid < 25 || tid == 28) {   // stencil weights of the channel (k3f_finalize: W1[25], W2[25], sum W2 at 50)
      const float* src = ws.wts + ((size_t)b * C + (oe >> 24)) * 64;
      const unsigned int wb = (unsigned int)__cvta_generic_to_shared(wst);
      if (tid < 25) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(wb + 4 * tid), "l"(src + tid));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(wb + 4 * (25 + tid)), "l"(src + 25 + tid));
      } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(wb + 4 * 50), "l"(src + 50));
      }
    }
    cp_async_commit();
  };

  // holders of the positive terms (fine target: min(A_t, B_m); mid target: min(C_h, B_m)) for the whole tile walk
  {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const unsigned char* hp = holdb + (long)(hg.nm + hg.nh + q) * BHW;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(hs_base + (2 + q) * SLOT + 4 * j), "l"(hp + roff[j]));
    }
  }
  prefetch(0);

  __syncthreads();                                     // label tile is complete
  // label structure of the block's 8 x 8 neighbourhood per level: uniform class (or 0xfe = mixed) and the classes present
  unsigned int ublk[3], pres[3];
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const unsigned char* lt = LT + (l * PR + 4 * rq) * LP + 4 * st;
    const unsigned int first = *reinterpret_cast<const unsigned int*>(lt) & 0xffu;
    const unsigned int pat = first * 0x01010101u;
    unsigned int diff = 0u, hash = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const unsigned int wd = *reinterpret_cast<const unsigned int*>(lt + i * LP + 4 * q);
        diff |= wd ^ pat;
        hash |= (1u << (wd & 31)) | (1u << ((wd >> 8) & 31)) | (1u << ((wd >> 16) & 31)) | (1u << ((wd >> 24) & 31));
      }
    }
    ublk[l] = diff == 0u ? first : 0xfeu;
    pres[l] = hash;
  }

  float g0[4][4];

  // phase A of channel (order index) ci: plane (ci & 1), gradient of BCE + CE -> g0
  auto phaseA = [&](int ci) {
    const unsigned int oe = s_order[ci], ax = s_aux[ci];
    const int kind = oe & 3;
    const unsigned int fl = (oe >> 16) & 0xffu, ch = oe >> 24;
    const unsigned int cc = ch * 0x01010101u;
    float* plane = planes + (ci & 1) * PLANE;
    if (fl & 1u) {       // where does the group's (1 - max) term count: everywhere but at pixels whose target is this mid
      const unsigned int midc = (unsigned int)(hg.nf + (ax & 0xffu)) * 0x01010101u;
      const uint4 hm = *reinterpret_cast<const uint4*>(hs_gen);
      hmN[0] = hm.x | __vcmpeq4(tc1[0], midc); hmN[1] = hm.y | __vcmpeq4(tc1[1], midc);
      hmN[2] = hm.z | __vcmpeq4(tc1[2], midc); hmN[3] = hm.w | __vcmpeq4(tc1[3], midc);
    }
    if (fl & 4u) {
      const unsigned int high = ax >> 8;
      if (high != 0xffu) {
        const unsigned int highc = (unsigned int)(hg.nf + hg.nm + high) * 0x01010101u;
        const uint4 hh = *reinterpret_cast<const uint4*>(hs_gen + SLOT);
        hhN[0] = hh.x | __vcmpeq4(tc2[0], highc); hhN[1] = hh.y | __vcmpeq4(tc2[1], highc);
        hhN[2] = hh.z | __vcmpeq4(tc2[2], highc); hhN[3] = hh.w | __vcmpeq4(tc2[3], highc);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) hhN[j] = 0xffffffffu;
      }
    }
    // ---- ring piece(s): sigmoid only ----
    {
      float xv[4];
      if (is_strip) staged_vec4<T>(rs_gen, xv);
      else {
        if (sizeof(T) == 4) { const float2 t2 = *reinterpret_cast<const float2*>(rs_gen); xv[0] = t2.x; xv[1] = t2.y; }
        else { xv[0] = staged_elem<T>(rs_gen, 0); xv[1] = staged_elem<T>(rs_gen, 1); }
        xv[2] = xv[3] = 0.f;
      }
      if (in0) {
        *reinterpret_cast<float2*>(plane + pidx0) = make_float2(fmaf(sig_only(xv[0]), hv[0], 1e-6f), fmaf(sig_only(xv[1]), hv[1], 1e-6f));
        if (is_strip)
          *reinterpret_cast<float2*>(plane + pidx0 + 2) = make_float2(fmaf(sig_only(xv[2]), hv[2], 1e-6f), fmaf(sig_only(xv[3]), hv[3], 1e-6f));
      }
      if (in1) {
        float a0, a1;
        if (sizeof(T) == 4) { const float2 t2 = *reinterpret_cast<const float2*>(rs_gen + SLOT); a0 = t2.x; a1 = t2.y; }
        else { a0 = staged_elem<T>(rs_gen + SLOT, 0); a1 = staged_elem<T>(rs_gen + SLOT, 1); }
        *reinterpret_cast<float2*>(plane + pidx1) = make_float2(fmaf(sig_only(a0), hv[4], 1e-6f), fmaf(sig_only(a1), hv[5], 1e-6f));
      }
    }
    // ---- stencil weights of the channel times the upstream gradient (every thread moves the words it staged itself) ----
    {
      float* dst = wsm + (ci & 1) * WS;
      if (tid < 25) { dst[tid] = wst[tid] * gscale; dst[28 + tid] = wst[25 + tid] * gscale; }
      else if (tid < 28) { dst[tid] = 0.f; dst[28 + tid] = 0.f; }
      else if (tid == 28) dst[56] = wst[50] * gscale;
    }
    const float wbase = kind == 0 ? wF : 0.f;
    const unsigned int clA = (oe >> 8) & 0xffu;
    const bool pos = ((kind == 0 ? presF : (kind == 1 ? presM : presH)) >> (clA & 63u)) & 1ull;
    float* prow = plane + (4 * rq + 2) * PW + 4 * st + 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float xv[4];
      staged_vec4<T>(xs_gen + j * SLOT, xv);
      const float4 iv4 = *reinterpret_cast<const float4*>(is_gen + j * SLOT);
      const float ivk[4] = {iv4.x, iv4.y, iv4.z, iv4.w};       // 1 / sum e^x of the level; 0 on void pixels
      const unsigned int zM = hmN[j] ^ cc, zH = hhN[j] ^ cc;
      float s[4], E[4], t[4], oh[4], ds[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        sig_exp3(xv[k], s[k], E[k]);
        t[k] = 1.0f - s[k];
        oh[k] = 0.f;
      }
      // d/ds of the -log(1 - . + eps) terms this channel holds: own fine term, its mid group's max, its high group's max
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float r = rcp(t[k] + eps);
        ds[k] = wbase * r;
        if (byte_is_zero(zM, k)) ds[k] = fmaf(wM, r, ds[k]);
        if (byte_is_zero(zH, k)) ds[k] = fmaf(wH, r, ds[k]);
      }
      if (pos) {
        const unsigned int zT = (kind == 0 ? tc0[j] : (kind == 1 ? tc1[j] : tc2[j])) ^ cc;
        const unsigned int zPF = *reinterpret_cast<const unsigned int*>(hs_gen + 2 * SLOT + 4 * j) ^ cc;
        const unsigned int zPM = *reinterpret_cast<const unsigned int*>(hs_gen + 3 * SLOT + 4 * j) ^ cc;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float Bp = byte_is_zero(zPF, k) ? wF : 0.f;
          if (byte_is_zero(zPM, k)) Bp += wM;
          if (byte_is_zero(zT, k)) {
            oh[k] = 1.f;
            ds[k] = fmaf(-wbase, rcp(t[k] + eps), ds[k]);      // the target has no own (1 - s) term
            if (kind == 2) Bp += wH;
          }
          ds[k] = fmaf(-Bp, rcp(s[k] + eps), ds[k]);
        }
      }
      float P[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float s2 = ivk[k] != 0.f ? s[k] : 0.f;        // s * valid
        P[k] = s2 + 1e-6f;                                    // = probs * valid + 1e-6 (rmi...py:487)
        const float q = s2 * t[k];
        const float ce = fmaf(E[k], ivk[k], -oh[k]);          // softmax - one-hot
        g0[j][k] = fmaf(ds[k], q, wCE * ce);
      }
      *reinterpret_cast<float2*>(prow + j * PW) = make_float2(P[0], P[1]);
      *reinterpret_cast<float2*>(prow + j * PW + 2) = make_float2(P[2], P[3]);
    }
  };

  // phase B of channel ci: stencil over plane (ci & 1), combine with g0, store
  auto phaseB = [&](int ci) {
    const unsigned int oe = s_order[ci];
    const int kind = oe & 3;
    const unsigned int cl = (oe >> 8) & 0xffu, ch = oe >> 24;
    const float* wp = wsm + (ci & 1) * WS;
    const float* pl = planes + (ci & 1) * PLANE + (4 * rq) * PW + 4 * st;
    const unsigned int ub = kind == 0 ? ublk[0] : (kind == 1 ? ublk[1] : ublk[2]);
    const unsigned int ph = kind == 0 ? pres[0] : (kind == 1 ? pres[1] : pres[2]);
    const float init = ub == cl ? wp[56] : 0.f;
    // mixed labels around the block and this class among them: the block needs the one-hot stencil too.  In images
    // where few strips are mixed it is deferred to a work list that the whole CTA drains 128 blocks at a time (inline
    // it would idle the other lanes of the warp); in noisy images (INLINE_OH) a second sweep right here is cheaper.
    const bool mixed_hit = ub == 0xfeu && ((ph >> (cl & 31)) & 1u);
    if (!INLINE_OH && mixed_hit) s_work[atomicAdd(s_nwork, 1)] = (unsigned short)(tid | (ci << 7));
    const int nsweep = (INLINE_OH && mixed_hit) ? 2 : 1;
    const unsigned char* lt = LT + (kind * PR + 4 * rq) * LP + 4 * st;
    float acc[4][4], q[4][4];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[o][k] = init;
#pragma unroll 1
    for (int sw = 0; sw < nsweep; ++sw) {
      float w[28];
#pragma unroll
      for (int v = 0; v < 7; ++v) {
        const float4 t4 = *reinterpret_cast<const float4*>(wp + sw * 28 + 4 * v);
        w[4 * v] = t4.x; w[4 * v + 1] = t4.y; w[4 * v + 2] = t4.z; w[4 * v + 3] = t4.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float win[8];
        if (sw == 0) {
          const float4 a = *reinterpret_cast<const float4*>(pl + i * PW);
          const float4 c4 = *reinterpret_cast<const float4*>(pl + i * PW + 4);
          win[0] = a.x; win[1] = a.y; win[2] = a.z; win[3] = a.w;
          win[4] = c4.x; win[5] = c4.y; win[6] = c4.z; win[7] = c4.w;
          if (i >= 2 && i < 6) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float s1 = win[k + 2] - 1e-6f;          // valid pixels: P = s + 1e-6 ; void: exactly 0
              q[i - 2][k] = s1 * (1.0f - s1);
            }
          }
        } else {
          const unsigned int l0 = *reinterpret_cast<const unsigned int*>(lt + i * LP);
          const unsigned int l1 = *reinterpret_cast<const unsigned int*>(lt + i * LP + 4);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            win[v] = ((l0 >> (8 * v)) & 0xffu) == cl ? 1.f : 0.f;
            win[4 + v] = ((l1 >> (8 * v)) & 0xffu) == cl ? 1.f : 0.f;
          }
        }
        // up to 16 independent accumulators between two updates of the same one
#pragma unroll
        for (int dx = 0; dx < 5; ++dx)
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int dyi = i - o;
            if (dyi < 0 || dyi > 4) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[o][k] = fmaf(w[dyi * 5 + dx], win[k + dx], acc[o][k]);
          }
      }
    }
    T* gp = grad + ((long)b * C + ch) * HW;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      if (border) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {     // frame pixels: no RMI term here (k3_frame2); their stencil sums may hold garbage
          const bool inter = (imask >> (4 * o + k)) & 1u;
          q[o][k] = inter ? q[o][k] : 0.f;
          acc[o][k] = inter ? acc[o][k] : 0.f;
        }
      }
      float g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) g[k] = fmaf(q[o][k], acc[o][k], g0[o][k]);
      if (rowok[o]) VecIO<T, 4>::store(gp + roff[o], g);
    }
  };

  // One deferred one-hot stencil: block `blk` (thread id of its owner), order index wci.  Adds
  //   gscale * valid * s(1-s) * sum_d W2[d] [L(r+d) == cl]   to the gradient the owner has already stored.
  auto onehot_item = [&](unsigned int item) {
    const int blk = item & 127, wci = item >> 7;
    const int brq = blk >> 4, bst = blk & 15;
    const unsigned int oe = s_order[wci];
    const int kind = oe & 3;
    const unsigned int cl = (oe >> 8) & 0xffu, ch = oe >> 24;
    const int bx = tx0 + 4 * bst;
    if (bx >= W) return;
    const float* wsrc = ws.wts + ((size_t)b * C + ch) * 64 + 25;
    const T* xc = x + ((long)b * C + ch) * HW;
    T* gc = grad + ((long)b * C + ch) * HW;
    // everything from global memory first: the latency hides behind the stencil sweep
    float w[25];
#pragma unroll
    for (int v = 0; v < 25; ++v) w[v] = __ldg(wsrc + v);
    uint4 xraw[4], graw[4];
    unsigned int t4[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int y = min(ty0 + 4 * brq + o, H - 1);
      const long off = (long)y * W + bx;
      t4[o] = *reinterpret_cast<const unsigned int*>(lab8 + off);
      if (sizeof(T) == 4) {
        xraw[o] = __ldg(reinterpret_cast<const uint4*>(xc + off));
        graw[o] = __ldcg(reinterpret_cast<const uint4*>(gc + off));     // written earlier in this kernel: coherent load
      } else {
        const uint2 xr = __ldg(reinterpret_cast<const uint2*>(xc + off));
        const uint2 gr = __ldcg(reinterpret_cast<const uint2*>(gc + off));
        xraw[o] = make_uint4(xr.x, xr.y, 0u, 0u);
        graw[o] = make_uint4(gr.x, gr.y, 0u, 0u);
      }
    }
    float acc[4][4];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[o][k] = 0.f;
    const unsigned char* lt = LT + (kind * PR + 4 * brq) * LP + 4 * bst;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned int l0 = *reinterpret_cast<const unsigned int*>(lt + i * LP);
      const unsigned int l1 = *reinterpret_cast<const unsigned int*>(lt + i * LP + 4);
      float win[8];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        win[v] = ((l0 >> (8 * v)) & 0xffu) == cl ? 1.f : 0.f;
        win[4 + v] = ((l1 >> (8 * v)) & 0xffu) == cl ? 1.f : 0.f;
      }
#pragma unroll
      for (int dx = 0; dx < 5; ++dx)
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int dyi = i - o;
          if (dyi < 0 || dyi > 4) continue;
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[o][k] = fmaf(w[dyi * 5 + dx], win[k + dx], acc[o][k]);
        }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int y = ty0 + 4 * brq + o;
      if (y >= H) break;
      const long off = (long)y * W + bx;
      float xv[4], g[4];
      staged_vec4<T>(&xraw[o], xv);
      staged_vec4<T>(&graw[o], g);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool inter = !border || (y >= 2 && y < H - 2 && bx + k >= 2 && bx + k < W - 2);
        const bool valid = ((t4[o] >> (8 * k)) & 0xffu) != SH_IGNORE;
        const float sg = sig_only(xv[k]);
        const float qk = (valid && inter) ? sg * (1.0f - sg) * gscale : 0.f;
        g[k] = fmaf(qk, acc[o][k], g[k]);
      }
      VecIO<T, 4>::store(gc + off, g);
    }
  };
  // drain 128 items (all threads busy) whenever that many are queued; `flush` takes whatever is left as well
  auto drain = [&](bool flush) {
#pragma unroll 1
    for (;;) {
      const int n = s_nwork[0];
      if (n < NT && !(flush && n > 0)) break;
      const int take = n < NT ? n : NT;
      unsigned int item = 0xffffffffu;
      if (tid < take) item = s_work[n - take + tid];
      __syncthreads();                                 // everyone has its item: the counter may move
      if (tid == 0) s_nwork[0] = n - take;
      if (item != 0xffffffffu) onehot_item(item);
      __syncthreads();                                 // counter visible before the next pushes / the next look
    }
  };

  // one copy of each phase in the instruction stream: iteration -1 only runs phase A of channel 0 (its logits were
  // requested above, before the label scan)
#pragma unroll 1
  for (int ci = -1; ci < C; ++ci) {
    if (ci >= 0 && ci + 1 < C) prefetch(ci + 1);
    if (ci >= 0) phaseB(ci);
    if (ci + 1 < C) {
      cp_async_wait<0>();
      phaseA(ci + 1);
    }
    __syncthreads();                                   // planes handed over; gradient stores of phase B visible to the CTA
    if (!INLINE_OH && ci >= 0) drain(ci + 1 == C);
  }
}

}  // namespace fast2
}  // namespace sh
