// Pass 2 of the three-level loss, TMA form (sm_100a): the same arithmetic and thread mapping as rmi3_fast_bwd.cuh
// (read that file first), with every per-channel global read moved out of the threads:
//   * the logit tile of the next channel INCLUDING its 2-pixel ring arrives as one cp.async.bulk.tensor box
//     (68 x 36, out-of-image elements zero-filled by the TMA unit), 1/sum e^x of the channel's level as a second box and
//     the holder bytes at group starts as a third / fourth; one elected thread issues them, an mbarrier with
//     complete_tx hands them over.  The legacy kernel spends ~215 of its ~1600 instructions per (thread, channel) on
//     cp.async address arithmetic and carries the addresses in ~20 registers.
//   * with those registers gone the kernel fits 128 registers -> 4 CTAs per SM (MINB = 4); measured, 3 CTAs with the
//     compiler's own 168 registers (MINB = 3, the default) is as fast: the kernel is bound by operand reads of the
//     register file, not by latency.
// Reference arithmetic: models/loss/rmi_hiera_triplet_loss.py:349-526 (autograd of it); analytic RMI backward in
// oracle/rmi_taps.py.
#pragma once
#include <type_traits>
#include "rmi3_fast_bwd.cuh"
#include "tma.cuh"

namespace sh {
namespace fast3 {

using fast2::Hier2;
using fast2::byte_is_zero;
constexpr int TW = fast2::TW, TH = fast2::TH, NT = fast2::NT, PW = fast2::PW, PR = fast2::PR, PLANE = fast2::PLANE;
constexpr int LP = fast2::LP, WS = fast2::WS, SLOT = fast2::SLOT;
constexpr int PLANE_B = ((PLANE * 4 + 127) / 128) * 128;          // bytes of one P plane, padded to the TMA alignment

// the logit box: columns x0-2-XO .. x0+65+XO; the XO extra columns on each side make the box START a multiple of 16 bytes
// (a TMA requirement for the innermost coordinate: an unaligned start is an illegal-instruction fault, measured with
// scripts/ubench/tma_probe.cu) and keep its inner extent one
template <typename T> struct XBox {
  static constexpr int XO = sizeof(T) == 4 ? 2 : 6;
  static constexpr int COLS = PW + 2 * XO;
  static constexpr int BYTES = PR * COLS * (int)sizeof(T);
  static constexpr int PAD = ((BYTES + 127) / 128) * 128;
};
constexpr int INV_B = TH * TW * 4, HOLD_B = TH * TW;

template <typename T>
inline size_t pass2_smem(int C, int nf, int nm, int nh) {
  size_t s = (size_t)2 * PLANE_B;                   // P planes
  s += XBox<T>::PAD;                                // logit box of the next channel
  s += INV_B + 2 * HOLD_B;                          // 1 / sum e^x of its level; holder bytes of its mid / high group
  s += (size_t)2 * SLOT;                            // holders of the positive terms (per-thread cp.async slots)
  s += (size_t)3 * PR * LP;                         // label tile
  s += (size_t)3 * WS * 4;                          // stencil weights (2 buffers) + their cp.async staging
  s += (size_t)C * 8 + (size_t)2 * nf * 4 + 64;     // tables
  s += (size_t)256 * 2 + 16 + 16;                   // deferred one-hot stencils: work list + counter; mbarrier
  return (s + 127) & ~(size_t)127;
}

// 4 / 2 consecutive staged elements -> fp32
template <typename T>
__device__ __forceinline__ void ld4(const T* p, float (&o)[4]) {
  staged_vec4<T>(p, o);      // 16-byte aligned for fp32 (column 4 + 4 st of the box), 8-byte for the 16-bit types
}
template <typename T>
__device__ __forceinline__ void ld2(const T* p, float& a, float& b) {
  if (sizeof(T) == 4) { const float2 t2 = *reinterpret_cast<const float2*>(p); a = t2.x; b = t2.y; }
  else { a = staged_elem<T>(p, 0); b = staged_elem<T>(p, 1); }
}

template <typename T, bool INLINE_OH, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k3t_pass2(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_inv,
          const __grid_constant__ CUtensorMap map_hold, const T* __restrict__ x, T* __restrict__ grad, int B, int H, int W,
          Hier2 hg, Ws3 ws, float eps, float loss_weight, const float* __restrict__ gscale_ptr, int tiles_x,
          int tiles_per_img) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int C = hg.nf + hg.nm + hg.nh;
  constexpr int XO = XBox<T>::XO, XC = XBox<T>::COLS;
  float* planes = reinterpret_cast<float*>(smem_raw);                                  // [2][PLANE_B / 4]
  T* xbox = reinterpret_cast<T*>(smem_raw + 2 * PLANE_B);                              // [PR][XC] logits of the next channel
  float* ibox = reinterpret_cast<float*>(smem_raw + 2 * PLANE_B + XBox<T>::PAD);       // [TH][TW] 1 / sum e^x
  unsigned char* hbox = reinterpret_cast<unsigned char*>(ibox) + INV_B;                // [2][TH][TW] holder bytes (mid, high group)
  unsigned char* hst = hbox + 2 * HOLD_B;                                              // [2 kinds][NT] x (4 rows x 4 B): positive-term holders
  unsigned char* LT = hst + 2 * SLOT;                                                  // [3][PR][LP]
  float* wsm = reinterpret_cast<float*>(LT + 3 * PR * LP);                             // [2][WS]
  float* wst = wsm + 2 * WS;                                                           // [WS] raw weights of the next channel
  unsigned int* s_order = reinterpret_cast<unsigned int*>(wst + WS);                   // [C]
  unsigned int* s_aux = s_order + C;                                                   // [C]
  int* s_f2m = reinterpret_cast<int*>(s_aux + C);                                      // [nf]
  int* s_f2h = s_f2m + hg.nf;                                                          // [nf]
  int* s_nwork = s_f2h + hg.nf;                                                        // [4] (one used)
  unsigned short* s_work = reinterpret_cast<unsigned short*>(s_nwork + 4);             // [256] block | order index << 7
  unsigned long long* s_mbar = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(s_work + 256) + 15) & ~(uintptr_t)15);               // [1]
  const unsigned int mbar = (unsigned int)__cvta_generic_to_shared(s_mbar);

  const int tid = threadIdx.x;
  {
    // Two instantiations are launched back to back; each image is served by one of them.  Images whose labels are
    // noisy (most 3x8 strips see more than one class, counted by k3f_prep) do their one-hot stencils inline, the
    // others defer them to a work list (see phase B).  Without statistics (generic pass 1) everything defers.
    const int bb = blockIdx.y;          // grid = (tiles of one image, images): the image index stays on the uniform datapath
    const bool noisy = 2u * ws.strips[2 * bb] > ws.strips[2 * bb + 1];
    if (noisy != INLINE_OH) return;
  }
  if (tid == 0) {
    s_nwork[0] = 0;
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const long HW = (long)H * W, BHW = (long)B * HW;
  const int b = blockIdx.y, tile = blockIdx.x;
  const int tyi = tile / tiles_x, ty0 = tyi * TH, tx0 = (tile - tyi * tiles_x) * TW;
  const unsigned char* lab8 = ws.lab8 + (long)b * HW;
  const bool border = ty0 < 2 || ty0 + TH > H - 2 || tx0 < 2 || tx0 + TW > W - 2;

  for (int i = tid; i < C; i += NT) {
    s_order[i] = hg.order[i];
    s_aux[i] = hg.aux[i];
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
  int roff[4];                        // pixel offset of the strip inside one channel plane (clamped into the image)
  unsigned int tc0[4], tc1[4], tc2[4], hmN[4], hhN[4];
  // classes (mod 64, per level) that are the target of some pixel of the block: a set bit sends the channel through
  // the positive-term path; a collision only costs that detour
  unsigned long long presF = 0ull, presM = 0ull, presH = 0ull;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = ty0 + 4 * rq + j;
    rowok[j] = y < H && colok;
    roff[j] = min(y, H - 1) * W + (colok ? xg : 0);
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
  int pidx0, pidx1, xidx0, xidx1;     // plane index / logit-box index of this thread's ring piece(s)
  int goff0, goff1;                   // only for the validity bytes below
  bool in0, in1;
  float hv[6];
  const bool is_strip = tid < 64;
  if (is_strip) {
    const int hrow = tid >> 4, pr = hrow < 2 ? hrow : TH + hrow, strip = tid & 15;
    const int yy = ty0 - 2 + pr, xx = tx0 + 4 * strip;
    pidx0 = pr * PW + 2 + 4 * strip;
    xidx0 = pr * XC + XO + 2 + 4 * strip;
    in0 = yy >= 0 && yy < H && xx < W;
    goff0 = in0 ? yy * W + xx : 0;
  } else {
    const int hl = tid - 64, srow = hl >> 1, side = hl & 1;
    const int yy = ty0 + srow, xx = side ? tx0 + TW : tx0 - 2;
    pidx0 = (2 + srow) * PW + (side ? TW + 2 : 0);
    xidx0 = (2 + srow) * XC + XO + (side ? TW + 2 : 0);
    in0 = yy < H && xx >= 0 && xx < W;
    goff0 = in0 ? yy * W + xx : 0;
  }
  {
    const int crow = (tid >> 1) & 3, pr = crow < 2 ? crow : TH + crow, side = tid & 1;
    const int yy = ty0 - 2 + pr, xx = side ? tx0 + TW : tx0 - 2;
    pidx1 = pr * PW + (side ? TW + 2 : 0);
    xidx1 = pr * XC + XO + (side ? TW + 2 : 0);
    in1 = tid < 8 && yy >= 0 && yy < H && xx >= 0 && xx < W;
    goff1 = in1 ? yy * W + xx : 0;
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

  // ---- staging ----
  const unsigned int hs_base = (unsigned int)__cvta_generic_to_shared(hst + tid * 16);
  const unsigned char* hs_gen = hst + tid * 16;
  const unsigned char* holdb = ws.hold + (long)b * HW;
  const unsigned int xbox_s = (unsigned int)__cvta_generic_to_shared(xbox);
  const unsigned int ibox_s = (unsigned int)__cvta_generic_to_shared(ibox);
  const unsigned int hbox_s = (unsigned int)__cvta_generic_to_shared(hbox);

  // everything phase A of channel ci needs from global memory: three or four TMA boxes issued by one thread (the
  // mbarrier's transaction count covers them), plus the channel's stencil weights (25 + 25 + 1 floats, cp.async)
  auto prefetch = [&](int ci) {
    const unsigned int oe = s_order[ci], ax = s_aux[ci];
    const int kind = oe & 3;
    const unsigned int fl = (oe >> 16) & 0xffu;
    if (tid == 0) {
      const bool hm = (fl & 1u) != 0u, hh = (fl & 4u) && (ax >> 8) != 0xffu;
      mbar_expect_tx(mbar, (unsigned int)(XBox<T>::BYTES + INV_B + (hm ? HOLD_B : 0) + (hh ? HOLD_B : 0)));
      tma_load_3d(xbox_s, &map_x, tx0 - 2 - XO, ty0 - 2, b * C + (int)(oe >> 24), mbar);
      tma_load_3d(ibox_s, &map_inv, tx0, ty0, kind * B + b, mbar);
      if (hm) tma_load_3d(hbox_s, &map_hold, tx0, ty0, (int)(ax & 0xffu) * B + b, mbar);
      if (hh) tma_load_3d(hbox_s + HOLD_B, &map_hold, tx0, ty0, (hg.nm + (int)(ax >> 8)) * B + b, mbar);
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
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(hs_base + q * SLOT + 4 * j), "l"(hp + roff[j]));
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
  int nq = 0, pushed = 0;             // deferred one-hot stencils: queue length (identical in every thread), own push flag

  // phase A of channel (order index) ci: plane (ci & 1), gradient of BCE + CE -> g0
  auto phaseA = [&](int ci) {
    const unsigned int oe = s_order[ci], ax = s_aux[ci];
    const int kind = oe & 3;
    const unsigned int fl = (oe >> 16) & 0xffu, ch = oe >> 24;
    const unsigned int cc = ch * 0x01010101u;
    float* plane = planes + (ci & 1) * (PLANE_B / 4);
    const unsigned char* hrow = hbox + (4 * rq) * TW + 4 * st;      // this thread's 4 x 4 bytes of a holder box
    if (fl & 1u) {       // where does the group's (1 - max) term count: everywhere but at pixels whose target is this mid
      const unsigned int midc = (unsigned int)(hg.nf + (ax & 0xffu)) * 0x01010101u;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        hmN[j] = *reinterpret_cast<const unsigned int*>(hrow + j * TW) | __vcmpeq4(tc1[j], midc);
    }
    if (fl & 4u) {
      const unsigned int high = ax >> 8;
      if (high != 0xffu) {
        const unsigned int highc = (unsigned int)(hg.nf + hg.nm + high) * 0x01010101u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          hhN[j] = *reinterpret_cast<const unsigned int*>(hrow + HOLD_B + j * TW) | __vcmpeq4(tc2[j], highc);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) hhN[j] = 0xffffffffu;
      }
    }
    // ---- ring piece(s): sigmoid only ----
    {
      float xv[4];
      if (is_strip) ld4<T>(xbox + xidx0, xv);
      else {
        ld2<T>(xbox + xidx0, xv[0], xv[1]);
        xv[2] = xv[3] = 0.f;
      }
      if (in0) {
        *reinterpret_cast<float2*>(plane + pidx0) = make_float2(fmaf(sig_only(xv[0]), hv[0], 1e-6f), fmaf(sig_only(xv[1]), hv[1], 1e-6f));
        if (is_strip)
          *reinterpret_cast<float2*>(plane + pidx0 + 2) = make_float2(fmaf(sig_only(xv[2]), hv[2], 1e-6f), fmaf(sig_only(xv[3]), hv[3], 1e-6f));
      }
      if (in1) {
        float a0, a1;
        ld2<T>(xbox + xidx1, a0, a1);
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
      ld4<T>(xbox + (4 * rq + 2 + j) * XC + XO + 2 + 4 * st, xv);
      const float4 iv4 = *reinterpret_cast<const float4*>(ibox + (4 * rq + j) * TW + 4 * st);
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
        const unsigned int zPF = *reinterpret_cast<const unsigned int*>(hs_gen + 4 * j) ^ cc;
        const unsigned int zPM = *reinterpret_cast<const unsigned int*>(hs_gen + SLOT + 4 * j) ^ cc;
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
    const float* pl = planes + (ci & 1) * (PLANE_B / 4) + (4 * rq) * PW + 4 * st;
    const unsigned int ub = kind == 0 ? ublk[0] : (kind == 1 ? ublk[1] : ublk[2]);
    const unsigned int ph = kind == 0 ? pres[0] : (kind == 1 ? pres[1] : pres[2]);
    const float init = ub == cl ? wp[56] : 0.f;
    // mixed labels around the block and this class among them: the block needs the one-hot stencil too.  In images
    // where few strips are mixed it is deferred to a work list that the whole CTA drains 128 blocks at a time (inline
    // it would idle the other lanes of the warp); in noisy images (INLINE_OH) a second sweep right here is cheaper.
    const bool mixed_hit = ub == 0xfeu && ((ph >> (cl & 31)) & 1u);
    if (!INLINE_OH && mixed_hit) { s_work[atomicAdd(s_nwork, 1)] = (unsigned short)(tid | (ci << 7)); pushed = 1; }
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
  // drain 128 items (all threads busy) whenever that many are queued; `flush` takes whatever is left as well.
  // The queue length is carried in a register that every thread updates identically (the number of pushes of a phase B
  // comes out of the barrier that ends it, __syncthreads_count): reading the shared counter here instead would race
  // with the pushes of warps that have already left for the next phase B, and warps that disagree about draining
  // fall out of step at the barriers inside.
  auto drain = [&](bool flush) {
#pragma unroll 1
    while (nq >= NT || (flush && nq > 0)) {
      const int take = nq < NT ? nq : NT;
      unsigned int item = 0xffffffffu;
      if (tid < take) item = s_work[nq - take + tid];
      nq -= take;
      __syncthreads();                                 // everyone has its item: the counter may move
      if (tid == 0) s_nwork[0] = nq;
      if (item != 0xffffffffu) onehot_item(item);
      __syncthreads();                                 // counter visible before the next pushes
    }
  };

  // one copy of each phase in the instruction stream: iteration -1 only runs phase A of channel 0 (its logits were
  // requested above, before the label scan)
#pragma unroll 1
  for (int ci = -1; ci < C; ++ci) {
    if (ci >= 0 && ci + 1 < C) prefetch(ci + 1);
    if (ci >= 0) phaseB(ci);
    if (ci + 1 < C) {
      cp_async_wait<0>();                              // this thread's weight words (and, once, the positive-term holders)
      mbar_wait_or_trap(mbar, (unsigned int)(ci + 1) & 1u);    // the TMA boxes of channel ci + 1 have landed
      phaseA(ci + 1);
    }
    // planes handed over; gradient stores of phase B visible to the CTA; number of deferred stencils queued by phase B
    if (INLINE_OH) __syncthreads();
    else nq += __syncthreads_count(pushed);
    pushed = 0;
    if (!INLINE_OH && ci >= 0) drain(ci + 1 == C);
  }
}

}  // namespace fast3
}  // namespace sh
