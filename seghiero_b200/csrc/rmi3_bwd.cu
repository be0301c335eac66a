// Three-level hierarchical loss, backward side:
//   k3_pass2   second streaming read of the logits, gradient written once:
//              tree-BCE + CE gradient from the per-pixel summaries of pass 1, RMI gradient as a
//              5x5 stencil over P (weights from k3_finalize) for interior pixels
//   k3_frame2  RMI gradient of the 2-pixel image frame (class-dependent stencil weights),
//              added in place
// Analytic RMI backward: see oracle/rmi_taps.py (checked against autograd on the CPU).
#include "rmi3_common.cuh"
#include "rmi3_fast_bwd.cuh"
#include "rmi3_fast_bwd2.cuh"
#include <cstdlib>
#include <cstring>

namespace sh {

// ---------------------------------------------------------------------------------------------
// k3_pass2: grid (tiles_x, tiles_y, B), 512 threads, tile 64x32, rounds of 4 channels.
// A thread owns a 4x2 pixel block for TWO of the round's channels in both phases:
//   phase A: sigmoid/exp, BCE+CE gradient from the summaries (kept in registers), P -> plane
//   phase B: 5x5 stencil over the plane, combine, store the gradient (128-bit, coalesced)
// Planes are double buffered: one __syncthreads per round.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads, 512 / kThreads)
k3_pass2(const T* __restrict__ x, T* __restrict__ grad, int B, int H, int W, Hier3 hg, Ws3 ws, float eps,
         float loss_weight, const float* __restrict__ gscale_ptr, int vec_ok) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int PX = kTH * kTW;
  constexpr int kPlane = (kTH + 4) * kPitch;
  float* planes = reinterpret_cast<float*>(smem_raw);        // [2][kNR][kPlane]
  float* wbuf = planes + 2 * kNR * kPlane;                   // [2][kNR][64]
  float* ivt = wbuf + 2 * kNR * 64;                          // [3][PX]  1/sum e^x per level
  unsigned char* labt = reinterpret_cast<unsigned char*>(ivt + 3 * PX);   // [3][kTH+4][kLabPitch]
  int* htab = reinterpret_cast<int*>(labt + 3 * (kTH + 4) * kLabPitch);   // hierarchy tables (6 KB)
  uint4* xstage = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(htab) + 6144);   // [4][kThreads] raw logits

  const int tid = threadIdx.x;
  const Hier3 h = stage_hier(hg, htab, tid, kThreads);
  __syncthreads();
  const int C = h.nf + h.nm + h.nh;
  const int b = blockIdx.z, y0 = blockIdx.y * kTH, x0 = blockIdx.x * kTW;
  const long HW = (long)H * W;
  const unsigned char* lab8 = ws.lab8 + (long)b * HW;
  const unsigned char* flg = ws.flags + (long)b * HW;
  const float gscale = *gscale_ptr;
  const T* xb = x + (long)b * C * HW;
  T* gb = grad + (long)b * C * HW;

  // thread -> (channel pair, 4x2 pixel block)
  const int pair = tid / (kThreads / 2), u = tid % (kThreads / 2), rp = u >> 4, st = u & 15;
  const int ty = 2 * rp, tx = 4 * st;
  const int xg = x0 + tx;
  int nvalid = W - xg;
  nvalid = nvalid > 4 ? 4 : (nvalid < 0 ? 0 : nvalid);
  const bool st_al = vec_ok && ((W & 3) == 0);
  bool row_ok[2];
  long own_off[2];
  unsigned int tf4[2], tm4[2], th4[2], hpf4[2], hpm4[2], ucode[3][2];
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    const int y = y0 + ty + o;
    row_ok[o] = y < H && nvalid > 0;
    own_off[o] = (long)y * W + xg;
    tf4[o] = tm4[o] = th4[o] = 0xffffffffu;
    hpf4[o] = hpm4[o] = 0;
#pragma unroll
    for (int l = 0; l < 3; ++l) ucode[l][o] = 0xffffffffu;
    if (row_ok[o]) {
      tf4[o] = load4_u8(lab8 + own_off[o], nvalid, st_al) | (nvalid < 4 ? (0xffffffffu << (8 * nvalid)) : 0u);
      const unsigned int fl4 = load4_u8(flg + own_off[o], nvalid, st_al);
      hpf4[o] = load4_u8(ws.hold + ((size_t)(h.nm + h.nh) * B + b) * HW + own_off[o], nvalid, st_al);
      hpm4[o] = load4_u8(ws.hold + ((size_t)(h.nm + h.nh + 1) * B + b) * HW + own_off[o], nvalid, st_al);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned int t = (tf4[o] >> (8 * k)) & 0xffu, fl = (fl4 >> (8 * k)) & 0xffu;
        const unsigned int m = t != SH_IGNORE ? (unsigned)h.f2m[t] : 0xffu, g = t != SH_IGNORE ? (unsigned)h.f2h[t] : 0xffu;
        tm4[o] = (tm4[o] & ~(0xffu << (8 * k))) | (m << (8 * k));
        th4[o] = (th4[o] & ~(0xffu << (8 * k))) | (g << (8 * k));
        const unsigned int r3[3] = {t != SH_IGNORE ? t : 0u, t != SH_IGNORE ? m : 0u, t != SH_IGNORE ? g : 0u};
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          unsigned int code = 0xffu;
          if (k < nvalid && (fl & kFlagInterior)) code = (fl & (kFlagUniF << l)) ? r3[l] : 0xfeu;
          ucode[l][o] = (ucode[l][o] & ~(0xffu << (8 * k))) | (code << (8 * k));
        }
      }
    }
  }
  load_label_tile(labt, kTH, lab8, H, W, y0, x0, h, tid, kThreads);
  for (int i = tid; i < 3 * PX / 4; i += kThreads) {
    const int l = i / (PX / 4), rem = i - l * (PX / 4), r = rem / kStrips, s4 = (rem % kStrips) * 4;
    const int y = y0 + r, xx = x0 + s4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < H && xx < W) {
      const float* ip = ws.inv + ((size_t)l * B + b) * HW + (long)y * W + xx;
      if (st_al && xx + 4 <= W) v = *reinterpret_cast<const float4*>(ip);
      else { float t4[4] = {0.f, 0.f, 0.f, 0.f}; for (int k = 0; k < 4 && xx + k < W; ++k) t4[k] = ip[k]; v = make_float4(t4[0], t4[1], t4[2], t4[3]); }
    }
    *reinterpret_cast<float4*>(ivt + l * PX + r * kTW + s4) = v;
  }
  // halo slots (same positions in every plane): rows 0,1 and kTH+2,kTH+3 of the plane, 2 columns either side
  constexpr int nhalo = kPlane - kTH * kTW;
  constexpr int kHS = (nhalo + kThreads - 1) / kThreads;
  int h_sidx[kHS];
  long h_goff[kHS];
  bool h_valid[kHS];
#pragma unroll
  for (int q = 0; q < kHS; ++q) {
    const int e = tid + q * kThreads;
    h_sidx[q] = -1; h_goff[q] = -1; h_valid[q] = false;
    if (e < nhalo) {
      int r, j;
      if (e < 4 * kPitch) { const int rr = e / kPitch; r = rr < 2 ? rr : kTH + rr; j = e % kPitch; }
      else { const int e2 = e - 4 * kPitch; r = 2 + (e2 >> 2); const int qq = e2 & 3; j = qq < 2 ? qq : kTW + qq; }
      h_sidx[q] = r * kPitch + j;
      const int yy = y0 - 2 + r, xx = x0 - 2 + j;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) { h_goff[q] = (long)yy * W + xx; h_valid[q] = lab8[h_goff[q]] != SH_IGNORE; }
    }
  }
  __syncthreads();
  unsigned int pres[3] = {0u, 0u, 0u};
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    bool any_nu = false;
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
      for (int k = 0; k < 4; ++k) any_nu |= ((ucode[l][o] >> (8 * k)) & 0xffu) == 0xfeu;
    if (any_nu) {
      for (int rr = 0; rr < 6; ++rr) {
        const unsigned int* row = reinterpret_cast<const unsigned int*>(labt + ((l * (kTH + 4)) + ty + rr) * kLabPitch + tx);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const unsigned int wd = row[q];
          pres[l] |= (1u << (wd & 31)) | (1u << ((wd >> 8) & 31)) | (1u << ((wd >> 16) & 31)) | (1u << ((wd >> 24) & 31));
        }
      }
    }
  }
  // byte masks used by the gradient logic (0xff per pixel byte where the condition holds)
  unsigned int valid4[2], cth4[2];
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    valid4[o] = __vcmpne4(tf4[o], 0xffffffffu);
    cth4[o] = ((th4[o] & valid4[o]) + (unsigned int)(h.nf + h.nm) * 0x01010101u) | ~valid4[o];   // channel of the high target
  }

  const float nv = fmaxf((float)ws.counts[0], 1.0f);
  const float wF = 2.5f * loss_weight * gscale / (nv * (float)h.nf);
  const float wM = 2.5f * loss_weight * gscale / (nv * (float)h.nm);
  const float wH = 2.5f * loss_weight * gscale / (nv * (float)h.nh);
  const float wCE = loss_weight * gscale / ((float)B * (float)HW);

  const int nrounds = (C + kNR - 1) / kNR;
  float4* gstage = reinterpret_cast<float4*>(xstage + 4 * kThreads);                        // [4][kThreads] BCE+CE gradient
  unsigned int* hstage = reinterpret_cast<unsigned int*>(gstage + 4 * kThreads);            // [kNR][kHS][kThreads]
  // next round's logits travel global -> shared with cp.async; ragged strips load synchronously on use
  bool fast_own[2];
#pragma unroll
  for (int o = 0; o < 2; ++o) fast_own[o] = vec_ok && row_ok[o] && nvalid == 4;
  const bool halo_async_ok = sizeof(T) == 4 || vec_ok;
  auto prefetch = [&](int r) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int c = r * kNR + 2 * pair + jj;
      if (c < C) {
        const T* xc = xb + (long)c * HW;
#pragma unroll
        for (int o = 0; o < 2; ++o)
          if (fast_own[o]) cp_async_vec4<T>(xstage + (jj * 2 + o) * kThreads + tid, xc + own_off[o]);
      }
    }
#pragma unroll
    for (int j = 0; j < kNR; ++j) {
      const int c = r * kNR + j;
#pragma unroll
      for (int q = 0; q < kHS; ++q)
        if (c < C && h_goff[q] >= 0 && halo_async_ok)
          cp_async_elem<T>(hstage + (j * kHS + q) * kThreads + tid, xb + (long)c * HW, h_goff[q]);
    }
    cp_async_commit();
  };

  prefetch(0);
  for (int r = 0; r < nrounds; ++r) {
    const int buf = r & 1;
    // ======================= phase A =======================
    cp_async_wait_all();
#pragma unroll 1     // rolled on purpose: the fully unrolled body overflows the instruction cache
    for (int jj = 0; jj < 2; ++jj) {
      const int c = r * kNR + 2 * pair + jj;
      if (c >= C) break;
      const int lvl = c < h.nf ? 0 : (c < h.nf + h.nm ? 1 : 2);
      const int cl = lvl == 0 ? c : (lvl == 1 ? c - h.nf : c - h.nf - h.nm);
      float* plane = planes + (buf * kNR + 2 * pair + jj) * kPlane;
      const int mid = lvl == 0 ? h.f2m[cl] : (lvl == 1 ? cl : -1);
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        float s[4], v[4], pk[4], xv[4];
        if (fast_own[o]) staged_vec4<T>(xstage + (jj * 2 + o) * kThreads + tid, xv);
        else if (row_ok[o]) load_n<T, 4>(xb + (long)c * HW, own_off[o], own_off[o] - xg + W, vec_ok != 0, xv);
        else { xv[0] = xv[1] = xv[2] = xv[3] = 0.f; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const SigExp se = sig_exp(xv[k]);
          s[k] = se.s; v[k] = se.v;
          const bool valid = ((tf4[o] >> (8 * k)) & 0xffu) != SH_IGNORE;
          pk[k] = (row_ok[o] && k < nvalid) ? ((valid ? se.s : 0.f) + 1e-6f) : 0.f;
        }
        *reinterpret_cast<float2*>(plane + (ty + o + 2) * kPitch + tx + 2) = make_float2(pk[0], pk[1]);
        *reinterpret_cast<float2*>(plane + (ty + o + 2) * kPitch + tx + 4) = make_float2(pk[2], pk[3]);
        // tree BCE + CE gradient of this channel at these 4 pixels, with per-byte SIMD masks:
        //   negative terms: own fine term (f != tf), mid-level term held by this channel (m != tm),
        //                   high-level term(s) held by this channel (h != th)
        //   positive terms: this channel holds min(A_tf, B_tm) / min(C_h.., B_tm) / is the high target
        float g0[4];
        const unsigned int cc = (unsigned int)c * 0x01010101u;
        const unsigned int v4 = valid4[o];
        unsigned int negO = 0, negM = 0, negH = 0, posH = 0;
        float extra[4] = {0.f, 0.f, 0.f, 0.f};
        if (lvl == 0) negO = __vcmpne4(tf4[o], cc) & v4;
        if (row_ok[o]) {
          if (lvl != 2) {
            const unsigned int hm = load4_u8(ws.hold + ((size_t)mid * B + b) * HW + own_off[o], nvalid, st_al);
            negM = __vcmpeq4(hm, cc) & __vcmpne4(tm4[o], (unsigned int)mid * 0x01010101u) & v4;
            const int q0 = h.mh_ptr[mid], q1e = h.mh_ptr[mid + 1];
            for (int q = q0; q < q1e; ++q) {
              const int hi = h.mh_idx[q];
              const unsigned int hh = load4_u8(ws.hold + ((size_t)(h.nm + hi) * B + b) * HW + own_off[o], nvalid, st_al);
              const unsigned int m4 = __vcmpeq4(hh, cc) & __vcmpne4(th4[o], (unsigned int)hi * 0x01010101u) & v4;
              if (q == q0) negH = m4;
              else {       // a mid that feeds several highs (non-tree maps): keep the multiplicity
#pragma unroll
                for (int k = 0; k < 4; ++k) extra[k] += (m4 >> (8 * k)) & 1u ? wH : 0.f;
              }
            }
          } else {
            const unsigned int hh = load4_u8(ws.hold + ((size_t)(h.nm + cl) * B + b) * HW + own_off[o], nvalid, st_al);
            posH = __vcmpeq4(cth4[o], cc) & v4;
            negH = __vcmpeq4(hh, cc) & ~posH & v4;
          }
        }
        const unsigned int posF = __vcmpeq4(hpf4[o], cc) & v4, posM = __vcmpeq4(hpm4[o], cc) & v4;
        const unsigned int pos_any = posF | posM | posH;
        const float4 iv4 = *reinterpret_cast<const float4*>(ivt + lvl * PX + (ty + o) * kTW + tx);
        const float ivk[4] = {iv4.x, iv4.y, iv4.z, iv4.w};
        const unsigned int tgt4 = __vcmpeq4(lvl == 0 ? tf4[o] : (lvl == 1 ? tm4[o] : th4[o]), (unsigned int)cl * 0x01010101u) & v4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned int bit = 1u << (8 * k);
          const float q1 = 1.0f - s[k];
          float dneg = extra[k];
          dneg += (negO & bit) ? wF : 0.f;
          dneg += (negM & bit) ? wM : 0.f;
          dneg += (negH & bit) ? wH : 0.f;
          float ds = dneg * rcp(q1 + eps);
          if (pos_any & bit) {
            float dpos = (posF & bit) ? wF : 0.f;
            dpos += (posM & bit) ? wM : 0.f;
            dpos += (posH & bit) ? wH : 0.f;
            ds -= dpos * rcp(s[k] + eps);
          }
          float ce = v[k] * ivk[k] - ((tgt4 & bit) ? 1.f : 0.f);
          ce = (v4 & bit) ? wCE * ce : 0.f;
          g0[k] = fmaf(ds, q1 * s[k], ce);
        }
        gstage[(jj * 2 + o) * kThreads + tid] = make_float4(g0[0], g0[1], g0[2], g0[3]);
      }
    }
    // halo of the 4 planes + stencil weights of the round
#pragma unroll
    for (int j = 0; j < kNR; ++j) {
#pragma unroll
      for (int q = 0; q < kHS; ++q)
        if (r * kNR + j < C && h_sidx[q] >= 0) {
          float p = 0.f;
          if (h_goff[q] >= 0) {
            const float hx = halo_async_ok ? staged_elem<T>(hstage + (j * kHS + q) * kThreads + tid, h_goff[q])
                                           : to_f32<T>(xb[(long)(r * kNR + j) * HW + h_goff[q]]);
            p = (h_valid[q] ? sig_exp(hx).s : 0.f) + 1e-6f;
          }
          planes[(buf * kNR + j) * kPlane + h_sidx[q]] = p;
        }
    }
    for (int q = tid; q < kNR * 64; q += kThreads) {
      const int j = q >> 6, c = r * kNR + j;
      if (c < C) wbuf[(buf * kNR + j) * 64 + (q & 63)] = ws.wts[((size_t)b * C + c) * 64 + (q & 63)];
    }
    if (r + 1 < nrounds) prefetch(r + 1);
    __syncthreads();

    // ======================= phase B =======================
#pragma unroll 1
    for (int jj = 0; jj < 2; ++jj) {
      const int c = r * kNR + 2 * pair + jj;
      if (c >= C) break;
      const int lvl = c < h.nf ? 0 : (c < h.nf + h.nm ? 1 : 2);
      const int cl = lvl == 0 ? c : (lvl == 1 ? c - h.nf : c - h.nf - h.nm);
      const float* plane = planes + (buf * kNR + 2 * pair + jj) * kPlane;
      const float* wb = wbuf + (buf * kNR + 2 * pair + jj) * 64;
      float w1[25];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const float4 t4 = *reinterpret_cast<const float4*>(wb + 4 * q);
        w1[4 * q] = t4.x; w1[4 * q + 1] = t4.y; w1[4 * q + 2] = t4.z; w1[4 * q + 3] = t4.w;
      }
      w1[24] = wb[24];
      float dP[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float pc[2][4];
#pragma unroll
      for (int rr = 0; rr < 6; ++rr) {      // plane rows ty+rr  <->  image rows y0+ty+rr-2
        const float4* p4 = reinterpret_cast<const float4*>(plane + (ty + rr) * kPitch + tx);
        const float4 a = p4[0], bq = p4[1];
        const float w[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int dyi = rr - o;            // = dy + 2
          if (dyi < 0 || dyi > 4) continue;
          if (dyi == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) pc[o][k] = w[k + 2];
          }
#pragma unroll
          for (int dx = 0; dx < 5; ++dx)
#pragma unroll
            for (int k = 0; k < 4; ++k) dP[o][k] = fmaf(w1[dyi * 5 + dx], w[k + dx], dP[o][k]);
        }
      }
      const float w2full = wb[50];
      const unsigned int pr = lvl == 0 ? pres[0] : (lvl == 1 ? pres[1] : pres[2]);
      const bool want = (pr >> (cl & 31)) & 1u;
      float add[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      if (want) {
#pragma unroll
        for (int rr = 0; rr < 6; ++rr) {
          const unsigned int* row = reinterpret_cast<const unsigned int*>(
              labt + ((lvl * (kTH + 4)) + ty + rr) * kLabPitch + tx);
          const unsigned long long wbits = (unsigned long long)row[0] | ((unsigned long long)row[1] << 32);
          float mt[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) mt[q] = (byte_of(wbits, q) == (unsigned int)cl) ? 1.f : 0.f;
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            const int dyi = rr - o;
            if (dyi < 0 || dyi > 4) continue;
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
              const float w2 = wb[25 + dyi * 5 + dx];
#pragma unroll
              for (int k = 0; k < 4; ++k) add[o][k] = fmaf(w2, mt[k + dx], add[o][k]);
            }
          }
        }
      }
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const unsigned int uc = lvl == 0 ? ucode[0][o] : (lvl == 1 ? ucode[1][o] : ucode[2][o]);
        float g[4];
        const float4 g4 = gstage[(jj * 2 + o) * kThreads + tid];
        const float g0[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned int code = (uc >> (8 * k)) & 0xffu;
          float d = dP[o][k];
          if (code == (unsigned int)cl) d += w2full;
          else if (code == 0xfeu) d += add[o][k];
          const bool valid = ((tf4[o] >> (8 * k)) & 0xffu) != SH_IGNORE;
          const float s = pc[o][k] - 1e-6f;            // valid pixels: P = s + 1e-6
          const float q = (valid && code != 0xffu) ? s * (1.0f - s) * gscale : 0.f;
          g[k] = fmaf(q, d, g0[k]);
        }
        if (row_ok[o]) store_n<T, 4>(gb + (long)c * HW, own_off[o], own_off[o] - xg + W, vec_ok != 0, g);
      }
    }
  }
}

// stencil of one frame pixel over the staged band segment; WITH_L = false when the segment holds no pixel of the class
template <bool WITH_L>
__device__ __forceinline__ float frame2_stencil(const BandSeg& bs, const float* w, int va, int i, int u, int N, bool is_row) {
  float dP = 0.f;
#pragma unroll
  for (int dv = -2; dv <= 2; ++dv) {
    const int vv = va + dv;
    if (vv < 0 || vv > 3) continue;
#pragma unroll
    for (int du = -2; du <= 2; ++du) {
      if (u + du < 0 || u + du >= N) continue;
      const int t = is_row ? (dv + 2) * 5 + du + 2 : (du + 2) * 5 + dv + 2;
      dP = fmaf(w[t], bs.P[vv][i + 2 + du], dP);
      if (WITH_L) dP = fmaf(w[25 + t], bs.L[vv][i + 2 + du], dP);
    }
  }
  return dP;
}

// grid (B*C, nseg), block 256: the frame pixels of one (b, c) plane, side by side; the 4-pixel band segments the
// stencil touches are staged in shared memory
template <typename T>
__global__ void __launch_bounds__(256) k3_frame2(T* __restrict__ grad, int B, int H, int W, Hier3 h, Ws3 ws,
                                                 const float* __restrict__ bandR, const float* __restrict__ bandC,
                                                 const float* __restrict__ gscale_ptr) {
  __shared__ float fw[25 * 50];
  __shared__ BandSeg bs;
  const int C = h.nf + h.nm + h.nh;
  const int bc = blockIdx.x, b = bc / C, c = bc % C;
  const int seg = blockIdx.y, nseg = gridDim.y;
  const int lvl = c < h.nf ? 0 : (c < h.nf + h.nm ? 1 : 2);
  const int cl = lvl == 0 ? c : (lvl == 1 ? c - h.nf : c - h.nf - h.nm);
  const long HW = (long)H * W;
  const float* bR = bandR + (size_t)bc * 8 * W;
  const float* bC = bandC + (size_t)bc * 8 * H;
  const unsigned char* lb = ws.labB + ((size_t)b * 3 + lvl) * 8 * ((size_t)W + H);
  T* gc = grad + ((long)b * C + c) * HW;
  const float gscale = *gscale_ptr;
  const int tid = threadIdx.x;
  for (int i = tid; i < 25 * 50; i += 256) fw[i] = ws.fwts[(size_t)bc * 25 * 50 + i];
#pragma unroll 1
  for (int side = 0; side < 4; ++side) {
    const bool is_row = side < 2;
    // rows 0,1,H-2,H-1 over all columns; columns 0,1,W-2,W-1 over the rows in between
    const int N = is_row ? W : H, first = is_row ? 0 : 2, len = is_row ? W : H - 4;
    const int per = (len + nseg - 1) / nseg, lo = first + seg * per, hi = min(first + len, lo + per);
    for (int u0 = lo; u0 < hi; u0 += kSegMax) {
      const int n = min(kSegMax, hi - u0);
      __syncthreads();
      const bool saw = stage_band(bs, side, u0, n, bR, bC, lb, cl, H, W, tid, 256);
      const bool any = __syncthreads_or(saw) != 0;       // a pixel of this class in the segment? (else L == 0)
      for (int idx = tid; idx < 2 * n; idx += 256) {
        const int line = idx / n, i = idx - line * n, u = u0 + i;
        const int va = line + 2 * (side & 1);
        const float s = bs.P[va][i + 2] - 1e-6f;              // void pixels: P = 1e-6, no gradient
        if (s == 0.f) continue;
        const int vg = (side & 1) ? (is_row ? H : W) - 4 + va : va;
        const int yy = is_row ? vg : u, xx = is_row ? u : vg;
        const float* w = fw + (axis_class(yy, H) * 5 + axis_class(xx, W)) * 50;
        const float dP = any ? frame2_stencil<true>(bs, w, va, i, u, N, is_row)
                             : frame2_stencil<false>(bs, w, va, i, u, N, is_row);
        const float add = dP * gscale * s * (1.0f - s);
        const long off = (long)yy * W + xx;
        gc[off] = from_f32<T>(to_f32<T>(gc[off]) + add);
      }
    }
  }
}

static size_t pass2_smem_bytes() {
  size_t s = (size_t)2 * kNR * (kTH + 4) * kPitch * 4 + 2 * kNR * 64 * 4 + 3 * kTH * kTW * 4 +
             3 * (kTH + 4) * kLabPitch + 6144 /* hierarchy tables */;
  s = (s + 15) & ~(size_t)15;
  constexpr int kHS = ((kTH + 4) * kPitch - kTH * kTW + kThreads - 1) / kThreads;
  s += (size_t)2 * 4 * kThreads * 16 + (size_t)kNR * kHS * kThreads * 4;   // cp.async staging + parked BCE gradient
  return (s + 15) & ~(size_t)15;
}

bool fast_path_ok(const void* x, const void* grad, int elem, int H, int W, int nf, int nm, int nh, int fast_tab_ok);
size_t fast_bwd_smem(int C, int nf, int nm, int nh) { return fast2::pass2_smem(C, nf, nm, nh); }

// TMA tensor maps: tma.cuh
// which pass-2 kernel serves the fast path: "tma" (default) or "legacy" (SEGHIERO_B200_PASS2, for A/B measurements)
static bool pass2_want_tma() {
  const char* e = std::getenv("SEGHIERO_B200_PASS2");
  return e == nullptr || std::strcmp(e, "legacy") != 0;
}

template <typename T>
static int run_backward3(const void* x, void* grad, int B, int H, int W, const Hier3& h, const Ws3& ws,
                         const float* bandR, const float* bandC, float eps, float lw, const float* gscale, int stages,
                         int fast_tab_ok, cudaStream_t st) {
  const int C = h.nf + h.nm + h.nh;
  const bool vec_ok = ((W & 3) == 0) && ((uintptr_t)x % (4 * sizeof(T)) == 0) && ((uintptr_t)grad % (4 * sizeof(T)) == 0);
  const size_t fsmem = fast2::pass2_smem(C, h.nf, h.nm, h.nh);
  // the tiled backward kernel takes any tree-shaped hierarchy up to 254 channels, whichever pass 1 ran
  const bool fast = fast_tab_ok && (W & 3) == 0 && C <= 254 && (uintptr_t)x % (4 * sizeof(T)) == 0 &&
                    (uintptr_t)grad % (4 * sizeof(T)) == 0 && fsmem <= 227 * 1024;
  if ((stages & 1) && fast) {
    fast2::Hier2 fh;
    fh.nf = h.nf; fh.nm = h.nm; fh.nh = h.nh; fh.f2m = h.f2m; fh.f2h = h.f2h;
    fh.order = h.order + C; fh.aux = h.order + 2 * C;
    const int tiles_x = (W + fast2::TW - 1) / fast2::TW, tiles_y = (H + fast2::TH - 1) / fast2::TH;
    // TMA form: needs 16-byte row pitches for the three maps (logits, 1/sum e^x, holder bytes)
    CUtensorMap mx, mi, mh;
    const size_t tsmem = fast3::pass2_smem<T>(C, h.nf, h.nm, h.nh);
    const char* ctas_e = std::getenv("SEGHIERO_B200_P2CTAS");          // "4": the 128-register variant (A/B measurements)
    const bool four = ctas_e != nullptr && std::atoi(ctas_e) == 4;
    if (pass2_want_tma() && tsmem <= 72 * 1024 &&      /* 3 CTAs per SM */
        make_plane_map(&mx, TmaType<T>::v, (int)sizeof(T), x, W, H, (long)B * C, fast3::XBox<T>::COLS, fast3::PR) &&
        make_plane_map(&mi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ws.inv, W, H, 3L * B, fast3::TW, fast3::TH) &&
        make_plane_map(&mh, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, ws.hold, W, H, (long)(h.nm + h.nh + 2) * B, fast3::TW, fast3::TH)) {
      // 3 CTAs per SM with the compiler's own register allocation (default), or 4 under a 128-register cap
      void (*t0)(CUtensorMap, CUtensorMap, CUtensorMap, const T*, T*, int, int, int, fast2::Hier2, Ws3, float, float,
                 const float*, int, int) = four ? fast3::k3t_pass2<T, false, 4> : fast3::k3t_pass2<T, false, 3>;
      void (*t1)(CUtensorMap, CUtensorMap, CUtensorMap, const T*, T*, int, int, int, fast2::Hier2, Ws3, float, float,
                 const float*, int, int) = four ? fast3::k3t_pass2<T, true, 4> : fast3::k3t_pass2<T, true, 3>;
      cudaFuncSetAttribute(t0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
      cudaFuncSetAttribute(t1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
      t0<<<dim3(tiles_x * tiles_y, B), fast3::NT, tsmem, st>>>(mx, mi, mh, (const T*)x, (T*)grad, B, H, W, fh, ws, eps, lw, gscale,
                                                         tiles_x, tiles_x * tiles_y);
      SH_CHECK_LAUNCH();
      t1<<<dim3(tiles_x * tiles_y, B), fast3::NT, tsmem, st>>>(mx, mi, mh, (const T*)x, (T*)grad, B, H, W, fh, ws, eps, lw, gscale,
                                                         tiles_x, tiles_x * tiles_y);
      SH_CHECK_LAUNCH();
      if (stages & 2) {
        k3_frame2<T><<<dim3(B * C, ws.nseg), 256, 0, st>>>((T*)grad, B, H, W, h, ws, bandR, bandC, gscale);
        SH_CHECK_LAUNCH();
      }
      return SH_OK;
    }
    auto kern0 = fast2::k3f_pass2<T, false>;
    auto kern1 = fast2::k3f_pass2<T, true>;
    cudaFuncSetAttribute(kern0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
    cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
    // every image is served by exactly one of the two (label-noise statistics of k3f_prep, read on the device)
    kern0<<<B * tiles_x * tiles_y, fast2::NT, fsmem, st>>>((const T*)x, (T*)grad, B, H, W, fh, ws, eps, lw, gscale, tiles_x,
                                                          tiles_x * tiles_y);
    SH_CHECK_LAUNCH();
    kern1<<<B * tiles_x * tiles_y, fast2::NT, fsmem, st>>>((const T*)x, (T*)grad, B, H, W, fh, ws, eps, lw, gscale, tiles_x,
                                                          tiles_x * tiles_y);
    SH_CHECK_LAUNCH();
  } else if (stages & 1) {
    // a fast forward pass skipped the per-pixel flags this kernel reads unless it knew the backward would land here
    if (fast_path_ok(x, nullptr, (int)sizeof(T), H, W, h.nf, h.nm, h.nh, fast_tab_ok) && fsmem <= 227 * 1024)
      return SH_ERR_UNSUPPORTED;     // only reachable with a gradient buffer that is not 16-byte aligned
    const size_t smem = pass2_smem_bytes();
    auto kern = k3_pass2<T>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 g2(ws.tiles_x, ws.tiles_y, B);
    kern<<<g2, kThreads, smem, st>>>((const T*)x, (T*)grad, B, H, W, h, ws, eps, lw, gscale, vec_ok ? 1 : 0);
    SH_CHECK_LAUNCH();
  }
  if (stages & 2) {
    k3_frame2<T><<<dim3(B * C, ws.nseg), 256, 0, st>>>((T*)grad, B, H, W, h, ws, bandR, bandC, gscale);
    SH_CHECK_LAUNCH();
  }
  return SH_OK;
}

}  // namespace sh

extern "C" {

// Which pass-2 kernel sh_rmi3_backward launches for this problem: 0 = generic k3_pass2, 1 = tiled cp.async kernel
// (k3f_pass2), 2 = tiled TMA kernel (k3t_pass2).  Pure host arithmetic (diagnostics, bench.py's stage names).
int sh_rmi3_pass2_kind(const void* logits, const void* grad, int dtype, int H, int W, int nf, int nm, int nh,
                       int fast_tab_ok) {
  const int es = dtype == SH_DT_F32 ? 4 : 2, C = nf + nm + nh;
  const bool fast = fast_tab_ok && (W & 3) == 0 && C <= 254 && (uintptr_t)logits % (4 * es) == 0 &&
                    (uintptr_t)grad % (4 * es) == 0 && sh::fast2::pass2_smem(C, nf, nm, nh) <= 227 * 1024;
  if (!fast) return 0;
  const size_t tsmem = es == 4 ? sh::fast3::pass2_smem<float>(C, nf, nm, nh) : sh::fast3::pass2_smem<__half>(C, nf, nm, nh);
  const bool tma = sh::pass2_want_tma() && tsmem <= 72 * 1024 && ((long)W * es) % 16 == 0 && W % 16 == 0 &&
                   ((uintptr_t)logits & 15) == 0 && sh::encode_tiled_fn() != nullptr;
  return tma ? 2 : 1;
}

int sh_rmi3_backward(const void* logits, int dtype, void* grad, int B, int H, int W, int nf, int nm, int nh,
                     const int* hier_tab, int n_mh, int fast_tab_ok, float loss_weight, void* workspace,
                     const float* grad_out, int stages, void* stream) {
  if (B <= 0 || H < 8 || W < 8 || nf + nm + nh > 254) return SH_ERR_BAD_ARG;
  sh::Ws3 ws = sh::ws3_layout(workspace, B, H, W, nf, nm, nh);
  const float* bandR = (const float*)((unsigned char*)workspace + ws.bytes);
  const float* bandC = bandR + (size_t)B * (nf + nm + nh) * 8 * W;
  sh::Hier3 h = sh::hier3_from_tab(hier_tab, nf, nm, nh, n_mh);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case SH_DT_F32: return sh::run_backward3<float>(logits, grad, B, H, W, h, ws, bandR, bandC, 1e-6f, loss_weight, grad_out, stages, fast_tab_ok, st);
    case SH_DT_BF16: return sh::run_backward3<__nv_bfloat16>(logits, grad, B, H, W, h, ws, bandR, bandC, 1e-6f, loss_weight, grad_out, stages, fast_tab_ok, st);
    case SH_DT_F16: return sh::run_backward3<__half>(logits, grad, B, H, W, h, ws, bandR, bandC, 1e-6f, loss_weight, grad_out, stages, fast_tab_ok, st);
  }
  return SH_ERR_UNSUPPORTED;
}

}  // extern "C"
