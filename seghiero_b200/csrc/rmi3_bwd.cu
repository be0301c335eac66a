// Three-level hierarchical loss, backward side:
//   k3_pass2   second streaming read of the logits, gradient written once:
//              tree-BCE + CE gradient from the per-pixel summaries of pass 1, RMI gradient as a
//              5x5 stencil over P (weights from k3_finalize) for interior pixels
//   k3_frame2  RMI gradient of the 2-pixel image frame (class-dependent stencil weights),
//              added in place
// Analytic RMI backward: see oracle/rmi_taps.py (checked against autograd on the CPU).
#include "rmi3_common.cuh"

namespace sh {

__device__ __forceinline__ unsigned int load4_u8(const unsigned char* p, int nvalid, bool aligned) {
  if (aligned && nvalid == 4) return *reinterpret_cast<const unsigned int*>(p);
  unsigned int r = 0;
  for (int k = 0; k < nvalid; ++k) r |= (unsigned int)p[k] << (8 * k);
  return r;
}

template <typename T>
__global__ void __launch_bounds__(256, 2)
k3_pass2(const T* __restrict__ x, T* __restrict__ grad, int B, int H, int W, Hier3 h, Ws3 ws, float eps,
         float loss_weight, const float* __restrict__ gscale_ptr, int vec_ok) {
  constexpr int TH = 16;
  __shared__ __align__(16) float plane[(TH + 4) * kPitch];
  __shared__ __align__(16) float wbuf[64];
  __shared__ __align__(8) unsigned char labt[3 * (TH + 4) * kLabPitch];

  const int C = h.nf + h.nm + h.nh;
  const int b = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * kTW;
  const long HW = (long)H * W;
  const int tid = threadIdx.x;
  const int ty = tid / kStrips, tx = (tid % kStrips) * 4;
  const int y = y0 + ty, xg = x0 + tx;
  const unsigned char* lab8 = ws.lab8 + (long)b * HW;
  const unsigned char* flg = ws.flags + (long)b * HW;
  const float gscale = *gscale_ptr;

  for (int e = tid; e < (TH + 4) * kPitch; e += 256) {
    const int r = e / kPitch, j = e - r * kPitch;
    const int yy = y0 - 2 + r, xx = x0 - 2 + j;
    unsigned char f = 0xff, m = 0xff, g = 0xff;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const int t = lab8[(long)yy * W + xx];
      f = m = g = 0;
      if (t != SH_IGNORE) { f = (unsigned char)t; m = (unsigned char)h.f2m[t]; g = (unsigned char)h.f2h[t]; }
    }
    labt[(0 * (TH + 4) + r) * kLabPitch + j] = f;
    labt[(1 * (TH + 4) + r) * kLabPitch + j] = m;
    labt[(2 * (TH + 4) + r) * kLabPitch + j] = g;
  }

  int tf[4], tm[4], thh[4];
  bool inimg[4], interior[4];
  unsigned int ulab[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};
  unsigned int nonuni[3] = {0u, 0u, 0u};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    inimg[k] = (y < H) && (xg + k < W);
    int t = SH_IGNORE, fl = 0;
    if (inimg[k]) { t = lab8[(long)y * W + xg + k]; fl = flg[(long)y * W + xg + k]; }
    tf[k] = t;
    tm[k] = t != SH_IGNORE ? h.f2m[t] : SH_IGNORE;
    thh[k] = t != SH_IGNORE ? h.f2h[t] : SH_IGNORE;
    interior[k] = (fl & kFlagInterior) != 0;
    const int r3[3] = {t != SH_IGNORE ? t : 0, t != SH_IGNORE ? tm[k] : 0, t != SH_IGNORE ? thh[k] : 0};
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      if (interior[k]) {
        if (fl & (kFlagUniF << l)) ulab[l] = (ulab[l] & ~(0xffu << (8 * k))) | ((unsigned int)r3[l] << (8 * k));
        else nonuni[l] |= 1u << k;
      }
    }
  }
  const bool row_ok = y < H;
  int nvalid = row_ok ? W - xg : 0;
  nvalid = nvalid > 4 ? 4 : (nvalid < 0 ? 0 : nvalid);
  const long own_off = (long)y * W + xg;
  const bool st_al = vec_ok && ((W & 3) == 0);
  // summaries of pass 1
  unsigned int hp_f = 0, hp_m = 0;
  float iv[3][4];
#pragma unroll
  for (int l = 0; l < 3; ++l)
#pragma unroll
    for (int k = 0; k < 4; ++k) iv[l][k] = 0.f;
  if (nvalid > 0) {
    hp_f = load4_u8(ws.hold + ((size_t)(h.nm + h.nh) * B + b) * HW + own_off, nvalid, st_al);
    hp_m = load4_u8(ws.hold + ((size_t)(h.nm + h.nh + 1) * B + b) * HW + own_off, nvalid, st_al);
    for (int l = 0; l < 3; ++l)
      for (int k = 0; k < nvalid; ++k) iv[l][k] = ws.inv[((size_t)l * B + b) * HW + own_off + k];
  }
  __syncthreads();
  unsigned int pres[3] = {0u, 0u, 0u};
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    if (nonuni[l]) {
      for (int rr = 0; rr < 5; ++rr) {
        const unsigned char* row = labt + ((l * (TH + 4)) + ty + rr) * kLabPitch + tx;
        for (int q = 0; q < 8; ++q) pres[l] |= 1u << (row[q] & 31);
      }
    }
  }

  const float nv = fmaxf((float)ws.counts[0], 1.0f);
  const float wL[3] = {2.5f * loss_weight * gscale / (nv * (float)h.nf), 2.5f * loss_weight * gscale / (nv * (float)h.nm),
                       2.5f * loss_weight * gscale / (nv * (float)h.nh)};
  const float wCE = loss_weight * gscale / ((float)B * (float)HW);

  const T* xb = x + (long)b * C * HW;
  T* gb = grad + (long)b * C * HW;
  constexpr int nhalo = (TH + 4) * kPitch - TH * kTW;

  float xv[4];
  if (row_ok) load_n<T, 4>(xb, own_off, (long)y * W + W, vec_ok != 0, xv);
  else { xv[0] = xv[1] = xv[2] = xv[3] = 0.f; }

  for (int c = 0; c < C; ++c) {
    const int lvl = c < h.nf ? 0 : (c < h.nf + h.nm ? 1 : 2);
    const int cl = lvl == 0 ? c : (lvl == 1 ? c - h.nf : c - h.nf - h.nm);
    const T* xc = xb + (long)c * HW;
    float s[4], v[4], pk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      SigExp se = sig_exp(xv[k]);
      s[k] = se.s; v[k] = se.v;
      pk[k] = inimg[k] ? ((tf[k] != SH_IGNORE ? se.s : 0.f) + 1e-6f) : 0.f;
    }
    *reinterpret_cast<float2*>(plane + (ty + 2) * kPitch + tx + 2) = make_float2(pk[0], pk[1]);
    *reinterpret_cast<float2*>(plane + (ty + 2) * kPitch + tx + 4) = make_float2(pk[2], pk[3]);
    for (int e = tid; e < nhalo; e += 256) {
      int r, j;
      if (e < 4 * kPitch) { const int rr = e / kPitch; r = rr < 2 ? rr : TH + rr; j = e % kPitch; }
      else { const int e2 = e - 4 * kPitch; r = 2 + (e2 >> 2); const int q = e2 & 3; j = q < 2 ? q : kTW + q; }
      const int yy = y0 - 2 + r, xx = x0 - 2 + j;
      float p = 0.f;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const bool valid = lab8[(long)yy * W + xx] != SH_IGNORE;
        p = (valid ? sig_exp(to_f32<T>(xc[(long)yy * W + xx])).s : 0.f) + 1e-6f;
      }
      plane[r * kPitch + j] = p;
    }
    if (tid < 64) wbuf[tid] = ws.wts[((size_t)b * C + c) * 64 + tid];
    if (c + 1 < C && row_ok) load_n<T, 4>(xc + HW, own_off, (long)y * W + W, vec_ok != 0, xv);
    __syncthreads();

    // ---- RMI gradient wrt P: 5x5 stencil -------------------------------------------------------
    float dP[4] = {0.f, 0.f, 0.f, 0.f};
    {
      float w1[25];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const float4 t4 = *reinterpret_cast<const float4*>(wbuf + 4 * q);
        w1[4 * q] = t4.x; w1[4 * q + 1] = t4.y; w1[4 * q + 2] = t4.z; w1[4 * q + 3] = t4.w;
      }
      w1[24] = wbuf[24];
#pragma unroll
      for (int rr = 0; rr < 5; ++rr) {
        const float4* pr = reinterpret_cast<const float4*>(plane + (ty + rr) * kPitch + tx);
        const float4 a = pr[0], bq = pr[1];
        const float w[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int dx = 0; dx < 5; ++dx)
#pragma unroll
          for (int k = 0; k < 4; ++k) dP[k] = fmaf(w1[rr * 5 + dx], w[k + dx], dP[k]);
      }
      const float w2full = wbuf[50];
      const unsigned int ul = ulab[lvl];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (((ul >> (8 * k)) & 0xffu) == (unsigned int)cl) dP[k] += w2full;
      if (nonuni[lvl] && ((pres[lvl] >> (cl & 31)) & 1u)) {
        float add[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int rr = 0; rr < 5; ++rr) {
          const unsigned int* row = reinterpret_cast<const unsigned int*>(
              labt + ((lvl * (TH + 4)) + ty + rr) * kLabPitch + tx);
          const unsigned long long wbits = (unsigned long long)row[0] | ((unsigned long long)row[1] << 32);
          float mt[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) mt[q] = (byte_of(wbits, q) == (unsigned int)cl) ? 1.f : 0.f;
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            const float w2 = wbuf[25 + rr * 5 + dx];
#pragma unroll
            for (int k = 0; k < 4; ++k) add[k] = fmaf(w2, mt[k + dx], add[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((nonuni[lvl] >> k) & 1u) dP[k] += add[k];
      }
    }

    // ---- tree BCE + CE gradient from the summaries ----------------------------------------------
    float g[4];
    {
      unsigned int hm = 0;       // holder of MCMBc for this channel's mid
      int mid = -1;
      if (lvl == 0) mid = h.f2m[cl]; else if (lvl == 1) mid = cl;
      if (mid >= 0 && nvalid > 0) hm = load4_u8(ws.hold + ((size_t)mid * B + b) * HW + own_off, nvalid, st_al);
      float dneg[4] = {0.f, 0.f, 0.f, 0.f}, dpos[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (tf[k] == SH_IGNORE) continue;
        const unsigned int hpf = (hp_f >> (8 * k)) & 0xffu, hpm = (hp_m >> (8 * k)) & 0xffu;
        if (lvl == 0) {
          if (cl == tf[k]) { if (hpf == (unsigned int)c) dpos[k] += wL[0]; }
          else dneg[k] += wL[0];
          if (mid != tm[k] && ((hm >> (8 * k)) & 0xffu) == (unsigned int)c) dneg[k] += wL[1];
        } else if (lvl == 1) {
          if (cl == tm[k]) {
            if (hpf == (unsigned int)c) dpos[k] += wL[0];
            if (hpm == (unsigned int)c) dpos[k] += wL[1];
          } else if (((hm >> (8 * k)) & 0xffu) == (unsigned int)c) dneg[k] += wL[1];
        } else {
          if (hpm == (unsigned int)c) dpos[k] += wL[1];
          if (cl == thh[k]) dpos[k] += wL[2];
        }
      }
      // high-level negative terms routed to this channel
      if (lvl == 2) {
        if (nvalid > 0) {
          const unsigned int hh = load4_u8(ws.hold + ((size_t)(h.nm + cl) * B + b) * HW + own_off, nvalid, st_al);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (tf[k] != SH_IGNORE && cl != thh[k] && ((hh >> (8 * k)) & 0xffu) == (unsigned int)c) dneg[k] += wL[2];
        }
      } else if (nvalid > 0) {
        for (int q = h.mh_ptr[mid]; q < h.mh_ptr[mid + 1]; ++q) {
          const int hi = h.mh_idx[q];
          const unsigned int hh = load4_u8(ws.hold + ((size_t)(h.nm + hi) * B + b) * HW + own_off, nvalid, st_al);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (tf[k] != SH_IGNORE && hi != thh[k] && ((hh >> (8 * k)) & 0xffu) == (unsigned int)c) dneg[k] += wL[2];
        }
      }
      const int tgt_sel = lvl;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float q1 = 1.0f - s[k];
        float ds = 0.f;
        if (dneg[k] != 0.f) ds += dneg[k] * rcp(q1 + eps);
        if (dpos[k] != 0.f) ds -= dpos[k] * rcp(s[k] + eps);
        const bool valid = tf[k] != SH_IGNORE;
        if (valid && interior[k]) ds += dP[k] * gscale;
        float ce = 0.f;
        if (valid) {
          const int tgt = tgt_sel == 0 ? tf[k] : (tgt_sel == 1 ? tm[k] : thh[k]);
          ce = wCE * (v[k] * iv[lvl][k] - (cl == tgt ? 1.f : 0.f));
        }
        g[k] = ds * (q1 * s[k]) + ce;
      }
    }
    if (row_ok) store_n<T, 4>(gb + (long)c * HW, own_off, (long)y * W + W, vec_ok != 0, g);
    __syncthreads();
  }
}

// grid (B*C, nseg), block 256: threads stride over the frame pixels of one (b, c) plane
template <typename T>
__global__ void __launch_bounds__(256) k3_frame2(const T* __restrict__ x, T* __restrict__ grad, int B, int H, int W,
                                                 Hier3 h, Ws3 ws, const float* __restrict__ gscale_ptr) {
  __shared__ float fw[25 * 50];
  const int C = h.nf + h.nm + h.nh;
  const int bc = blockIdx.x, b = bc / C, c = bc % C;
  const int lvl = c < h.nf ? 0 : (c < h.nf + h.nm ? 1 : 2);
  const int cl = lvl == 0 ? c : (lvl == 1 ? c - h.nf : c - h.nf - h.nm);
  const int* lmap = lvl == 0 ? nullptr : (lvl == 1 ? h.f2m : h.f2h);
  const long HW = (long)H * W;
  const T* xc = x + ((long)b * C + c) * HW;
  T* gc = grad + ((long)b * C + c) * HW;
  const unsigned char* lab8 = ws.lab8 + (long)b * HW;
  const float gscale = *gscale_ptr;
  for (int i = threadIdx.x; i < 25 * 50; i += 256) fw[i] = ws.fwts[(size_t)bc * 25 * 50 + i];
  __syncthreads();
  const int nframe = 4 * W + 4 * (H - 4);
  const int per = (nframe + gridDim.y - 1) / gridDim.y;
  const int lo = blockIdx.y * per, hi = min(nframe, lo + per);
  for (int idx = lo + threadIdx.x; idx < hi; idx += 256) {
    int yy, xx;
    if (idx < 4 * W) { const int r = idx / W; yy = r < 2 ? r : H - 4 + r; xx = idx - r * W; }
    else { const int i2 = idx - 4 * W; const int q = i2 / (H - 4); xx = q < 2 ? q : W - 4 + q; yy = 2 + (i2 - q * (H - 4)); }
    const int tr = lab8[(long)yy * W + xx];
    if (tr == SH_IGNORE) continue;   // dP/ds = valid
    const float* w = fw + (axis_class(yy, H) * 5 + axis_class(xx, W)) * 50;
    float dP = 0.f;
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        const int y2 = yy + dy, x2 = xx + dx;
        if (y2 < 0 || y2 >= H || x2 < 0 || x2 >= W) continue;
        const int tn = lab8[(long)y2 * W + x2];
        const bool vn = tn != SH_IGNORE;
        const float pn = (vn ? sig_exp(to_f32<T>(xc[(long)y2 * W + x2])).s : 0.f) + 1e-6f;
        const int rl_n = vn ? (lmap ? lmap[tn] : tn) : 0;
        const int t = (dy + 2) * 5 + dx + 2;
        dP = fmaf(w[t], pn, dP);
        if (rl_n == cl) dP += w[25 + t];
      }
    }
    const float s = sig_exp(to_f32<T>(xc[(long)yy * W + xx])).s;
    const float add = dP * gscale * s * (1.0f - s);
    gc[(long)yy * W + xx] = from_f32<T>(to_f32<T>(gc[(long)yy * W + xx]) + add);
  }
}

template <typename T>
static int run_backward3(const void* x, void* grad, int B, int H, int W, const Hier3& h, const Ws3& ws, float eps,
                         float lw, const float* gscale, int stages, cudaStream_t st) {
  const int C = h.nf + h.nm + h.nh;
  const bool vec_ok = ((W & 3) == 0) && ((uintptr_t)x % (4 * sizeof(T)) == 0) && ((uintptr_t)grad % (4 * sizeof(T)) == 0);
  dim3 g2((W + kTW - 1) / kTW, (H + 15) / 16, B);
  if (stages & 1) {
    k3_pass2<T><<<g2, 256, 0, st>>>((const T*)x, (T*)grad, B, H, W, h, ws, eps, lw, gscale, vec_ok ? 1 : 0);
    SH_CHECK_LAUNCH();
  }
  if (stages & 2) {
    k3_frame2<T><<<dim3(B * C, ws.nseg), 256, 0, st>>>((const T*)x, (T*)grad, B, H, W, h, ws, gscale);
    SH_CHECK_LAUNCH();
  }
  return SH_OK;
}

}  // namespace sh

extern "C" {

int sh_rmi3_backward(const void* logits, int dtype, void* grad, int B, int H, int W, int nf, int nm, int nh,
                     const int* hier_tab, int n_mh, float loss_weight, void* workspace, const float* grad_out,
                     int stages, void* stream) {
  if (B <= 0 || H < 5 || W < 5 || nf + nm + nh > 255) return SH_ERR_BAD_ARG;
  sh::Ws3 ws = sh::ws3_layout(workspace, B, H, W, nf, nm, nh);
  sh::Hier3 h = sh::hier3_from_tab(hier_tab, nf, nm, nh, n_mh);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case SH_DT_F32: return sh::run_backward3<float>(logits, grad, B, H, W, h, ws, 1e-6f, loss_weight, grad_out, stages, st);
    case SH_DT_BF16: return sh::run_backward3<__nv_bfloat16>(logits, grad, B, H, W, h, ws, 1e-6f, loss_weight, grad_out, stages, st);
    case SH_DT_F16: return sh::run_backward3<__half>(logits, grad, B, H, W, h, ws, 1e-6f, loss_weight, grad_out, stages, st);
  }
  return SH_ERR_UNSUPPORTED;
}

}  // extern "C"
