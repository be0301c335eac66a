// Three-level hierarchical loss, forward side:
//   k3_prep     labels int64 -> uint8 + per-pixel flags (interior / label-uniform 5x5 per level)
//   k3_pass1    ONE streaming read of the logits: tree BCE + CE sums, per-pixel
//               summaries for the backward pass, RMI interior taps per (tile, channel)
//   k3_frame1   RMI taps of the 2-pixel image frame, per border class
//   k3_finalize per (b,c): fp64 reduction, 9x9 algebra (inverse, Schur, log-det),
//               analytic adjoints -> 5x5 stencil weights for the backward pass
//   k3_loss     scalar loss assembly (device side, no host sync)
// Reference arithmetic: models/loss/rmi_hiera_triplet_loss.py:323-546.
#include "rmi3_common.cuh"

namespace sh {

// ---------------------------------------------------------------------------------------------
// k3_prep
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_prep(const long long* __restrict__ label, int B, int H, int W, Hier3 h,
                                               unsigned char* __restrict__ lab8, unsigned char* __restrict__ flags,
                                               unsigned long long* __restrict__ counts) {
  constexpr int TH = 16;
  __shared__ __align__(8) unsigned char rl[3][TH + 4][kLabPitch];
  const int b = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * kTW;
  const long long* lb = label + (long)b * H * W;
  bool bad = false;
  for (int e = threadIdx.x; e < (TH + 4) * kPitch; e += 256) {
    const int r = e / kPitch, j = e - r * kPitch;
    const int y = y0 - 2 + r, x = x0 - 2 + j;
    unsigned char f = 0xff, m = 0xff, g = 0xff;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const long long t = lb[(long)y * W + x];
      f = m = g = 0;  // void pixels are one-hot of class 0 at every level inside RMI
      if (t != SH_IGNORE) {
        if (t >= 0 && t < h.nf) { f = (unsigned char)t; m = (unsigned char)h.f2m[t]; g = (unsigned char)h.f2h[t]; }
        else bad = true;
      }
    }
    rl[0][r][j] = f; rl[1][r][j] = m; rl[2][r][j] = g;
  }
  __syncthreads();
  const int ty = threadIdx.x / kStrips, tx = (threadIdx.x % kStrips) * 4;
  const int y = y0 + ty;
  long long nv = 0;
  if (y < H) {
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + tx + k;
      if (x >= W) break;
      const long long t = lb[(long)y * W + x];
      const bool valid = (t != SH_IGNORE);
      nv += valid;
      unsigned char fl = 0;
      if (y >= 2 && y < H - 2 && x >= 2 && x < W - 2) {
        fl = kFlagInterior;
#pragma unroll
        for (int lvl = 0; lvl < 3; ++lvl) {
          const unsigned char ctr = rl[lvl][ty + 2][tx + k + 2];
          bool uni = true;
#pragma unroll
          for (int dy = 0; dy < 5; ++dy)
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) uni &= (rl[lvl][ty + dy][tx + k + dx] == ctr);
          if (uni) fl |= (kFlagUniF << lvl);
        }
      }
      lab8[(long)b * H * W + (long)y * W + x] = (valid && t >= 0 && t < h.nf) ? (unsigned char)t : SH_IGNORE;
      flags[(long)b * H * W + (long)y * W + x] = fl;
    }
  }
  nv = warp_sum(nv);
  if ((threadIdx.x & 31) == 0 && nv) atomicAdd(counts, (unsigned long long)nv);
  if (bad) atomicOr((unsigned int*)(counts + 2), 1u);
}

// 16 values per lane -> every even lane holds the warp total of element
// e(lane) = 8*bit4 + 4*bit3 + 2*bit2 + bit1 in the return value (61 instructions instead of 160).
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int n = 8, off = 16; n >= 1; n >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? v[i] : v[i + n];
      const float keep = up ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int reduce16_slot(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

__device__ __forceinline__ void store4_u8(unsigned char* p, const unsigned char (&v)[4], int nvalid, bool aligned) {
  if (aligned && nvalid == 4) {
    *reinterpret_cast<unsigned int*>(p) = v[0] | (v[1] << 8) | (v[2] << 16) | ((unsigned int)v[3] << 24);
  } else {
    for (int k = 0; k < nvalid; ++k) p[k] = v[k];
  }
}

// ---------------------------------------------------------------------------------------------
// k3_pass1: grid (tiles_x, tiles_y, B), block th*16 threads, one thread = 4 consecutive pixels.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256, 2)
k3_pass1(const T* __restrict__ x, int B, int H, int W, Hier3 h, Ws3 ws, float eps, int vec_ok) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int th = blockDim.x / kStrips;
  const int PX = th * kTW;
  const int C = h.nf + h.nm + h.nh;
  float* plane = reinterpret_cast<float*>(smem_raw);                 // [(th+2)][kPitch]
  float* maxA = plane + (th + 2) * kPitch;                           // [nm][PX]
  float* maxB = maxA + (size_t)h.nm * PX;                            // [nh][PX]
  float* chacc = maxB + (size_t)h.nh * PX;                           // [64]
  unsigned char* holdA = reinterpret_cast<unsigned char*>(chacc + 64);  // [nm][PX]
  unsigned char* holdB = holdA + (size_t)h.nm * PX;                  // [nh][PX]
  unsigned char* labt = holdB + (size_t)h.nh * PX;                   // [3][(th+4)][kLabPitch]

  const int b = blockIdx.z, y0 = blockIdx.y * th, x0 = blockIdx.x * kTW;
  const long HW = (long)H * W;
  const int tid = threadIdx.x, lane = tid & 31;
  const int ty = tid / kStrips, tx = (tid % kStrips) * 4;
  const int y = y0 + ty, xg = x0 + tx;
  const long tile_id = ((long)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const unsigned char* lab8 = ws.lab8 + (long)b * HW;
  const unsigned char* flg = ws.flags + (long)b * HW;

  // ---- label tile (3 levels, halo 2) ------------------------------------------------------------
  for (int e = tid; e < (th + 4) * kPitch; e += blockDim.x) {
    const int r = e / kPitch, j = e - r * kPitch;
    const int yy = y0 - 2 + r, xx = x0 - 2 + j;
    unsigned char f = 0xff, m = 0xff, g = 0xff;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const int t = lab8[(long)yy * W + xx];
      f = m = g = 0;
      if (t != SH_IGNORE) { f = (unsigned char)t; m = (unsigned char)h.f2m[t]; g = (unsigned char)h.f2h[t]; }
    }
    labt[(0 * (th + 4) + r) * kLabPitch + j] = f;
    labt[(1 * (th + 4) + r) * kLabPitch + j] = m;
    labt[(2 * (th + 4) + r) * kLabPitch + j] = g;
  }
  // ---- per-pixel label state ----------------------------------------------------------------------
  int tf[4], tm[4], thh[4];
  unsigned int hsm[4];
  bool inimg[4], interior[4];
  unsigned int ulab[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};  // byte k: class if (interior & uniform) else 0xff
  unsigned int nonuni[3] = {0u, 0u, 0u};                           // bit k: interior & !uniform
  unsigned int rlab[3] = {0u, 0u, 0u};                             // byte k: RMI label (void -> 0)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    inimg[k] = (y < H) && (xg + k < W);
    int t = SH_IGNORE, fl = 0;
    if (inimg[k]) { t = lab8[(long)y * W + xg + k]; fl = flg[(long)y * W + xg + k]; }
    tf[k] = t;
    tm[k] = t != SH_IGNORE ? h.f2m[t] : SH_IGNORE;
    thh[k] = t != SH_IGNORE ? h.f2h[t] : SH_IGNORE;
    hsm[k] = t != SH_IGNORE ? h.hsmask[tm[k]] : 0u;
    interior[k] = (fl & kFlagInterior) != 0;
    const int r3[3] = {t != SH_IGNORE ? t : 0, t != SH_IGNORE ? tm[k] : 0, t != SH_IGNORE ? thh[k] : 0};
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      rlab[l] |= (unsigned int)r3[l] << (8 * k);
      if (interior[k]) {
        if (fl & (kFlagUniF << l)) ulab[l] = (ulab[l] & ~(0xffu << (8 * k))) | ((unsigned int)r3[l] << (8 * k));
        else nonuni[l] |= 1u << k;
      }
    }
  }
  for (int i = tid; i < h.nm * PX; i += blockDim.x) { maxA[i] = -1.f; holdA[i] = 0; }
  for (int i = tid; i < h.nh * PX; i += blockDim.x) { maxB[i] = -1.f; holdB[i] = 0; }
  __syncthreads();
  // bloom filter of the labels in this strip's 5x8 window (only if some pixel needs the slow path)
  unsigned int pres[3] = {0u, 0u, 0u};
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    if (nonuni[l]) {
      for (int rr = 0; rr < 5; ++rr) {
        const unsigned char* row = labt + ((l * (th + 4)) + ty + rr) * kLabPitch + tx;
        for (int q = 0; q < 8; ++q) pres[l] |= 1u << (row[q] & 31);
      }
    }
  }

  // ---- streaming state --------------------------------------------------------------------------
  float sumv[3][4], vt[3][4], a_t[4], b_t[4], c_t[4], min_c[4], prod[4];
  int hold_minc[4];
  float lacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // BCE fine/mid/high, CE fine/mid/high
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int l = 0; l < 3; ++l) { sumv[l][k] = 0.f; vt[l][k] = 1.f; }
    a_t[k] = b_t[k] = c_t[k] = 1.f; min_c[k] = 3.0e38f; hold_minc[k] = 0; prod[k] = 1.f;
  }
  const int px0 = ty * kTW + tx;
  const T* xb = x + (long)b * C * HW;
  const long own_off = (long)y * W + xg;
  const bool row_ok = y < H;
  const int nhalo = (th + 2) * kPitch - th * kTW;
  const bool st_al = vec_ok && ((W & 3) == 0);

  float xv[4];
  if (row_ok) load_n<T, 4>(xb, own_off, (long)y * W + W, vec_ok != 0, xv);
  else { xv[0] = xv[1] = xv[2] = xv[3] = 0.f; }

  for (int c = 0; c < C; ++c) {
    const int lvl = c < h.nf ? 0 : (c < h.nf + h.nm ? 1 : 2);
    const int cl = lvl == 0 ? c : (lvl == 1 ? c - h.nf : c - h.nf - h.nm);
    const T* xc = xb + (long)c * HW;
    // ---------------- phase A: sigmoid/exp, BCE/CE streaming, park P in the plane ---------------
    float pk[4];
    {
      float s[4], v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        SigExp se = sig_exp(xv[k]);
        s[k] = se.s; v[k] = se.v;
        pk[k] = inimg[k] ? ((tf[k] != SH_IGNORE ? se.s : 0.f) + 1e-6f) : 0.f;
      }
      *reinterpret_cast<float2*>(plane + ty * kPitch + tx + 2) = make_float2(pk[0], pk[1]);
      *reinterpret_cast<float2*>(plane + ty * kPitch + tx + 4) = make_float2(pk[2], pk[3]);
      if (lvl == 0) {
        const int m = h.f2m[cl];
        float4 cur = *reinterpret_cast<float4*>(maxA + (size_t)m * PX + px0);
        unsigned int hd = *reinterpret_cast<unsigned int*>(holdA + (size_t)m * PX + px0);
        float curv[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sumv[0][k] += v[k];
          if (cl == tf[k]) { a_t[k] = s[k]; vt[0][k] = v[k]; }
          else prod[k] *= (1.0f - s[k]) + eps;
          if (s[k] > curv[k]) { curv[k] = s[k]; hd = (hd & ~(0xffu << (8 * k))) | ((unsigned int)c << (8 * k)); }
        }
        *reinterpret_cast<float4*>(maxA + (size_t)m * PX + px0) = make_float4(curv[0], curv[1], curv[2], curv[3]);
        *reinterpret_cast<unsigned int*>(holdA + (size_t)m * PX + px0) = hd;
        if ((cl & 3) == 3 || cl == h.nf - 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { if (tf[k] != SH_IGNORE) lacc[0] -= fast_log(prod[k]); prod[k] = 1.f; }
        }
      } else if (lvl == 1) {
        float4 cur = *reinterpret_cast<float4*>(maxA + (size_t)cl * PX + px0);
        unsigned int hd = *reinterpret_cast<unsigned int*>(holdA + (size_t)cl * PX + px0);
        float curv[4] = {cur.x, cur.y, cur.z, cur.w};
        unsigned char hv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sumv[1][k] += v[k];
          int holder = (hd >> (8 * k)) & 0xff;
          if (s[k] > curv[k]) { curv[k] = s[k]; holder = c; }   // fine max wins ties
          hv[k] = (unsigned char)holder;
          if (cl == tm[k]) { b_t[k] = s[k]; vt[1][k] = v[k]; }
          else prod[k] *= (1.0f - curv[k]) + eps;
        }
        for (int q = h.mh_ptr[cl]; q < h.mh_ptr[cl + 1]; ++q) {
          const int hh = h.mh_idx[q];
          float4 cb = *reinterpret_cast<float4*>(maxB + (size_t)hh * PX + px0);
          unsigned int hb = *reinterpret_cast<unsigned int*>(holdB + (size_t)hh * PX + px0);
          float cbv[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (curv[k] > cbv[k]) { cbv[k] = curv[k]; hb = (hb & ~(0xffu << (8 * k))) | ((unsigned int)hv[k] << (8 * k)); }
          *reinterpret_cast<float4*>(maxB + (size_t)hh * PX + px0) = make_float4(cbv[0], cbv[1], cbv[2], cbv[3]);
          *reinterpret_cast<unsigned int*>(holdB + (size_t)hh * PX + px0) = hb;
        }
        if (row_ok) {
          int nvalid = W - xg; nvalid = nvalid > 4 ? 4 : nvalid;
          if (nvalid > 0) store4_u8(ws.hold + ((size_t)cl * B + b) * HW + own_off, hv, nvalid, st_al);
        }
        if ((cl & 3) == 3 || cl == h.nm - 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { if (tf[k] != SH_IGNORE) lacc[1] -= fast_log(prod[k]); prod[k] = 1.f; }
        }
      } else {
        float4 cur = *reinterpret_cast<float4*>(maxB + (size_t)cl * PX + px0);
        unsigned int hd = *reinterpret_cast<unsigned int*>(holdB + (size_t)cl * PX + px0);
        float curv[4] = {cur.x, cur.y, cur.z, cur.w};
        unsigned char hv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sumv[2][k] += v[k];
          int holder = (hd >> (8 * k)) & 0xff;
          if (s[k] > curv[k]) { curv[k] = s[k]; holder = c; }   // mid max wins ties
          hv[k] = (unsigned char)holder;
          if (cl == thh[k]) { c_t[k] = s[k]; vt[2][k] = v[k]; }
          else prod[k] *= (1.0f - curv[k]) + eps;
          if ((hsm[k] >> cl) & 1u) { if (s[k] < min_c[k]) { min_c[k] = s[k]; hold_minc[k] = c; } }
        }
        if (row_ok) {
          int nvalid = W - xg; nvalid = nvalid > 4 ? 4 : nvalid;
          if (nvalid > 0) store4_u8(ws.hold + ((size_t)(h.nm + cl) * B + b) * HW + own_off, hv, nvalid, st_al);
        }
        if ((cl & 3) == 3 || cl == h.nh - 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { if (tf[k] != SH_IGNORE) lacc[2] -= fast_log(prod[k]); prod[k] = 1.f; }
        }
      }
    }
    // halo of the plane: 2 rows below, 2 columns either side
    for (int e = tid; e < nhalo; e += blockDim.x) {
      int r, j;
      if (e < 2 * kPitch) { r = th + e / kPitch; j = e % kPitch; }
      else { const int e2 = e - 2 * kPitch; r = e2 >> 2; const int q = e2 & 3; j = q < 2 ? q : kTW + q; }
      const int yy = y0 + r, xx = x0 - 2 + j;
      float p = 0.f;
      if (yy < H && xx >= 0 && xx < W) {
        const bool valid = lab8[(long)yy * W + xx] != SH_IGNORE;
        p = (valid ? sig_exp(to_f32<T>(xc[(long)yy * W + xx])).s : 0.f) + 1e-6f;
      }
      plane[r * kPitch + j] = p;
    }
    if (tid < kNPart) chacc[tid] = 0.f;
    // prefetch next channel's own pixels
    if (c + 1 < C && row_ok) load_n<T, 4>(xc + HW, own_off, (long)y * W + W, vec_ok != 0, xv);
    __syncthreads();

    // ---------------- phase B: interior taps -----------------------------------------------------
    {
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      float w0[8], w1[8], w2[8];
      {
        const float4* p0 = reinterpret_cast<const float4*>(plane + (ty + 0) * kPitch + tx);
        const float4* p1 = reinterpret_cast<const float4*>(plane + (ty + 1) * kPitch + tx);
        const float4* p2 = reinterpret_cast<const float4*>(plane + (ty + 2) * kPitch + tx);
        float4 a = p0[0], bq = p0[1]; w0[0]=a.x; w0[1]=a.y; w0[2]=a.z; w0[3]=a.w; w0[4]=bq.x; w0[5]=bq.y; w0[6]=bq.z; w0[7]=bq.w;
        a = p1[0]; bq = p1[1]; w1[0]=a.x; w1[1]=a.y; w1[2]=a.z; w1[3]=a.w; w1[4]=bq.x; w1[5]=bq.y; w1[6]=bq.z; w1[7]=bq.w;
        a = p2[0]; bq = p2[1]; w2[0]=a.x; w2[1]=a.y; w2[2]=a.z; w2[3]=a.w; w2[4]=bq.x; w2[5]=bq.y; w2[6]=bq.z; w2[7]=bq.w;
      }
      const unsigned int ul = ulab[lvl];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = interior[k] ? pk[k] : 0.f;
#pragma unroll
        for (int dx = 0; dx <= 2; ++dx) acc[dx] = fmaf(a, w0[k + 2 + dx], acc[dx]);
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
          acc[3 + dx + 2] = fmaf(a, w1[k + 2 + dx], acc[3 + dx + 2]);
          acc[8 + dx + 2] = fmaf(a, w2[k + 2 + dx], acc[8 + dx + 2]);
        }
        const bool hit = ((ul >> (8 * k)) & 0xffu) == (unsigned int)cl;
        acc[kLPFull] += hit ? pk[k] : 0.f;
        acc[kLLFull] += hit ? 1.f : 0.f;
      }
      const float tot = warp_reduce16(acc, lane);
      if ((lane & 1) == 0) {
        const int slot = reduce16_slot(lane);
        if (slot < 15 && tot != 0.f) atomicAdd(&chacc[slot], tot);
      }
      // slow path: anchors whose 5x5 label neighbourhood is not uniform at this level
      const bool want = nonuni[lvl] && ((pres[lvl] >> (cl & 31)) & 1u);
      if (__any_sync(0xffffffffu, want)) {
        float sl[48];  // [0,25) lp taps, [25,38) ll half-plane taps, rest padding
#pragma unroll
        for (int i = 0; i < 48; ++i) sl[i] = 0.f;
        if (want) {
          float bk[4], lk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool nu = (nonuni[lvl] >> k) & 1u;
            bk[k] = nu ? pk[k] : 0.f;
            lk[k] = (nu && ((rlab[lvl] >> (8 * k)) & 0xffu) == (unsigned int)cl) ? 1.f : 0.f;
          }
#pragma unroll
          for (int rr = 0; rr < 5; ++rr) {
            const unsigned int* row = reinterpret_cast<const unsigned int*>(
                labt + ((lvl * (th + 4)) + ty + rr) * kLabPitch + tx);
            const unsigned long long wbits = (unsigned long long)row[0] | ((unsigned long long)row[1] << 32);
            float mt[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) mt[q] = (byte_of(wbits, q) == (unsigned int)cl) ? 1.f : 0.f;
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                sl[rr * 5 + dx + 2] = fmaf(bk[k], mt[k + 2 + dx], sl[rr * 5 + dx + 2]);
                if (rr > 2 || (rr == 2 && dx >= 0)) {
                  const int hi = rr == 2 ? dx : 3 + (rr - 3) * 5 + (dx + 2);
                  sl[25 + hi] = fmaf(lk[k], mt[k + 2 + dx], sl[25 + hi]);
                }
              }
            }
          }
        }
#pragma unroll
        for (int base = 0; base < 48; base += 16) {
          const float t2 = warp_reduce16(*reinterpret_cast<float(*)[16]>(sl + base), lane);
          if ((lane & 1) == 0) {
            const int idx = base + reduce16_slot(lane);
            if (idx < 38 && t2 != 0.f) atomicAdd(&chacc[kLPS + idx], t2);
          }
        }
      }
    }
    __syncthreads();
    if (tid < kNPart) ws.part1[((size_t)tile_id * C + c) * kNPart + tid] = chacc[tid];
  }

  // ---- per-pixel epilogue: positive terms, CE, summaries ----------------------------------------
  {
    unsigned char hp_f[4], hp_m[4];
    float iv[3][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      hp_f[k] = 0; hp_m[k] = 0;
#pragma unroll
      for (int l = 0; l < 3; ++l) iv[l][k] = rcp(sumv[l][k]);
      if (tf[k] != SH_IGNORE) {
        const bool a_holds = a_t[k] <= b_t[k];                 // fine wins ties (rmi...py:421-425)
        const float mcla = a_holds ? a_t[k] : b_t[k];
        hp_f[k] = (unsigned char)(a_holds ? tf[k] : h.nf + tm[k]);
        const bool c_holds = min_c[k] <= b_t[k];               // high wins ties (rmi...py:439-440)
        const float mclb = c_holds ? min_c[k] : b_t[k];
        hp_m[k] = (unsigned char)(c_holds ? hold_minc[k] : h.nf + tm[k]);
        lacc[0] -= fast_log(mcla + eps);
        lacc[1] -= fast_log(mclb + eps);
        lacc[2] -= fast_log(c_t[k] + eps);
#pragma unroll
        for (int l = 0; l < 3; ++l) lacc[3 + l] += fast_log(sumv[l][k]) - fast_log(vt[l][k]);
      }
    }
    if (row_ok) {
      int nvalid = W - xg; nvalid = nvalid > 4 ? 4 : nvalid;
      if (nvalid > 0) {
        store4_u8(ws.hold + ((size_t)(h.nm + h.nh) * B + b) * HW + own_off, hp_f, nvalid, st_al);
        store4_u8(ws.hold + ((size_t)(h.nm + h.nh + 1) * B + b) * HW + own_off, hp_m, nvalid, st_al);
        for (int l = 0; l < 3; ++l)
          for (int k = 0; k < nvalid; ++k) ws.inv[((size_t)l * B + b) * HW + own_off + k] = iv[l][k];
      }
    }
  }
  __syncthreads();
  float* red = plane;  // reuse (needs 6 * nwarps floats)
  const float r = block_sum_k<6>(lacc, red);
  if (tid < 6) ws.bcepart[(size_t)tile_id * 8 + tid] = r;
  else if (tid < 8) ws.bcepart[(size_t)tile_id * 8 + tid] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// k3_frame1: taps of the image frame.  grid (B*C, nseg), block 256 = 8 warps; warp w owns run w
// (rows 0,1,H-2,H-1 x middle columns; columns 0,1,W-2,W-1 x middle rows); warp 0 of segment 0
// then does the 16 corner pixels.  Output: frameT[seg][b*C+c][class][75] = pp[25] lp[25] ll[25].
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void frame_pixel_taps(const T* __restrict__ xc, const unsigned char* __restrict__ lab8,
                                                 const int* __restrict__ lmap, int cl, int H, int W, int y, int x,
                                                 float (&acc)[75]) {
  const int tr = lab8[(long)y * W + x];
  const bool vr = tr != SH_IGNORE;
  const float pr = (vr ? sig_exp(to_f32<T>(xc[(long)y * W + x])).s : 0.f) + 1e-6f;
  const int rl_r = vr ? (lmap ? lmap[tr] : tr) : 0;
  const float lr = rl_r == cl ? 1.f : 0.f;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const int yy = y + dy, xx = x + dx;
      float pn = 0.f, ln = 0.f;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const int tn = lab8[(long)yy * W + xx];
        const bool vn = tn != SH_IGNORE;
        pn = (vn ? sig_exp(to_f32<T>(xc[(long)yy * W + xx])).s : 0.f) + 1e-6f;
        const int rl_n = vn ? (lmap ? lmap[tn] : tn) : 0;
        ln = rl_n == cl ? 1.f : 0.f;
      }
      const int t = (dy + 2) * 5 + dx + 2;
      acc[t] = fmaf(pr, pn, acc[t]);
      acc[25 + t] = fmaf(pr, ln, acc[25 + t]);
      acc[50 + t] = fmaf(lr, ln, acc[50 + t]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k3_frame1(const T* __restrict__ x, int B, int H, int W, Hier3 h, Ws3 ws) {
  const int C = h.nf + h.nm + h.nh;
  const int bc = blockIdx.x, b = bc / C, c = bc % C;
  const int seg = blockIdx.y, nseg = gridDim.y;
  const int lvl = c < h.nf ? 0 : (c < h.nf + h.nm ? 1 : 2);
  const int cl = lvl == 0 ? c : (lvl == 1 ? c - h.nf : c - h.nf - h.nm);
  const int* lmap = lvl == 0 ? nullptr : (lvl == 1 ? h.f2m : h.f2h);
  const long HW = (long)H * W;
  const T* xc = x + ((long)b * C + c) * HW;
  const unsigned char* lab8 = ws.lab8 + (long)b * HW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* out = ws.frameT + ((size_t)seg * B * C + bc) * 25 * 75;

  float acc[75];
#pragma unroll
  for (int i = 0; i < 75; ++i) acc[i] = 0.f;
  const bool is_row = warp < 4;
  const int line = (warp & 3) < 2 ? (warp & 3) : (is_row ? H : W) - 4 + (warp & 3);  // 0,1,n-2,n-1
  const int len = (is_row ? W : H) - 4;
  const int per = (len + nseg - 1) / nseg;
  const int lo = 2 + seg * per, hi = min(2 + len, lo + per);
  for (int i = lo + lane; i < hi; i += 32) {
    const int yy = is_row ? line : i, xx = is_row ? i : line;
    frame_pixel_taps<T>(xc, lab8, lmap, cl, H, W, yy, xx, acc);
  }
  const int kline = (warp & 3) < 2 ? (warp & 3) : 1 + (warp & 3);  // class 0,1,3,4
  const int cls = is_row ? kline * 5 + 2 : 2 * 5 + kline;
#pragma unroll
  for (int i = 0; i < 75; ++i) {
    const float r = warp_sum(acc[i]);
    if (lane == 0) out[cls * 75 + i] = r;
  }
  if (warp == 0) {
    // corners: one lane per pixel, each its own class
#pragma unroll
    for (int i = 0; i < 75; ++i) acc[i] = 0.f;
    const int ky = (lane >> 2) & 3, kx = lane & 3;
    const int cy = ky < 2 ? ky : ky + 1, cx = kx < 2 ? kx : kx + 1;  // classes 0,1,3,4
    if (lane < 16 && seg == 0) {
      const int yy = ky < 2 ? ky : H - 4 + ky, xx = kx < 2 ? kx : W - 4 + kx;
      frame_pixel_taps<T>(xc, lab8, lmap, cl, H, W, yy, xx, acc);
    }
    if (lane < 16) {
      for (int i = 0; i < 75; ++i) out[(cy * 5 + cx) * 75 + i] = acc[i];
    }
    if (lane == 16) {
      for (int i = 0; i < 75; ++i) out[(2 * 5 + 2) * 75 + i] = 0.f;  // interior class unused
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k3_finalize: one CTA (64 threads) per (b, c).
// ---------------------------------------------------------------------------------------------
__device__ void gj_inverse9(double* A, double* Inv, double* piv, double* fac, int lane) {
  for (int e = lane; e < 81; e += 32) Inv[e] = (e / 9 == e % 9) ? 1.0 : 0.0;
  __syncwarp();
  for (int k = 0; k < 9; ++k) {
    const double p = A[k * 9 + k];
    const double ip = 1.0 / p;
    __syncwarp();
    if (lane == 0) piv[k] = p;
    if (lane < 9) { A[k * 9 + lane] *= ip; Inv[k * 9 + lane] *= ip; }
    __syncwarp();
    if (lane < 9) fac[lane] = A[lane * 9 + k];
    __syncwarp();
    for (int e = lane; e < 81; e += 32) {
      const int i = e / 9, l = e % 9;
      if (i != k) {
        A[e] -= fac[i] * A[k * 9 + l];
        Inv[e] -= fac[i] * Inv[k * 9 + l];
      }
    }
    __syncwarp();
  }
}

// Out[i][j] = sum_k X[i][k] * Y[k][j]   (optionally X transposed / Y transposed)
__device__ void mm9(const double* X, bool xt, const double* Y, bool yt, double* Out, int lane) {
  for (int e = lane; e < 81; e += 32) {
    const int i = e / 9, j = e % 9;
    double a = 0.0;
    for (int k = 0; k < 9; ++k) a += (xt ? X[k * 9 + i] : X[i * 9 + k]) * (yt ? Y[j * 9 + k] : Y[k * 9 + j]);
    Out[e] = a;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(64) k3_finalize(int B, int C, long tiles_per_img, int nseg, Ws3 ws, double scale) {
  __shared__ double part[kNPart];
  __shared__ double fr[25 * 75];
  __shared__ double Spp[81], Slp[81], Sll[81], K[81], T1[81], M[81], Wm[81], U[81], Gpp[81], tmp[81];
  __shared__ double piv[9], fac[9];
  const int bc = blockIdx.x, b = bc / C, c = bc % C;
  const int tid = threadIdx.x;
  if (tid < kNPart) {
    double a = 0.0;
    const float* p = ws.part1 + ((size_t)b * tiles_per_img * C + c) * kNPart + tid;
    for (long t = 0; t < tiles_per_img; ++t) a += (double)p[(size_t)t * C * kNPart];
    part[tid] = a;
  }
  for (int e = tid; e < 25 * 75; e += 64) {
    double a = 0.0;
    for (int s = 0; s < nseg; ++s) a += (double)ws.frameT[((size_t)s * B * C + bc) * 25 * 75 + e];
    fr[e] = a;
  }
  __syncthreads();
  // assemble
  for (int e = tid; e < 81; e += 64) {
    const int i = e / 9, j = e % 9;
    const int yi = i / 3, xi = i % 3, yj = j / 3, xj = j % 3;
    const int dy = yi - yj, dx = xi - xj;
    const int t = tap_index(dy, dx);
    double fpp = 0.0, flp = 0.0, fll = 0.0;
    for (int ky = 0; ky < 5; ++ky)
      for (int kx = 0; kx < 5; ++kx) {
        if (ky == 2 && kx == 2) continue;
        if (!offset_valid(ky, yj) || !offset_valid(kx, xj)) continue;
        const double* f = fr + (ky * 5 + kx) * 75;
        fpp += f[t]; flp += f[25 + t]; fll += f[50 + t];
      }
    Slp[e] = part[kLPFull] + part[kLPS + t] + flp;
    if (in_half_plane(dy, dx)) {
      const int ht = half_tap_index(dy, dx);
      Spp[e] = part[kPP + ht] + fpp;
      Sll[e] = part[kLLFull] + part[kLLS + ht] + fll;
    }
  }
  __syncthreads();
  for (int e = tid; e < 81; e += 64) {
    const int i = e / 9, j = e % 9;
    const int dy = i / 3 - j / 3, dx = i % 3 - j % 3;
    if (!in_half_plane(dy, dx)) { Spp[e] = Spp[j * 9 + i]; Sll[e] = Sll[j * 9 + i]; }
  }
  __syncthreads();
  if (tid < 32) {
    const int lane = tid;
    for (int e = lane; e < 81; e += 32) tmp[e] = Spp[e] + ((e / 9 == e % 9) ? 1e-3 : 0.0);
    __syncwarp();
    gj_inverse9(tmp, K, piv, fac, lane);           // K = (Spp + aI)^-1
    mm9(Slp, false, K, false, T1, lane);           // T1 = Slp K
    mm9(T1, false, Slp, true, tmp, lane);          // tmp = Slp K Slp^T
    for (int e = lane; e < 81; e += 32) M[e] = Sll[e] - tmp[e] + ((e / 9 == e % 9) ? 1e-3 : 0.0);
    __syncwarp();
    for (int e = lane; e < 81; e += 32) tmp[e] = M[e];
    __syncwarp();
    gj_inverse9(tmp, Wm, piv, fac, lane);          // W = M^-1, pivots = chol(M)_kk^2
    if (lane == 0) {
      double r = 0.0;
      for (int k = 0; k < 9; ++k) r += log(sqrt(piv[k]) + 1e-8);
      ws.rbc[bc] = r;                              // = 0.5 * log_det_by_cholesky(M)
    }
    mm9(Wm, false, T1, false, U, lane);            // U = W Slp K ;  G_lp = -U
    mm9(T1, true, U, false, Gpp, lane);            // G_pp = 0.5 * T1^T U
    // stencil weights
    float* wout = ws.wts + (size_t)bc * 64;
    float* fout = ws.fwts + (size_t)bc * 25 * 50;
    if (lane < 25) {
      double w1 = 0.0, w2 = 0.0;
      for (int e = 0; e < 81; ++e) {
        const int i = e / 9, j = e % 9;
        if (tap_index(i / 3 - j / 3, i % 3 - j % 3) == lane) { w1 += Gpp[e]; w2 -= U[e]; }
      }
      wout[lane] = (float)(scale * w1);            // 2 * (0.5 * T1^T U) * scale
      wout[25 + lane] = (float)(scale * w2);
      tmp[lane] = scale * w2;
    }
    __syncwarp();
    if (lane == 0) {
      double s = 0.0;
      for (int t = 0; t < 25; ++t) s += tmp[t];
      wout[50] = (float)s;
    }
    for (int q = lane; q < 25 * 25; q += 32) {
      const int cls = q / 25, t = q % 25, ky = cls / 5, kx = cls % 5;
      double w1 = 0.0, w2 = 0.0;
      for (int e = 0; e < 81; ++e) {
        const int i = e / 9, j = e % 9;
        if (tap_index(i / 3 - j / 3, i % 3 - j % 3) != t) continue;
        if (!offset_valid(ky, j / 3) || !offset_valid(kx, j % 3)) continue;
        w1 += Gpp[e]; w2 -= U[e];
      }
      fout[cls * 50 + t] = (float)(scale * w1);
      fout[cls * 50 + 25 + t] = (float)(scale * w2);
    }
  }
}

__device__ __forceinline__ double schedule_factor3(double step, double total) {
  if (step < total) return 0.25 * (1.0 + cos((step - total) / total * 3.141592653589793));
  return 0.5;
}

// loss = (lam*rmi + 0.5*5*(Lf/(Nv nf)+Lm/(Nv nm)+Lh/(Nv nh)) + sum CE/Npx + ready*factor*trip) * lw
__global__ void __launch_bounds__(256) k3_loss(int B, int C, int nf, int nm, int nh, double npx, Ws3 ws, double lam,
                                               const double* __restrict__ step, double total_steps,
                                               const float* __restrict__ trip, const int* __restrict__ ready,
                                               float lw, float* __restrict__ out) {
  __shared__ float cls[256];
  float rmi = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) {
    double a = 0.0;
    for (int b = 0; b < B; ++b) a += ws.rbc[(size_t)b * C + c];
    rmi += (float)(a / B) / 9.0f;                 // .mean(dim=0).float() / half_d  (rmi...py:515-516)
  }
  cls[threadIdx.x] = rmi;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int i = 0; i < 256; ++i) r += cls[i];
    const double nv = fmax((double)ws.counts[0], 1.0);
    const double* s = ws.sums;
    double loss = lam * (double)r + 2.5 * (s[0] / (nv * nf) + s[1] / (nv * nm) + s[2] / (nv * nh)) +
                  (s[3] + s[4] + s[5]) / npx;
    double tscale = 0.0;
    if (trip != nullptr && ready != nullptr && *ready > 0 && trip[1] > 0.f) {
      const double f = schedule_factor3(*step, total_steps);
      loss += f * (double)trip[0];
      tscale = f * lw;
    }
    loss *= lw;
    if (ws.counts[2] != 0) loss = __longlong_as_double(0x7ff8000000000000LL);
    out[0] = (float)loss;
    out[1] = (float)tscale;
    out[2] = r;
  }
}

__global__ void __launch_bounds__(256) k_reduce_partials3(const float* __restrict__ part, long n, int K,
                                                          double* __restrict__ out) {
  __shared__ double sm[256];
  for (int k = 0; k < K; ++k) {
    double a = 0.0;
    for (long i = threadIdx.x; i < n; i += 256) a += (double)part[(size_t)i * K + k];
    sm[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = sm[0];
    __syncthreads();
  }
}

template <typename T>
static int run_forward3(const void* x, const long long* label, int B, int H, int W, const Hier3& h, const Ws3& ws,
                        float eps, double scale, int stages, cudaStream_t st) {
  const int C = h.nf + h.nm + h.nh;
  if (stages & 1) {
    cudaError_t e = cudaMemsetAsync(ws.counts, 0, 4 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    dim3 gp((W + kTW - 1) / kTW, (H + 15) / 16, B);
    k3_prep<<<gp, 256, 0, st>>>(label, B, H, W, h, ws.lab8, ws.flags, ws.counts);
    SH_CHECK_LAUNCH();
  }
  const int th = ws.th, PX = th * kTW;
  size_t smem = (size_t)(th + 2) * kPitch * 4 + (size_t)(h.nm + h.nh) * PX * 5 + 64 * 4 + 3 * (size_t)(th + 4) * kLabPitch;
  smem = (smem + 15) & ~(size_t)15;
  if (smem > 227 * 1024) return SH_ERR_UNSUPPORTED;
  const bool vec_ok = ((W & 3) == 0) && ((uintptr_t)x % (4 * sizeof(T)) == 0);
  auto kern = k3_pass1<T>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 g1(ws.tiles_x, ws.tiles_y, B);
  if (stages & 2) {
    kern<<<g1, th * kStrips, smem, st>>>((const T*)x, B, H, W, h, ws, eps, vec_ok ? 1 : 0);
    SH_CHECK_LAUNCH();
  }
  const int nseg = ws.nseg;
  if (stages & 4) {
    k3_frame1<T><<<dim3(B * C, nseg), 256, 0, st>>>((const T*)x, B, H, W, h, ws);
    SH_CHECK_LAUNCH();
  }
  if (stages & 8) {
    const long ntiles = (long)ws.tiles_x * ws.tiles_y * B;
    k_reduce_partials3<<<1, 256, 0, st>>>(ws.bcepart, ntiles, 8, ws.sums);
    SH_CHECK_LAUNCH();
    k3_finalize<<<B * C, 64, 0, st>>>(B, C, (long)ws.tiles_x * ws.tiles_y, nseg, ws, scale);
    SH_CHECK_LAUNCH();
  }
  return SH_OK;
}

}  // namespace sh

extern "C" {

size_t sh_rmi3_workspace_bytes(int B, int H, int W, int nf, int nm, int nh) {
  return sh::ws3_layout(nullptr, B, H, W, nf, nm, nh).bytes;
}

// Byte offsets of the workspace fields (diagnostics / tests):
// out[0..10] = counts, sums, lab8, flags, hold, inv, part1, bcepart, frameT, rbc, wts; out[11] = fwts
int sh_rmi3_workspace_offsets(int B, int H, int W, int nf, int nm, int nh, size_t* out) {
  sh::Ws3 w = sh::ws3_layout(nullptr, B, H, W, nf, nm, nh);
  const void* f[12] = {w.counts, w.sums, w.lab8, w.flags, w.hold, w.inv, w.part1, w.bcepart, w.frameT, w.rbc, w.wts, w.fwts};
  for (int i = 0; i < 12; ++i) out[i] = (size_t)f[i];
  return SH_OK;
}

// hier_tab (device int32): [f2m nf][f2h nf][mh_ptr nm+1][mh_idx n_mh][hsmask nm]
int sh_rmi3_forward(const void* logits, int dtype, const long long* label, int B, int H, int W, int nf, int nm, int nh,
                    const int* hier_tab, int n_mh, float lam, float loss_weight, void* workspace, int stages,
                    void* stream) {
  if (B <= 0 || H < 5 || W < 5 || nf <= 0 || nm <= 0 || nh <= 0 || nf + nm + nh > 255 || nh > 32)
    return SH_ERR_BAD_ARG;
  sh::Ws3 ws = sh::ws3_layout(workspace, B, H, W, nf, nm, nh);
  sh::Hier3 h = sh::hier3_from_tab(hier_tab, nf, nm, nh, n_mh);
  const double scale = (double)lam * (double)loss_weight / (9.0 * B);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case SH_DT_F32: return sh::run_forward3<float>(logits, label, B, H, W, h, ws, 1e-6f, scale, stages, st);
    case SH_DT_BF16: return sh::run_forward3<__nv_bfloat16>(logits, label, B, H, W, h, ws, 1e-6f, scale, stages, st);
    case SH_DT_F16: return sh::run_forward3<__half>(logits, label, B, H, W, h, ws, 1e-6f, scale, stages, st);
  }
  return SH_ERR_UNSUPPORTED;
}

int sh_loss3_final(int B, int H, int W, int nf, int nm, int nh, void* workspace, float lam, const double* step,
                   double total_steps, const float* trip, const int* ready, float loss_weight, float* out,
                   void* stream) {
  sh::Ws3 ws = sh::ws3_layout(workspace, B, H, W, nf, nm, nh);
  sh::k3_loss<<<1, 256, 0, (cudaStream_t)stream>>>(B, nf + nm + nh, nf, nm, nh, (double)B * H * W, ws, (double)lam,
                                                   step, total_steps, trip, ready, loss_weight, out);
  SH_CHECK_LAUNCH();
  return SH_OK;
}

}  // extern "C"
