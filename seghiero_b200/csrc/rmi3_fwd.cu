// Three-level hierarchical loss, forward side:
//   k3_prep     labels int64 -> uint8 + per-pixel flags (interior / label-uniform 5x5 per level)
//   k3_pass1    ONE streaming read of the logits: tree BCE + CE sums, per-pixel
//               summaries for the backward pass, RMI interior taps per (tile, channel)
//   k3_band     P = sigmoid*valid + 1e-6 of the 4-pixel image border bands (feeds the frame kernels)
//   k3_frame1   RMI taps of the 2-pixel image frame, per border class
//   k3_finalize per (b,c): fp64 reduction, 9x9 algebra (inverse, Schur, log-det),
//               analytic adjoints -> 5x5 stencil weights for the backward pass
//   k3_loss     scalar loss assembly (device side, no host sync)
// Reference arithmetic: models/loss/rmi_hiera_triplet_loss.py:323-546.
#include "rmi3_common.cuh"
#include "rmi3_fast.cuh"

namespace sh {

// ---------------------------------------------------------------------------------------------
// k3_prep: tile 64x16, 256 threads, one thread = 4 consecutive pixels
// ---------------------------------------------------------------------------------------------
template <typename L>
__global__ void __launch_bounds__(256) k3_prep(const L* __restrict__ label, int B, int H, int W, Hier3 h,
                                               unsigned char* __restrict__ lab8, unsigned char* __restrict__ flags,
                                               unsigned long long* __restrict__ counts) {
  constexpr int TH = 16;
  __shared__ __align__(8) unsigned char rl[3][TH + 4][kLabPitch];
  const int b = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * kTW;
  const L* lb = label + (long)b * H * W;
  bool bad = false;
  for (int e = threadIdx.x; e < (TH + 4) * kPitch; e += 256) {
    const int r = e / kPitch, j = e - r * kPitch;
    const int y = y0 - 2 + r, x = x0 - 2 + j;
    unsigned char f = 0xff, m = 0xff, g = 0xff;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const long long t = lab_ld(lb, (long)y * W + x);
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
  const int y = y0 + ty, xg = x0 + tx;
  long long nv = 0;
  if (y < H && xg < W) {
    unsigned int fl4 = 0, lab4 = 0xffffffffu;
    unsigned int uni[3] = {0xfu, 0xfu, 0xfu};   // bit k: 5x5 neighbourhood of pixel k uniform at this level
#pragma unroll
    for (int lvl = 0; lvl < 3; ++lvl) {
      const unsigned int* c2 = reinterpret_cast<const unsigned int*>(&rl[lvl][ty + 2][tx]);
      const unsigned long long ctr = (unsigned long long)c2[0] | ((unsigned long long)c2[1] << 32);
#pragma unroll
      for (int rr = 0; rr < 5; ++rr) {
        const unsigned int* rp = reinterpret_cast<const unsigned int*>(&rl[lvl][ty + rr][tx]);
        const unsigned long long w = (unsigned long long)rp[0] | ((unsigned long long)rp[1] << 32);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned long long pat = (unsigned long long)byte_of(ctr, k + 2) * 0x0101010101ULL;
          if (((w >> (8 * k)) & 0xffffffffffULL) != pat) uni[lvl] &= ~(1u << k);
        }
      }
    }
    const int nvalid = min(4, W - xg);
    for (int k = 0; k < nvalid; ++k) {
      const int x = xg + k;
      const long long t = lab_ld(lb, (long)y * W + x);
      const bool valid = (t != SH_IGNORE);
      nv += valid;
      unsigned int fl = 0;
      if (y >= 2 && y < H - 2 && x >= 2 && x < W - 2) {
        fl = kFlagInterior;
#pragma unroll
        for (int lvl = 0; lvl < 3; ++lvl)
          if ((uni[lvl] >> k) & 1u) fl |= (kFlagUniF << lvl);
      }
      const unsigned int l8 = (valid && t >= 0 && t < h.nf) ? (unsigned int)t : SH_IGNORE;
      lab4 = (lab4 & ~(0xffu << (8 * k))) | (l8 << (8 * k));
      fl4 |= fl << (8 * k);
    }
    const bool al = (W & 3) == 0;
    const long off = (long)b * H * W + (long)y * W + xg;
    store4_u8(lab8 + off, lab4, nvalid, al);
    store4_u8(flags + off, fl4, nvalid, al);
  }
  nv = warp_sum(nv);
  if (counts != nullptr) {        // null: a fast forward pass that already counted runs this kernel for the flags only
    if ((threadIdx.x & 31) == 0 && nv) atomicAdd(counts, (unsigned long long)nv);
    if (bad) atomicOr((unsigned int*)(counts + 2), 1u);
  }
}

// Slow path of k3_pass1 (kept out of line so that its registers do not burden the streaming loop):
// lp / ll taps of the interior anchors of one 4x4 block whose 5x5 label neighbourhood is NOT uniform.
// All lanes of the warp must call it (warp-level reductions); lanes with want == false add zeros.
// (A tap-per-lane redistribution over the lanes that have work was measured slower: label-byte gathers
// with 38 different offsets serialise in the shared-memory banks.)
__device__ __noinline__ void pass1_slow_path(const float* plane, const unsigned char* labl, unsigned int u0,
                                             unsigned int u1, unsigned int u2, unsigned int u3, int br, int bs,
                                             int cl, bool want, int lane, float* sp) {
  float sl[48];   // [0,25) lp taps, [25,38) ll half-plane taps, rest padding
#pragma unroll
  for (int i = 0; i < 48; ++i) sl[i] = 0.f;
  if (want) {
    const unsigned int ub4[4] = {u0, u1, u2, u3};
    float bk[4][4], lk[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* prow = plane + (4 * br + i) * kPitch + 4 * bs + 2;
      const unsigned char* crow = labl + (4 * br + i + 2) * kLabPitch + 4 * bs + 2;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool nu = ((ub4[i] >> (8 * k)) & 0xffu) == 0xfeu;
        bk[i][k] = nu ? prow[k] : 0.f;
        lk[i][k] = (nu && crow[k] == cl) ? 1.f : 0.f;
      }
    }
#pragma unroll
    for (int R = 0; R < 8; ++R) {    // label rows 4br-2 .. 4br+5 of the tile
      const unsigned int* row = reinterpret_cast<const unsigned int*>(labl + (4 * br + R) * kLabPitch + 4 * bs);
      const unsigned long long wbits = (unsigned long long)row[0] | ((unsigned long long)row[1] << 32);
      float mt[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) mt[q] = (byte_of(wbits, q) == (unsigned int)cl) ? 1.f : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = R - i;        // dy + 2
        if (rr < 0 || rr > 4) continue;
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            sl[rr * 5 + dx] = fmaf(bk[i][k], mt[k + dx], sl[rr * 5 + dx]);
            if (rr > 2 || (rr == 2 && dx >= 2)) {
              const int hi = rr == 2 ? dx - 2 : 3 + (rr - 3) * 5 + dx;
              sl[25 + hi] = fmaf(lk[i][k], mt[k + dx], sl[25 + hi]);
            }
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
      if (idx < 38 && t2 != 0.f) atomicAdd(sp + idx, t2);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k3_pass1: grid (tiles_x, tiles_y, B), 512 threads, tile 64x32.
//   phase A (thread = 4 consecutive pixels, all channels): sigmoid/exp, tree BCE + CE streaming
//            (channels visited in tree order so running maxima stay in registers), P -> plane
//   phase B (thread = one 4x4 block of ONE of the round's 4 channel planes): interior taps
// Planes are double buffered: one __syncthreads per round of 4 channels.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads, 512 / kThreads)
k3_pass1(const T* __restrict__ x, int B, int H, int W, Hier3 hg, Ws3 ws, float eps, int vec_ok) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int PX = kTH * kTW;
  constexpr int kPlane = (kTH + 2) * kPitch;
  const int C = hg.nf + hg.nm + hg.nh;
  float* planes = reinterpret_cast<float*>(smem_raw);                    // [2][kNR][kPlane]
  float* red = planes + 2 * kNR * kPlane;                                // [2][kNR][kGroupWarps][20]
  float* slow = red + 2 * kNR * kGroupWarps * 20;                        // [2][kNR][40]
  float* maxB = slow + 2 * kNR * 40;                                     // [nh][PX]
  unsigned char* holdB = reinterpret_cast<unsigned char*>(maxB + (size_t)hg.nh * PX);  // [nh][PX]
  unsigned char* U = holdB + (size_t)hg.nh * PX;                         // [3][kTH][kTW]
  unsigned char* labt = U + 3 * PX;                                      // [3][kTH+4][kLabPitch]
  int* htab = reinterpret_cast<int*>(labt + 3 * (kTH + 4) * kLabPitch);  // hierarchy tables (6 KB)
  uint4* xstage = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(htab) + 6144);   // [kNR][kThreads] raw logits
  unsigned int* hstage = reinterpret_cast<unsigned int*>(xstage + kNR * kThreads);           // [kNR][kThreads] halo logits
  const Hier3 h = stage_hier(hg, htab, threadIdx.x, kThreads);
  __syncthreads();

  const int b = blockIdx.z, y0 = blockIdx.y * kTH, x0 = blockIdx.x * kTW;
  const long HW = (long)H * W;
  const int tid = threadIdx.x, lane = tid & 31;
  const long tile_id = ((long)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const unsigned char* lab8 = ws.lab8 + (long)b * HW;
  const unsigned char* flg = ws.flags + (long)b * HW;
  const T* xb = x + (long)b * C * HW;

  // ---- phase-A role: own 4 pixels -------------------------------------------------------------------
  const int ty = tid / kStrips, tx = (tid % kStrips) * 4;
  const int y = y0 + ty, xg = x0 + tx;
  const int px0 = ty * kTW + tx;
  const bool row_ok = y < H;
  const long own_off = (long)y * W + xg;
  int nvalid = row_ok ? W - xg : 0;
  nvalid = nvalid > 4 ? 4 : (nvalid < 0 ? 0 : nvalid);
  const bool st_al = vec_ok && ((W & 3) == 0);

  load_label_tile(labt, kTH, lab8, H, W, y0, x0, h, tid, kThreads);
  int tf[4], tm[4], thh[4];
  unsigned int hsm[4];
  {
    unsigned int u4[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int t = SH_IGNORE, fl = 0;
      if (k < nvalid) { t = lab8[own_off + k]; fl = flg[own_off + k]; }
      tf[k] = t;
      tm[k] = t != SH_IGNORE ? h.f2m[t] : SH_IGNORE;
      thh[k] = t != SH_IGNORE ? h.f2h[t] : SH_IGNORE;
      hsm[k] = t != SH_IGNORE ? h.hsmask[tm[k]] : 0u;
      const unsigned int r3[3] = {t != SH_IGNORE ? (unsigned)t : 0u, t != SH_IGNORE ? (unsigned)tm[k] : 0u,
                                  t != SH_IGNORE ? (unsigned)thh[k] : 0u};
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        unsigned int code = 0xffu;                                   // not an interior anchor
        if (fl & kFlagInterior) code = (fl & (kFlagUniF << l)) ? r3[l] : 0xfeu;
        u4[l] = (u4[l] & ~(0xffu << (8 * k))) | (code << (8 * k));
      }
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) *reinterpret_cast<unsigned int*>(U + l * PX + px0) = u4[l];
  }
  for (int i = tid; i < h.nh * PX; i += kThreads) { maxB[i] = -1.f; holdB[i] = 0; }
  for (int i = tid; i < 2 * kNR * 40; i += kThreads) slow[i] = 0.f;
  // halo slot of this thread (same position in every plane)
  constexpr int nhalo = kPlane - kTH * kTW;   // 264
  int h_sidx = -1;
  long h_goff = -1;
  bool h_valid = false;
  if (tid < nhalo) {
    int r, j;
    if (tid < 2 * kPitch) { r = kTH + tid / kPitch; j = tid % kPitch; }
    else { const int e2 = tid - 2 * kPitch; r = e2 >> 2; const int q = e2 & 3; j = q < 2 ? q : kTW + q; }
    h_sidx = r * kPitch + j;
    const int yy = y0 + r, xx = x0 - 2 + j;
    if (yy < H && xx >= 0 && xx < W) { h_goff = (long)yy * W + xx; h_valid = lab8[h_goff] != SH_IGNORE; }
  }
  __syncthreads();

  // ---- phase-B role: one 4x4 block of plane g ------------------------------------------------------
  const int g = tid / kGroup, u = tid % kGroup, br = u >> 4, bs = u & 15;
  unsigned int ub[3][4];        // U codes of the block rows, per level
  unsigned int pres[3] = {0u, 0u, 0u};
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    bool any_nu = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ub[l][i] = *reinterpret_cast<const unsigned int*>(U + l * PX + (4 * br + i) * kTW + 4 * bs);
#pragma unroll
      for (int k = 0; k < 4; ++k) any_nu |= ((ub[l][i] >> (8 * k)) & 0xffu) == 0xfeu;
    }
    if (any_nu) {
      for (int rr = 0; rr < 8; ++rr) {
        const unsigned int* row = reinterpret_cast<const unsigned int*>(labt + ((l * (kTH + 4)) + 4 * br + rr) * kLabPitch + 4 * bs);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const unsigned int wd = row[q];
          pres[l] |= (1u << (wd & 31)) | (1u << ((wd >> 8) & 31)) | (1u << ((wd >> 16) & 31)) | (1u << ((wd >> 24) & 31));
        }
      }
    }
  }

  // ---- streaming state (phase-A role) ----------------------------------------------------------------
  float sumvF[4], sumvM[4], sumvH[4], prodF[4], prodM[4], prodH[4], a_t[4], b_t[4], c_t[4], min_c[4], runmax[4], validf[4];
  unsigned int runhold[4] = {0u, 0u, 0u, 0u}, hold_minc = 0;
  float lacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // BCE fine/mid/high, CE fine/mid/high
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sumvF[k] = sumvM[k] = sumvH[k] = 0.f;
    prodF[k] = prodM[k] = prodH[k] = 1.f;
    a_t[k] = b_t[k] = c_t[k] = 1.f;
    min_c[k] = 3.0e38f;
    runmax[k] = -1.f;
    validf[k] = tf[k] != SH_IGNORE ? 1.f : 0.f;
  }

  const int nrounds = (C + kNR - 1) / kNR;
  // next round's logits travel global -> shared with cp.async (no registers held across phase B);
  // ragged / unaligned strips fall back to synchronous loads when they are consumed
  const bool fast_own = vec_ok && row_ok && nvalid == 4;
  const bool halo_async = h_goff >= 0 && (sizeof(T) == 4 || vec_ok);
  auto chan_of = [&](unsigned int oe) {
    const int kind = oe & 0xff, cl = (oe >> 8) & 0xff;
    return kind == 0 ? cl : (kind == 1 ? h.nf + cl : h.nf + h.nm + cl);
  };
  auto prefetch = [&](int r) {
#pragma unroll
    for (int j = 0; j < kNR; ++j) {
      const int ci = r * kNR + j;
      if (ci < C) {
        const T* xc = xb + (long)chan_of(h.order[ci]) * HW;
        if (fast_own) cp_async_vec4<T>(xstage + j * kThreads + tid, xc + own_off);
        if (halo_async) cp_async_elem<T>(hstage + j * kThreads + tid, xc, h_goff);
      }
    }
    cp_async_commit();
  };
  auto flush = [&](int r) {
    // combine the warps of every plane of round r and write the per-(tile, channel) records
    const int buf = r & 1;
    for (int q = tid; q < kNR * 56; q += kThreads) {
      const int j = q / 56, k = q % 56, ci = r * kNR + j;
      if (ci >= C) continue;
      float* rec = ws.part1 + ((size_t)tile_id * C + chan_of(h.order[ci])) * kRec;
      const float* rp = red + ((buf * kNR + j) * kGroupWarps) * 20;
      if (k < 13) {
        float a = 0.f;
#pragma unroll
        for (int wq = 0; wq < kGroupWarps; ++wq) a += rp[wq * 20 + k];
        rec[k] = a;
      } else if (k < 15) {
        double a = 0.0;
#pragma unroll
        for (int wq = 0; wq < kGroupWarps; ++wq) a += reinterpret_cast<const double*>(rp + wq * 20 + 16)[k - 13];
        *reinterpret_cast<double*>(rec + kT0 + 2 * (k - 13)) = a;
      } else if (k < 15 + 38) {
        float* sp = slow + (buf * kNR + j) * 40 + (k - 15);
        rec[kLPS + (k - 15)] = *sp;
        *sp = 0.f;
      }
    }
  };

  prefetch(0);
  for (int r = 0; r < nrounds; ++r) {
    const int buf = r & 1;
    // ======================= phase A =======================
    cp_async_wait_all();
#pragma unroll 1     // keep the round body small: the kernel is instruction-cache bound when fully unrolled
    for (int j = 0; j < kNR; ++j) {
      const int ci = r * kNR + j;
      if (ci >= C) break;
      const unsigned int oe = h.order[ci];
      const int kind = oe & 0xff, cl = (oe >> 8) & 0xff, fl = (oe >> 16) & 0xff;
      float* plane = planes + (buf * kNR + j) * kPlane;
      float s[4], v[4], pk[4], xv[4];
      if (fast_own) staged_vec4<T>(xstage + j * kThreads + tid, xv);
      else if (row_ok) load_n<T, 4>(xb + (long)chan_of(oe) * HW, own_off, (long)y * W + W, vec_ok != 0, xv);
      else { xv[0] = xv[1] = xv[2] = xv[3] = 0.f; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const SigExp se = sig_exp(xv[k]);
        s[k] = se.s; v[k] = se.v;
        pk[k] = fmaf(se.s, validf[k], 1e-6f);      // literally probs * valid + 1e-6 (rmi...py:487)
      }
      *reinterpret_cast<float2*>(plane + ty * kPitch + tx + 2) = make_float2(pk[0], pk[1]);
      *reinterpret_cast<float2*>(plane + ty * kPitch + tx + 4) = make_float2(pk[2], pk[3]);
      if (h_sidx >= 0) {
        float p = 0.f;
        if (h_goff >= 0) {
          const float hx = halo_async ? staged_elem<T>(hstage + j * kThreads + tid, h_goff)
                                      : to_f32<T>(xb[(long)chan_of(oe) * HW + h_goff]);
          p = (h_valid ? sig_exp(hx).s : 0.f) + 1e-6f;
        }
        plane[h_sidx] = p;
      }
      if (fl & 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { runmax[k] = -1.f; runhold[k] = 0u; }
      }
      if (kind == 0) {
        const unsigned int c = cl;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sumvF[k] += v[k];
          if (cl == tf[k]) { a_t[k] = s[k]; lacc[3] -= fast_log(v[k]); }
          else prodF[k] *= (1.0f - s[k]) + eps;
          if (s[k] > runmax[k]) { runmax[k] = s[k]; runhold[k] = c; }
        }
        if (fl & 2) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { if (tf[k] != SH_IGNORE) lacc[0] -= fast_log(prodF[k]); prodF[k] = 1.f; }
        }
      } else if (kind == 1) {
        const unsigned int c = h.nf + cl;
        float cur[4];
        unsigned int hd = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sumvM[k] += v[k];
          cur[k] = runmax[k];
          unsigned int hk = runhold[k];
          if (s[k] > cur[k]) { cur[k] = s[k]; hk = c; }   // fine max wins ties
          hd |= hk << (8 * k);
          if (cl == tm[k]) { b_t[k] = s[k]; lacc[4] -= fast_log(v[k]); }
          else prodM[k] *= (1.0f - cur[k]) + eps;
        }
        for (int q = h.mh_ptr[cl]; q < h.mh_ptr[cl + 1]; ++q) {
          const int hh = h.mh_idx[q];
          float4 cb = *reinterpret_cast<float4*>(maxB + (size_t)hh * PX + px0);
          unsigned int hb = *reinterpret_cast<unsigned int*>(holdB + (size_t)hh * PX + px0);
          float cbv[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (cur[k] > cbv[k]) { cbv[k] = cur[k]; hb = (hb & ~(0xffu << (8 * k))) | (hd & (0xffu << (8 * k))); }
          *reinterpret_cast<float4*>(maxB + (size_t)hh * PX + px0) = make_float4(cbv[0], cbv[1], cbv[2], cbv[3]);
          *reinterpret_cast<unsigned int*>(holdB + (size_t)hh * PX + px0) = hb;
        }
        if (nvalid > 0) store4_u8(ws.hold + ((size_t)cl * B + b) * HW + own_off, hd, nvalid, st_al);
        if (fl & 2) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { if (tf[k] != SH_IGNORE) lacc[1] -= fast_log(prodM[k]); prodM[k] = 1.f; }
        }
      } else {
        const unsigned int c = h.nf + h.nm + cl;
        const float4 cb = *reinterpret_cast<const float4*>(maxB + (size_t)cl * PX + px0);
        unsigned int hd = *reinterpret_cast<const unsigned int*>(holdB + (size_t)cl * PX + px0);
        float cur[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sumvH[k] += v[k];
          if (s[k] > cur[k]) { cur[k] = s[k]; hd = (hd & ~(0xffu << (8 * k))) | (c << (8 * k)); }   // mid max wins ties
          if (cl == thh[k]) { c_t[k] = s[k]; lacc[5] -= fast_log(v[k]); }
          else prodH[k] *= (1.0f - cur[k]) + eps;
          if (((hsm[k] >> cl) & 1u) && s[k] < min_c[k]) {
            min_c[k] = s[k];
            hold_minc = (hold_minc & ~(0xffu << (8 * k))) | (c << (8 * k));
          }
        }
        if (nvalid > 0) store4_u8(ws.hold + ((size_t)(h.nm + cl) * B + b) * HW + own_off, hd, nvalid, st_al);
        if (fl & 2) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { if (tf[k] != SH_IGNORE) lacc[2] -= fast_log(prodH[k]); prodH[k] = 1.f; }
        }
      }
    }
    if (r + 1 < nrounds) prefetch(r + 1);
    __syncthreads();
    if (r > 0) flush(r - 1);

    // ======================= phase B =======================
    const int ci = r * kNR + g;
    if (ci < C) {
      const unsigned int oe = h.order[ci];
      const int lvl = oe & 0xff, cl = (oe >> 8) & 0xff;
      const float* plane = planes + (buf * kNR + g) * kPlane;
      float acc[16];          // [0..11] difference taps, [12] count of uniform anchors, rest 0
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      float t0 = 0.f, lpf = 0.f;
      float w[6][8];
#pragma unroll
      for (int rr = 0; rr < 6; ++rr) {
        const float4* p4 = reinterpret_cast<const float4*>(plane + (4 * br + rr) * kPitch + 4 * bs);
        const float4 a = p4[0], bq = p4[1];
        w[rr][0] = a.x; w[rr][1] = a.y; w[rr][2] = a.z; w[rr][3] = a.w;
        w[rr][4] = bq.x; w[rr][5] = bq.y; w[rr][6] = bq.z; w[rr][7] = bq.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned int ui = lvl == 0 ? ub[0][i] : (lvl == 1 ? ub[1][i] : ub[2][i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned int code = (ui >> (8 * k)) & 0xffu;
          const float p = w[i][k + 2];
          const float a = code != 0xffu ? p : 0.f;
          t0 = fmaf(a, p, t0);
          acc[0] = fmaf(a, w[i][k + 3] - p, acc[0]);
          acc[1] = fmaf(a, w[i][k + 4] - p, acc[1]);
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            acc[2 + dx] = fmaf(a, w[i + 1][k + dx] - p, acc[2 + dx]);
            acc[7 + dx] = fmaf(a, w[i + 2][k + dx] - p, acc[7 + dx]);
          }
          const bool hit = code == (unsigned int)cl;
          lpf += hit ? p : 0.f;
          acc[12] += hit ? 1.f : 0.f;
        }
      }
      // warp totals: 13 floats through the transposed reduction, the two big sums in double
      float* rp = red + ((buf * kNR + g) * kGroupWarps + ((tid >> 5) % kGroupWarps)) * 20;
      const float tot = warp_reduce16(acc, lane);
      if ((lane & 1) == 0) {
        const int slot = reduce16_slot(lane);
        if (slot < 13) rp[slot] = tot;
      }
      const double t0d = warp_sum((double)t0), lpd = warp_sum((double)lpf);
      if (lane == 0) {
        reinterpret_cast<double*>(rp + 16)[0] = t0d;
        reinterpret_cast<double*>(rp + 16)[1] = lpd;
      }
      // slow path: anchors whose 5x5 label neighbourhood is not uniform at this level
      const unsigned int pr = lvl == 0 ? pres[0] : (lvl == 1 ? pres[1] : pres[2]);
      const bool want = (pr >> (cl & 31)) & 1u;
      if (__any_sync(0xffffffffu, want)) {
        const unsigned int u0 = lvl == 0 ? ub[0][0] : (lvl == 1 ? ub[1][0] : ub[2][0]);
        const unsigned int u1 = lvl == 0 ? ub[0][1] : (lvl == 1 ? ub[1][1] : ub[2][1]);
        const unsigned int u2 = lvl == 0 ? ub[0][2] : (lvl == 1 ? ub[1][2] : ub[2][2]);
        const unsigned int u3 = lvl == 0 ? ub[0][3] : (lvl == 1 ? ub[1][3] : ub[2][3]);
        pass1_slow_path(plane, labt + (lvl * (kTH + 4)) * kLabPitch, u0, u1, u2, u3, br, bs, cl, want, lane,
                        slow + (buf * kNR + g) * 40);
      }
    }
  }
  __syncthreads();
  flush(nrounds - 1);

  // ---- per-pixel epilogue: positive terms, CE, summaries ----------------------------------------
  {
    unsigned int hp_f = 0, hp_m = 0;
    float iv[3][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // 1 / sum e^x, zero on void pixels (their CE gradient is zero; the fast backward pass relies on it)
      const float vk = tf[k] != SH_IGNORE ? 1.f : 0.f;
      iv[0][k] = rcp(sumvF[k]) * vk; iv[1][k] = rcp(sumvM[k]) * vk; iv[2][k] = rcp(sumvH[k]) * vk;
      if (tf[k] != SH_IGNORE) {
        const bool a_holds = a_t[k] <= b_t[k];                 // fine wins ties (rmi...py:421-425)
        const float mcla = a_holds ? a_t[k] : b_t[k];
        hp_f |= (unsigned int)(a_holds ? tf[k] : h.nf + tm[k]) << (8 * k);
        const bool c_holds = min_c[k] <= b_t[k];               // high wins ties (rmi...py:439-440)
        const float mclb = c_holds ? min_c[k] : b_t[k];
        hp_m |= (c_holds ? ((hold_minc >> (8 * k)) & 0xffu) : (unsigned int)(h.nf + tm[k])) << (8 * k);
        lacc[0] -= fast_log(mcla + eps);
        lacc[1] -= fast_log(mclb + eps);
        lacc[2] -= fast_log(c_t[k] + eps);
        lacc[3] += fast_log(sumvF[k]);
        lacc[4] += fast_log(sumvM[k]);
        lacc[5] += fast_log(sumvH[k]);
      }
    }
    if (nvalid > 0) {
      store4_u8(ws.hold + ((size_t)(h.nm + h.nh) * B + b) * HW + own_off, hp_f, nvalid, st_al);
      store4_u8(ws.hold + ((size_t)(h.nm + h.nh + 1) * B + b) * HW + own_off, hp_m, nvalid, st_al);
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        float* ip = ws.inv + ((size_t)l * B + b) * HW + own_off;
        if (st_al && nvalid == 4) *reinterpret_cast<float4*>(ip) = make_float4(iv[l][0], iv[l][1], iv[l][2], iv[l][3]);
        else for (int k = 0; k < nvalid; ++k) ip[k] = iv[l][k];
      }
    }
  }
  __syncthreads();
  const float r6 = block_sum_k<6>(lacc, planes);   // planes are free now (needs 6 * 16 floats)
  if (tid < 6) ws.bcepart[(size_t)tile_id * 8 + tid] = r6;
  else if (tid < 8) ws.bcepart[(size_t)tile_id * 8 + tid] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// k3_band: P of the 4-pixel border bands of every (b, c) plane.
//   bandR[bc][q][x]  q = 0..3 -> rows 0..3, q = 4..7 -> rows H-4..H-1
//   bandC[bc][q][y]  q = 0..3 -> cols 0..3, q = 4..7 -> cols W-4..W-1
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k3_band(const T* __restrict__ x, int B, int C, int H, int W,
                                               const unsigned char* __restrict__ lab8_all, float* __restrict__ bandR,
                                               float* __restrict__ bandC, Hier3 h, unsigned char* __restrict__ labB) {
  const int bc = blockIdx.x, b = bc / C, c = bc - b * C;
  // the first channel of each level also writes the level's RMI labels of the bands (Ws3::labB)
  const int lvl = c == 0 ? 0 : (c == h.nf ? 1 : (c == h.nf + h.nm ? 2 : -1));
  const int* lmap = lvl == 1 ? h.f2m : h.f2h;
  unsigned char* lbo = labB + ((size_t)b * 3 + (lvl < 0 ? 0 : lvl)) * 8 * ((size_t)W + H);
  const long HW = (long)H * W;
  const T* xc = x + (long)bc * HW;
  const unsigned char* lab8 = lab8_all + (long)b * HW;
  const int n = 8 * W + 8 * H;
#pragma unroll 4
  for (int i = blockIdx.y * 256 + threadIdx.x; i < n; i += gridDim.y * 256) {
    int yy, xx;
    float* dst;
    if (i < 8 * W) { const int q = i / W; xx = i - q * W; yy = q < 4 ? q : H - 8 + q; dst = bandR + ((size_t)bc * 8 + q) * W + xx; }
    else { const int i2 = i - 8 * W; const int q = i2 / H; yy = i2 - q * H; xx = q < 4 ? q : W - 8 + q; dst = bandC + ((size_t)bc * 8 + q) * H + yy; }
    const long off = (long)yy * W + xx;
    const float xv = to_f32<T>(xc[off]);          // both loads issue together (the kernel is load-latency bound)
    const int t = lab8[off];
    const bool valid = t != SH_IGNORE;
    *dst = (valid ? sig_exp(xv).s : 0.f) + 1e-6f;
    if (lvl >= 0) lbo[i] = (unsigned char)(valid ? (lvl == 0 ? t : lmap[t]) : 0);
  }
}

// ---------------------------------------------------------------------------------------------
// k3_frame1: taps of the image frame.  grid (B*C, nseg), block 256 = 8 warps; warp w owns run w
// (rows 0,1,H-2,H-1 x middle columns; columns 0,1,W-2,W-1 x middle rows); warp 0 of segment 0
// then does the 16 corner pixels.  Output: frameT[seg][b*C+c][class][kFrameRec].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void frame_pixel_taps(const BandView& bv, int cl, int y, int x, double& t0, double& lp0,
                                                 float (&acc)[75]) {
  const float pr = bv.P(y, x);
  const float lr = bv.L(y, x) == cl ? 1.f : 0.f;
  t0 += (double)pr * (double)pr;
  lp0 += (double)(pr * lr);
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const int yy = y + dy, xx = x + dx;
      if (yy < 0 || yy >= bv.H || xx < 0 || xx >= bv.W) continue;   // such taps have no valid window
      const float pn = bv.P(yy, xx);
      const float ln = bv.L(yy, xx) == cl ? 1.f : 0.f;
      const int t = (dy + 2) * 5 + dx + 2;
      acc[t] = fmaf(pr, pn - pr, acc[t]);
      acc[25 + t] = fmaf(pr, ln - lr, acc[25 + t]);
      acc[50 + t] = fmaf(lr, ln, acc[50 + t]);
    }
  }
}

__global__ void __launch_bounds__(256) k3_frame1(int B, int H, int W, Hier3 h, Ws3 ws, const float* __restrict__ bandR,
                                                 const float* __restrict__ bandC) {
  __shared__ BandSeg bs[4];
  const int C = h.nf + h.nm + h.nh;
  const int bc = blockIdx.x, b = bc / C, c = bc % C;
  const int seg = blockIdx.y, nseg = gridDim.y;
  const int lvl = c < h.nf ? 0 : (c < h.nf + h.nm ? 1 : 2);
  const int cl = lvl == 0 ? c : (lvl == 1 ? c - h.nf : c - h.nf - h.nm);
  BandView bv;
  bv.bandR = bandR + (size_t)bc * 8 * W; bv.bandC = bandC + (size_t)bc * 8 * H;
  bv.lab8 = ws.lab8 + (long)b * H * W; bv.lmap = lvl == 0 ? nullptr : (lvl == 1 ? h.f2m : h.f2h);
  bv.H = H; bv.W = W;
  const unsigned char* lb = ws.labB + ((size_t)b * 3 + lvl) * 8 * ((size_t)W + H);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* out = ws.frameT + ((size_t)seg * B * C + bc) * 25 * kFrameRec;

  // warp -> line: warps 0,1 rows 0,1 ; 2,3 rows H-2,H-1 ; 4,5 cols 0,1 ; 6,7 cols W-2,W-1
  const int side = warp >> 1, va = (warp & 1) + 2 * (side & 1);
  const bool is_row = side < 2;
  float acc[75];     // canonical order: index (dv + 2) * 5 + (du + 2), dv across the band, du along it
#pragma unroll
  for (int i = 0; i < 75; ++i) acc[i] = 0.f;
  double t0 = 0.0, lp0 = 0.0;
  int lo[2], hi[2];
#pragma unroll
  for (int a = 0; a < 2; ++a) {      // anchors are the middle pixels of the line: 2 .. N-3
    const int len = (a == 0 ? W : H) - 4, per = (len + nseg - 1) / nseg;
    lo[a] = 2 + seg * per;
    hi[a] = min(2 + len, lo[a] + per);
  }
  const int span = max(hi[0] - lo[0], hi[1] - lo[1]);
  for (int c0 = 0; c0 < span; c0 += kSegMax) {
    __syncthreads();
#pragma unroll 1
    for (int sd = 0; sd < 4; ++sd) {
      const int a = sd >> 1, u0 = lo[a] + c0, n = min(kSegMax, hi[a] - u0);
      if (n > 0) stage_band(bs[sd], sd, u0, n, bv.bandR, bv.bandC, lb, cl, H, W, tid, 256);
    }
    __syncthreads();
    const int a = side >> 1, u0 = lo[a] + c0, n = min(kSegMax, hi[a] - u0);
    const BandSeg& s = bs[side];
    for (int i = lane; i < n; i += 32) {
      const float pr = s.P[va][i + 2], lr = s.L[va][i + 2];
      t0 += (double)pr * (double)pr;
      lp0 += (double)(pr * lr);
#pragma unroll
      for (int dv = -2; dv <= 2; ++dv) {
        const int vv = va + dv;
        if (vv < 0 || vv > 3) continue;                      // such taps have no valid window
#pragma unroll
        for (int du = -2; du <= 2; ++du) {
          const float pn = s.P[vv][i + 2 + du], ln = s.L[vv][i + 2 + du];
          const int t = (dv + 2) * 5 + du + 2;
          acc[t] = fmaf(pr, pn - pr, acc[t]);
          acc[25 + t] = fmaf(pr, ln - lr, acc[25 + t]);
          acc[50 + t] = fmaf(lr, ln, acc[50 + t]);
        }
      }
    }
  }
  const int kline = (warp & 3) < 2 ? (warp & 3) : 1 + (warp & 3);  // class 0,1,3,4
  const int cls = is_row ? kline * 5 + 2 : 2 * 5 + kline;
  float* oc = out + cls * kFrameRec;
#pragma unroll
  for (int i = 0; i < 75; ++i) {
    const float r = warp_sum(acc[i]);
    // canonical (dv, du) -> tap (dy, dx): rows: dy = dv, dx = du ; columns: dy = du, dx = dv
    const int grp = i / 25, t = i % 25, tt = is_row ? t : (t % 5) * 5 + t / 5;
    if (lane == 0) oc[kFD + grp * 25 + tt] = r;
  }
  t0 = warp_sum(t0); lp0 = warp_sum(lp0);
  if (lane == 0) { reinterpret_cast<double*>(oc)[0] = t0; reinterpret_cast<double*>(oc)[1] = lp0; }
  if (warp == 0) {
    // corners: one lane per pixel, each its own class
#pragma unroll
    for (int i = 0; i < 75; ++i) acc[i] = 0.f;
    t0 = 0.0; lp0 = 0.0;
    const int ky = (lane >> 2) & 3, kx = lane & 3;
    const int cy = ky < 2 ? ky : ky + 1, cx = kx < 2 ? kx : kx + 1;  // classes 0,1,3,4
    if (lane < 16 && seg == 0)
      frame_pixel_taps(bv, cl, ky < 2 ? ky : H - 4 + ky, kx < 2 ? kx : W - 4 + kx, t0, lp0, acc);
    if (lane < 16) {
      float* o2 = out + (cy * 5 + cx) * kFrameRec;
      reinterpret_cast<double*>(o2)[0] = t0; reinterpret_cast<double*>(o2)[1] = lp0;
      for (int i = 0; i < 75; ++i) o2[kFD + i] = acc[i];
    }
    if (lane == 16) {
      float* o2 = out + (2 * 5 + 2) * kFrameRec;
      for (int i = 0; i < kFrameRec; ++i) o2[i] = 0.f;   // interior class unused
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k3_finalize: one CTA (256 threads) per (b, c).
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

// frame records of (b, c) summed over the segments -> fr[25][kFrameRec] (doubles)
__device__ void load_frame_records(double* fr, int bc, int B, int C, int nseg, const Ws3& ws, int tid) {
  for (int e = tid; e < 25 * kFrameRec; e += 256) {
    const int k = e % kFrameRec;
    double a = 0.0;
    if (k < 4) {
      if ((k & 1) == 0)
        for (int s = 0; s < nseg; ++s)
          a += *reinterpret_cast<const double*>(ws.frameT + ((size_t)s * B * C + bc) * 25 * kFrameRec + e);
    } else {
      for (int s = 0; s < nseg; ++s) a += (double)ws.frameT[((size_t)s * B * C + bc) * 25 * kFrameRec + e];
    }
    fr[e] = a;
  }
}

__device__ void finalize_core(const double* part, const double* fr, int bc, const Ws3& ws, double scale, int tid);

__global__ void __launch_bounds__(256) k3_finalize(int B, int C, long tiles_per_img, int nseg, Ws3 ws, double scale) {
  __shared__ double part4[4][kRec];
  __shared__ double part[kRec];
  __shared__ double fr[25 * kFrameRec];
  const int bc = blockIdx.x, b = bc / C, c = bc % C;
  const int tid = threadIdx.x;
  {
    // 4 slices x 64 record slots; slots kT0 .. kT0+3 hold two doubles
    const int k = tid & 63, sl = tid >> 6;
    const bool is_dbl = k >= kT0 && k < kT0 + 4;
    double a = 0.0;
    const float* p = ws.part1 + ((size_t)b * tiles_per_img * C + c) * kRec;
    if (!is_dbl) {
      for (long t = sl; t < tiles_per_img; t += 4) a += (double)p[(size_t)t * C * kRec + k];
    } else if ((k & 1) == 0) {
      for (long t = sl; t < tiles_per_img; t += 4) a += *reinterpret_cast<const double*>(p + (size_t)t * C * kRec + k);
    }
    part4[sl][k] = a;
  }
  load_frame_records(fr, bc, B, C, nseg, ws, tid);
  __syncthreads();
  if (tid < kRec) part[tid] = part4[0][tid] + part4[1][tid] + part4[2][tid] + part4[3][tid];
  __syncthreads();
  finalize_core(part, fr, bc, ws, scale, tid);
}

// fast path: records are per (persistent CTA, channel): fp64 product taps + label-anchored lp taps
// (ws.rec2) and integer label-label counts (ws.llrec); rewritten into the record layout of the generic path
__global__ void __launch_bounds__(256) k3f_finalize(int B, int C, int cpi, int nseg, Ws3 ws, double scale) {
  __shared__ double part[kRec];
  __shared__ double fr[25 * kFrameRec];
  __shared__ double raw[64];
  const int bc = blockIdx.x, b = bc / C, c = bc % C;
  const int tid = threadIdx.x;
  if (tid < 64) {
    double a = 0.0;
    if (tid < 38) {
      for (int j = 0; j < cpi; ++j) a += ws.rec2[((size_t)(b * cpi + j) * C + c) * kFastRec + tid];
    } else if (tid < 38 + 13) {
      unsigned long long n = 0;
      for (int j = 0; j < cpi; ++j) n += ws.llrec[((size_t)(b * cpi + j) * C + c) * 16 + (tid - 38)];
      a = (double)n;
    }
    raw[tid] = a;
  }
  load_frame_records(fr, bc, B, C, nseg, ws, tid);
  __syncthreads();
  if (tid < kRec) {
    double v = 0.0;
    if (tid >= kD && tid < kD + 12) v = raw[1 + (tid - kD)] - raw[0];   // difference form of the generic records
    else if (tid == kT0) v = raw[0];
    else if (tid >= kLPS && tid < kLPS + 25) v = raw[13 + (tid - kLPS)];
    else if (tid >= kLLS && tid < kLLS + 13) v = raw[38 + (tid - kLLS)];
    part[tid] = v;                                                       // kLLFull, kLPFull = 0
  }
  __syncthreads();
  finalize_core(part, fr, bc, ws, scale, tid);
}

__device__ void finalize_core(const double* part, const double* fr, int bc, const Ws3& ws, double scale, int tid) {
  __shared__ double Spp[81], Slp[81], Sll[81], K[81], T1[81], M[81], Wm[81], U[81], Gpp[81], tmp[81];
  __shared__ double piv[9], fac[9];
  // assemble
  for (int e = tid; e < 81; e += 256) {
    const int i = e / 9, j = e % 9;
    const int yi = i / 3, xi = i % 3, yj = j / 3, xj = j % 3;
    const int dy = yi - yj, dx = xi - xj;
    const int t = tap_index(dy, dx);
    double fpp = 0.0, flp = 0.0, fll = 0.0;
    for (int ky = 0; ky < 5; ++ky)
      for (int kx = 0; kx < 5; ++kx) {
        if (ky == 2 && kx == 2) continue;
        if (!offset_valid(ky, yj) || !offset_valid(kx, xj)) continue;
        const double* f = fr + (ky * 5 + kx) * kFrameRec;
        fpp += f[kFT0] + f[kFD + t];
        flp += f[kFLP0] + f[kFDL + t];
        fll += f[kFLL + t];
      }
    Slp[e] = part[kLPFull] + part[kLPS + t] + flp;
    if (in_half_plane(dy, dx)) {
      const int ht = half_tap_index(dy, dx);
      Spp[e] = part[kT0] + (ht > 0 ? part[kD + ht - 1] : 0.0) + fpp;
      Sll[e] = part[kLLFull] + part[kLLS + ht] + fll;
    }
  }
  __syncthreads();
  for (int e = tid; e < 81; e += 256) {
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
    mm9(T1, true, U, false, Gpp, lane);            // 2 * G_pp = T1^T U
  }
  __syncthreads();
  // adjoints -> stencil weights, spread over the whole CTA.  Tap t = (dy, dx) collects the matrix entries
  // e = (i, j) with window offsets i = j + (dy, dx): at most 9, walked in ascending e.
  float* wout = ws.wts + (size_t)bc * 64;
  float* fout = ws.fwts + (size_t)bc * 25 * 50;
  __shared__ double s_w[50];
  if (tid < 25) {
    double w1 = 0.0, w2 = 0.0;
    const int dy = tid / 5 - 2, dx = tid % 5 - 2;
    for (int yj = 0; yj < 3; ++yj)
      for (int xj = 0; xj < 3; ++xj) {
        const int yi = yj + dy, xi = xj + dx;
        if (yi < 0 || yi > 2 || xi < 0 || xi > 2) continue;
        const int e = (yi * 3 + xi) * 9 + yj * 3 + xj;
        w1 += Gpp[e]; w2 -= U[e];
      }
    wout[tid] = (float)(scale * w1);
    wout[25 + tid] = (float)(scale * w2);
    s_w[tid] = scale * w1;
    s_w[25 + tid] = scale * w2;
  }
  __syncthreads();
  if (tid == 32) {
    double s2 = 0.0, s1 = 0.0;
    for (int t = 0; t < 25; ++t) { s1 += s_w[t]; s2 += s_w[25 + t]; }
    wout[50] = (float)s2;      // sum of the lp weights: label-uniform pixels
    wout[51] = (float)s1;      // sum of the pp weights: difference form of the stencil
  }
  for (int q = tid; q < 25 * 25; q += 256) {
    const int cls = q / 25, t = q % 25, ky = cls / 5, kx = cls % 5;
    double w1 = 0.0, w2 = 0.0;
    const int dy = t / 5 - 2, dx = t % 5 - 2;
    for (int yj = 0; yj < 3; ++yj) {
      const int yi = yj + dy;
      if (yi < 0 || yi > 2 || !offset_valid(ky, yj)) continue;
      for (int xj = 0; xj < 3; ++xj) {
        const int xi = xj + dx;
        if (xi < 0 || xi > 2 || !offset_valid(kx, xj)) continue;
        const int e = (yi * 3 + xi) * 9 + yj * 3 + xj;
        w1 += Gpp[e]; w2 -= U[e];
      }
    }
    fout[cls * 50 + t] = (float)(scale * w1);
    fout[cls * 50 + 25 + t] = (float)(scale * w2);
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
    // a present class in neither id list: the reference raises ValueError from list.remove; without a host
    // sync the device-side equivalent is a poisoned loss (strict=True on the module turns it into the exception)
    if (ready != nullptr && ready[1] != 0) loss = __longlong_as_double(0x7ff8000000000000LL);
    out[0] = (float)loss;
    out[1] = (float)tscale;
    out[2] = r;
  }
}

// one CTA per output k: out[k] = sum_i part[i][k] in double, fixed order
__global__ void __launch_bounds__(256) k_reduce_partials3(const float* __restrict__ part, long n, int K,
                                                          double* __restrict__ out) {
  __shared__ double sm[256];
  const int k = blockIdx.x;
  double a = 0.0;
  for (long i = threadIdx.x; i < n; i += 256) a += (double)part[(size_t)i * K + k];
  sm[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[k] = sm[0];
}

static size_t pass1_smem_bytes(int nh) {
  size_t s = (size_t)2 * kNR * (kTH + 2) * kPitch * 4 + 2 * kNR * kGroupWarps * 20 * 4 + 2 * kNR * 40 * 4;
  s += (size_t)nh * kTH * kTW * 5 + 3 * kTH * kTW + 3 * (kTH + 4) * kLabPitch + 6144 /* hierarchy tables */;
  s = (s + 15) & ~(size_t)15;
  s += (size_t)kNR * kThreads * (16 + 4);   // cp.async staging of the next round's logits
  return (s + 15) & ~(size_t)15;
}

size_t fast_bwd_smem(int C, int nf, int nm, int nh);   // rmi3_bwd.cu

// Can the warp-specialised kernels of rmi3_fast.cuh run this problem?
bool fast_path_ok(const void* x, const void* grad, int elem, int H, int W, int nf, int nm, int nh, int fast_tab_ok) {
  const int C = nf + nm + nh;
  if (!fast_tab_ok || C <= fast::NR || C > fast::kFastMaxC) return false;
  if (fast::pass1_smem(C, nf, 2) > 227 * 1024) return false;
  if ((W & 3) != 0 || H < 8 || W < 8) return false;
  if ((uintptr_t)x % (4 * (size_t)elem) != 0) return false;
  if (grad != nullptr && (uintptr_t)grad % (4 * (size_t)elem) != 0) return false;
  return true;
}

template <typename T, typename L>
static int run_forward3_fast(const void* x, const L* label, int B, int H, int W, const Hier3& h, const Ws3& ws,
                             float* bandR, float* bandC, float eps, double scale, int stages, cudaStream_t st) {
  const int C = h.nf + h.nm + h.nh;
  const bool fast_bwd_ok = fast_bwd_smem(C, h.nf, h.nm, h.nh) <= 227 * 1024;
  fast::FastHier fh;
  fh.nf = h.nf; fh.nm = h.nm; fh.nh = h.nh; fh.f2m = h.f2m; fh.f2h = h.f2h; fh.order = h.order + C;
  const int cpi = ws.cpi, grid = B * cpi;
  if (stages & 1) {
    cudaError_t e = cudaMemsetAsync(ws.counts, 0, 4 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    // uint8 labels, #valid, range check and the label-label counts in one kernel; the per-pixel uniformity flags of
    // k3_prep are only read by the generic kernels
    const size_t smem = fast::prep_smem(C, h.nf);
    cudaFuncSetAttribute(fast::k3f_prep<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e = cudaMemsetAsync(ws.llrec, 0, (size_t)grid * C * 16 * sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(ws.strips, 0, (size_t)B * 2 * sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    fast::k3f_prep<L><<<grid * fast::PREP_MULT, 256, smem, st>>>(label, B, H, W, fh, ws, cpi,
                                                                (uintptr_t)label % 16 == 0 ? 1 : 0);
    SH_CHECK_LAUNCH();
    if (!fast_bwd_ok) {     // the backward pass will run the generic kernel, which wants the flags
      dim3 gp((W + kTW - 1) / kTW, (H + 15) / 16, B);
      k3_prep<L><<<gp, 256, 0, st>>>(label, B, H, W, h, ws.lab8, ws.flags, nullptr);
      SH_CHECK_LAUNCH();
    }
  }
  if (stages & 2) {
    const int nbuf = fast::pass1_smem(C, h.nf, fast::NBUF) <= 227 * 1024 ? fast::NBUF : 2;
    const size_t smem = fast::pass1_smem(C, h.nf, nbuf);
    if (smem > 227 * 1024) return SH_ERR_UNSUPPORTED;
    auto kern = fast::k3f_pass1<T>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, fast::NTHREADS, smem, st>>>((const T*)x, B, H, W, fh, ws, eps, cpi, nbuf);
    SH_CHECK_LAUNCH();
  }
  if (stages & 4) {
    k3_band<T><<<dim3(B * C, 4), 256, 0, st>>>((const T*)x, B, C, H, W, ws.lab8, bandR, bandC, h, ws.labB);
    SH_CHECK_LAUNCH();
    k3_frame1<<<dim3(B * C, ws.nseg), 256, 0, st>>>(B, H, W, h, ws, bandR, bandC);
    SH_CHECK_LAUNCH();
  }
  if (stages & 8) {
    k_reduce_partials3<<<6, 256, 0, st>>>(ws.bce2, (long)grid, 8, ws.sums);
    SH_CHECK_LAUNCH();
    k3f_finalize<<<B * C, 256, 0, st>>>(B, C, cpi, ws.nseg, ws, scale);
    SH_CHECK_LAUNCH();
  }
  return SH_OK;
}

template <typename T, typename L>
static int run_forward3(const void* x, const L* label, int B, int H, int W, const Hier3& h, const Ws3& ws,
                        float* bandR, float* bandC, float eps, double scale, int stages, cudaStream_t st) {
  const int C = h.nf + h.nm + h.nh;
  if (stages & 1) {
    cudaError_t e = cudaMemsetAsync(ws.counts, 0, 4 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(ws.strips, 0, (size_t)B * 2 * sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    dim3 gp((W + kTW - 1) / kTW, (H + 15) / 16, B);
    k3_prep<L><<<gp, 256, 0, st>>>(label, B, H, W, h, ws.lab8, ws.flags, ws.counts);
    SH_CHECK_LAUNCH();
  }
  if (stages & 2) {
    const size_t smem = pass1_smem_bytes(h.nh);
    if (smem > 227 * 1024) return SH_ERR_UNSUPPORTED;
    const bool vec_ok = ((W & 3) == 0) && ((uintptr_t)x % (4 * sizeof(T)) == 0);
    auto kern = k3_pass1<T>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 g1(ws.tiles_x, ws.tiles_y, B);
    kern<<<g1, kThreads, smem, st>>>((const T*)x, B, H, W, h, ws, eps, vec_ok ? 1 : 0);
    SH_CHECK_LAUNCH();
  }
  if (stages & 4) {
    k3_band<T><<<dim3(B * C, 4), 256, 0, st>>>((const T*)x, B, C, H, W, ws.lab8, bandR, bandC, h, ws.labB);
    SH_CHECK_LAUNCH();
    k3_frame1<<<dim3(B * C, ws.nseg), 256, 0, st>>>(B, H, W, h, ws, bandR, bandC);
    SH_CHECK_LAUNCH();
  }
  if (stages & 8) {
    const long ntiles = (long)ws.tiles_x * ws.tiles_y * B;
    k_reduce_partials3<<<6, 256, 0, st>>>(ws.bcepart, ntiles, 8, ws.sums);
    SH_CHECK_LAUNCH();
    k3_finalize<<<B * C, 256, 0, st>>>(B, C, (long)ws.tiles_x * ws.tiles_y, ws.nseg, ws, scale);
    SH_CHECK_LAUNCH();
  }
  return SH_OK;
}

}  // namespace sh

extern "C" {

// workspace = Ws3 fields, then bandR [B*C][8][W] and bandC [B*C][8][H] floats
size_t sh_rmi3_workspace_bytes(int B, int H, int W, int nf, int nm, int nh) {
  const size_t base = sh::ws3_layout(nullptr, B, H, W, nf, nm, nh).bytes;
  return base + sh::align256((size_t)B * (nf + nm + nh) * 8 * ((size_t)W + H) * 4);
}

// Byte offsets of the workspace fields (diagnostics / tests):
// out[0..10] = counts, sums, lab8, flags, hold, inv, part1, bcepart, frameT, rbc, wts; out[11] = fwts
int sh_rmi3_workspace_offsets(int B, int H, int W, int nf, int nm, int nh, size_t* out) {
  sh::Ws3 w = sh::ws3_layout(nullptr, B, H, W, nf, nm, nh);
  const void* f[12] = {w.counts, w.sums, w.lab8, w.flags, w.hold, w.inv, w.part1, w.bcepart, w.frameT, w.rbc, w.wts, w.fwts};
  for (int i = 0; i < 12; ++i) out[i] = (size_t)f[i];
  return SH_OK;
}

// hier_tab (device int32): [f2m nf][f2h nf][mh_ptr nm+1][mh_idx n_mh][hsmask nm][order C]
int sh_rmi3_fast_path(const void* logits, const void* grad, int dtype, int H, int W, int nf, int nm, int nh,
                      int fast_tab_ok) {
  return sh::fast_path_ok(logits, grad, dtype == SH_DT_F32 ? 4 : 2, H, W, nf, nm, nh, fast_tab_ok) ? 1 : 0;
}

int sh_rmi3_forward(const void* logits, int dtype, const void* label, int label_dtype, int B, int H, int W, int nf, int nm, int nh,
                    const int* hier_tab, int n_mh, int fast_tab_ok, float lam, float loss_weight, void* workspace,
                    int stages, void* stream) {
  if (B <= 0 || H < 8 || W < 8 || nf <= 0 || nm <= 0 || nh <= 0 || nf + nm + nh > 254 || nh > 32)
    return SH_ERR_BAD_ARG;
  if (sh::fast_path_ok(logits, nullptr, dtype == SH_DT_F32 ? 4 : 2, H, W, nf, nm, nh, fast_tab_ok)) {
    sh::Ws3 wsf = sh::ws3_layout(workspace, B, H, W, nf, nm, nh);
    float* bR = (float*)((unsigned char*)workspace + wsf.bytes);
    float* bC = bR + (size_t)B * (nf + nm + nh) * 8 * W;
    sh::Hier3 hf = sh::hier3_from_tab(hier_tab, nf, nm, nh, n_mh);
    const double sc = (double)lam * (double)loss_weight / (9.0 * B);
    cudaStream_t s2 = (cudaStream_t)stream;
    switch (dtype) {
      case SH_DT_F32: SH_LABEL_SWITCH(label_dtype, L, { return sh::run_forward3_fast<float, L>(logits, (const L*)label, B, H, W, hf, wsf, bR, bC, 1e-6f, sc, stages, s2); }) break;
      case SH_DT_BF16: SH_LABEL_SWITCH(label_dtype, L, { return sh::run_forward3_fast<__nv_bfloat16, L>(logits, (const L*)label, B, H, W, hf, wsf, bR, bC, 1e-6f, sc, stages, s2); }) break;
      case SH_DT_F16: SH_LABEL_SWITCH(label_dtype, L, { return sh::run_forward3_fast<__half, L>(logits, (const L*)label, B, H, W, hf, wsf, bR, bC, 1e-6f, sc, stages, s2); }) break;
    }
    return SH_ERR_UNSUPPORTED;
  }
  sh::Ws3 ws = sh::ws3_layout(workspace, B, H, W, nf, nm, nh);
  float* bandR = (float*)((unsigned char*)workspace + ws.bytes);
  float* bandC = bandR + (size_t)B * (nf + nm + nh) * 8 * W;
  sh::Hier3 h = sh::hier3_from_tab(hier_tab, nf, nm, nh, n_mh);
  const double scale = (double)lam * (double)loss_weight / (9.0 * B);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case SH_DT_F32: SH_LABEL_SWITCH(label_dtype, L, { return sh::run_forward3<float, L>(logits, (const L*)label, B, H, W, h, ws, bandR, bandC, 1e-6f, scale, stages, st); }) break;
    case SH_DT_BF16: SH_LABEL_SWITCH(label_dtype, L, { return sh::run_forward3<__nv_bfloat16, L>(logits, (const L*)label, B, H, W, h, ws, bandR, bandC, 1e-6f, scale, stages, st); }) break;
    case SH_DT_F16: SH_LABEL_SWITCH(label_dtype, L, { return sh::run_forward3<__half, L>(logits, (const L*)label, B, H, W, h, ws, bandR, bandC, 1e-6f, scale, stages, st); }) break;
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
