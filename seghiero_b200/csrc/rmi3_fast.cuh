// Fast path of the three-level (BCE + RMI + CE) loss, forward side: warp-specialised persistent kernel + label prep.
//
// Applies when the hierarchy is a tree (every mid has one high), W % 4 == 0, pointers are 16-byte
// aligned and NR < C <= kFastMaxC; everything else runs the generic kernels of rmi3_fwd.cu / rmi3_bwd.cu.
// Reference arithmetic: models/loss/rmi_hiera_triplet_loss.py:323-546.
// The backward side of the fast path is rmi3_fast_bwd.cuh.
//
// k3f_pass1: one CTA per SM (512 threads), each walking tiles (64 x 18 pixels) of ONE image:
//   warps 0-8   producers : thread = 4 consecutive pixels, all channels in tree order (fine children,
//                           their mid, ..., then the high).  sigmoid / e^x from 3 MUFU ops, tree BCE +
//                           CE sums in registers, P = s*valid + 1e-6 -> shared-memory channel plane.
//   warps 9-15  consumers : warp = one channel plane of the round, thread = 4 x 9 block.  The 2-pixel ring of
//                           the plane (sigmoid only, own cp.async ring), the tile's label bytes, then RMI moments
//                           of the interior anchors as 13 product taps (pr_cov) + 25 label-anchored taps
//                           (la_pr), warp-reduced and accumulated in fp64 per CTA.
// Planes travel producer -> consumer through a ring of NBUF round buffers (NR = 7 planes each, 28 channels =
// 4 rounds) guarded by named barriers (bar.arrive / bar.sync); logits are prefetched with cp.async.
#pragma once
#include "rmi3_common.cuh"

namespace sh {
namespace fast {

constexpr int TW = 64;               // tile width (pixels)
constexpr int PWARPS = 9;            // producer warps
constexpr int CWARPS = 7;            // consumer warps
constexpr int PPC = 1;               // planes per consumer warp and round
constexpr int TH = 2 * PWARPS;       // tile height of the streaming kernels: one producer thread per 4-pixel strip
constexpr int BR = TH / 2;           // rows of a consumer thread's 4-wide block
constexpr int PW = TW + 4;           // plane pitch: cols x0-2 .. x0+65
constexpr int PR = TH + 4;           // plane rows:  y0-2 .. y0+TH+1
constexpr int PLANE = PR * PW;
constexpr int NR = CWARPS * PPC;     // planes per round
constexpr int NBUF = 3;              // round buffers (2 when the per-channel fp64 totals of a large hierarchy need the room)
constexpr int NPROD = 32 * PWARPS, NCONS = 32 * CWARPS;
constexpr int NTHREADS = NPROD + NCONS;   // 512 threads, 128 registers each
constexpr int XD = 4;                // cp.async ring depth of the producers (power of two; XD-2 channels ahead)
constexpr int HD = 4;                // cp.async ring depth of the consumers' halo logits (rounds; HD-1 ahead)
constexpr int HSLOT = 48;            // staged halo bytes per lane and plane: 2 x 16 (strips) + 2 x 8 (column pairs)
constexpr int BAR_FULL = 1, BAR_EMPTY = 1 + NBUF, BAR_PROD = 1 + 2 * NBUF, BAR_CONS = 2 + 2 * NBUF;
constexpr int kFastMaxC = 254;

// order entry: kind (0 fine / 1 mid / 2 high) | class << 8 | flags << 16 | channel << 24 ; flags bit1 = flush products
struct FastHier {
  int nf, nm, nh;
  const int* f2m;            // [nf]
  const int* f2h;            // [nf]
  const unsigned int* order; // [C] tree order of the fast path
};

struct TileCoord { int y0, x0; };
__device__ __forceinline__ TileCoord tile_coord(int tile, int tiles_x, int th) {
  TileCoord t;
  const int tyi = tile / tiles_x;
  t.y0 = tyi * th;
  t.x0 = (tile - tyi * tiles_x) * TW;
  return t;
}

// -------------------------------------------------------------------------------------------------
// k3f_prep: labels int64 -> uint8, #valid, range check, and the label-only part of RMI:
//   ll[c][d] = #{ interior anchors r : L(r) = c and L(r + d) = c },  d in the half plane (13 taps),
// with L = the RMI label of the level (void pixels are class 0, rmi...py:360-370).  Counted as
//   ll[c][d] = A[c] - M[c][d],   A[c] = #anchors of class c,  M[c][d] = #anchors of class c whose tap d is NOT c:
// a strip whose 3 x 8 label window is uniform adds to A only (run-length accumulated in registers along the
// thread's column); only strips on a label boundary look at single taps.  Persistent CTAs (cpi per image),
// integer counts per (CTA, channel) -> ws.llrec (summed by k3f_finalize).
// -------------------------------------------------------------------------------------------------
constexpr int LPITCH = TW + 8;   // label tile pitch: cols x0-2 .. x0+65 (+4 pad)
constexpr int PTH2 = 32;         // tile height of k3f_prep

constexpr int PREP_MULT = 4;     // CTAs of k3f_prep per persistent-CTA record (occupancy: the kernel is load-latency bound)

template <typename L>
__global__ void __launch_bounds__(256) k3f_prep(const L* __restrict__ label, int B, int H, int W, FastHier hg,
                                                Ws3 ws, int cpi, int lab_vec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = hg.nf + hg.nm + hg.nh;
  unsigned int* hist = reinterpret_cast<unsigned int*>(smem_raw);                       // [C][16]: A, M[1..12]
  unsigned int* lut = hist + (size_t)C * 16;                                            // [256] fine | mid << 8 | high << 16 | lab8 << 24
  unsigned char* rl = reinterpret_cast<unsigned char*>(lut + 256);                      // [3][PTH2+2][LPITCH]
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < C * 16; i += 256) hist[i] = 0u;
  {
    // label byte -> RMI labels of the 3 levels (void = class 0 at every level) and the uint8 label; bit 31..24 = 0xfe marks
    // a value F.one_hot would reject
    unsigned int e = 0xfe000000u;
    if (tid == SH_IGNORE) e = 0xff000000u;
    else if (tid < hg.nf) e = (unsigned)tid | ((unsigned)hg.f2m[tid] << 8) | ((unsigned)hg.f2h[tid] << 16) | ((unsigned)tid << 24);
    lut[tid] = e;
  }
  __syncthreads();
  const int cpb = cpi * PREP_MULT;
  const int b = blockIdx.x / cpb, j0 = blockIdx.x - b * cpb;
  const int tiles_x = (W + TW - 1) / TW, ntiles = tiles_x * ((H + PTH2 - 1) / PTH2);
  const long HW = (long)H * W;
  const L* lb = label + (long)b * HW;
  unsigned char* lab8 = ws.lab8 + (long)b * HW;
  const int ty = tid >> 4, tx = (tid & 15) << 2;
  const int level_base[3] = {0, hg.nf, hg.nf + hg.nm};
  unsigned int run_c[3] = {0xffu, 0xffu, 0xffu}, run_n[3] = {0u, 0u, 0u};
  unsigned int nv = 0, n_strip = 0, n_mixed = 0;
  bool bad = false;
  constexpr int NP = (TW + 4) / 2, NITEM = (PTH2 + 2) * NP, NIT = (NITEM + 255) / 256;   // pixel pairs of the label tile
#pragma unroll 1
  for (int tile = j0; tile < ntiles; tile += cpb) {
    const TileCoord tc = tile_coord(tile, tiles_x, PTH2);
    struct { long long x, y; } v[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int e = tid + 256 * it, r = e / NP, j = (e - r * NP) * 2;
      const int y = tc.y0 + r, xx = tc.x0 - 2 + j;
      v[it].x = v[it].y = -1;                              // outside the image
      if (e < NITEM && y < H && xx >= 0 && xx < W) {
        lab_ld2<L>(lb + (long)y * W + xx, lab_vec != 0, v[it].x, v[it].y);
      }
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int e = tid + 256 * it, r = e / NP, j = (e - r * NP) * 2;
      if (e >= NITEM) continue;
      const int y = tc.y0 + r, xx = tc.x0 - 2 + j;
      unsigned int f2 = 0xffffu, m2 = 0xffffu, g2 = 0xffffu;
      if (y < H && xx >= 0 && xx < W) {
        const unsigned long long t0 = (unsigned long long)v[it].x, t1 = (unsigned long long)v[it].y;
        const unsigned int e0 = t0 < 256ull ? lut[(unsigned int)t0] : 0xfe000000u;
        const unsigned int e1 = t1 < 256ull ? lut[(unsigned int)t1] : 0xfe000000u;
        bad |= (e0 >> 24) == 0xfeu || (e1 >> 24) == 0xfeu;       // F.one_hot would raise in the reference
        f2 = (e0 & 0xffu) | ((e1 & 0xffu) << 8);
        m2 = ((e0 >> 8) & 0xffu) | (e1 & 0xff00u);
        g2 = ((e0 >> 16) & 0xffu) | ((e1 >> 8) & 0xff00u);
        if (r < PTH2 && j >= 2 && j < TW + 2) {
          // out-of-range labels are stored as void (the loss is poisoned through the error flag)
          const unsigned int l0 = (e0 >> 24) == 0xfeu ? 0xffu : e0 >> 24, l1 = (e1 >> 24) == 0xfeu ? 0xffu : e1 >> 24;
          *reinterpret_cast<unsigned short*>(lab8 + (long)y * W + xx) = (unsigned short)(l0 | (l1 << 8));
          nv += (t0 != SH_IGNORE) + (t1 != SH_IGNORE);
        }
      }
      *reinterpret_cast<unsigned short*>(rl + (0 * (PTH2 + 2) + r) * LPITCH + j) = (unsigned short)f2;
      *reinterpret_cast<unsigned short*>(rl + (1 * (PTH2 + 2) + r) * LPITCH + j) = (unsigned short)m2;
      *reinterpret_cast<unsigned short*>(rl + (2 * (PTH2 + 2) + r) * LPITCH + j) = (unsigned short)g2;
    }
    __syncthreads();
#pragma unroll 1
    for (int rr = 0; rr < PTH2 / 16; ++rr) {
      const int row = ty + 16 * rr, y = tc.y0 + row;
      if (y < 2 || y >= H - 2) continue;
      unsigned int imask = 0u;       // interior anchors among the strip's 4 pixels: byte mask
#pragma unroll
      for (int k = 0; k < 4; ++k) { const int xx = tc.x0 + tx + k; if (xx >= 2 && xx < W - 2) imask |= 0xffu << (8 * k); }
      if (imask == 0u) continue;
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const unsigned char* base = rl + (l * (PTH2 + 2) + row) * LPITCH + tx;
        unsigned int lo[3], hi[3];
        unsigned int diff = 0u;
        const unsigned int c0 = base[2];
        const unsigned int pat = c0 * 0x01010101u;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const unsigned int* p = reinterpret_cast<const unsigned int*>(base + q * LPITCH);
          lo[q] = p[0]; hi[q] = p[1];
          diff |= (lo[q] ^ pat) | (hi[q] ^ pat);
        }
        if (l == 0) { ++n_strip; n_mixed += diff != 0u; }
        if (diff == 0u) {            // uniform window (and therefore all 4 pixels interior): anchors only
          if (c0 == run_c[l]) run_n[l] += 4u;
          else {
            if (run_n[l]) atomicAdd(hist + (size_t)(level_base[l] + run_c[l]) * 16, run_n[l]);
            run_c[l] = c0; run_n[l] = 4u;
          }
        } else {
          // the 4 anchors' classes as one word; tap (q, dx): neighbours = window bytes dx .. dx+3 of row q.
          // A nonzero byte of (anchors ^ neighbours) is a mismatch of that pixel and tap.
          const unsigned int cw = __funnelshift_r(lo[0], hi[0], 16);
          unsigned int hrow[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) hrow[k] = (unsigned int)(level_base[l] + ((cw >> (8 * k)) & 0xffu)) * 16u;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((imask >> (8 * k)) & 1u) atomicAdd(hist + hrow[k], 1u);
#pragma unroll
          for (int t = 1; t < 13; ++t) {
            const int q = t < 3 ? 0 : (t < 8 ? 1 : 2), sh = t < 3 ? t + 2 : (t < 8 ? t - 3 : t - 8);
            const unsigned int nw = sh == 0 ? lo[q] : (sh == 4 ? hi[q] : __funnelshift_r(lo[q], hi[q], 8 * sh));
            const unsigned int mm = (cw ^ nw) & imask;
            if (mm) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if ((mm >> (8 * k)) & 0xffu) atomicAdd(hist + hrow[k] + t, 1u);
            }
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int l = 0; l < 3; ++l)
    if (run_n[l]) atomicAdd(hist + (size_t)(level_base[l] + run_c[l]) * 16, run_n[l]);
  nv = __reduce_add_sync(0xffffffffu, nv);
  if (lane == 0 && nv) atomicAdd(ws.counts, (unsigned long long)nv);
  n_strip = __reduce_add_sync(0xffffffffu, n_strip);
  n_mixed = __reduce_add_sync(0xffffffffu, n_mixed);
  if (lane == 0 && n_strip) { atomicAdd(ws.strips + 2 * b, n_mixed); atomicAdd(ws.strips + 2 * b + 1, n_strip); }
  if (bad) atomicOr((unsigned int*)(ws.counts + 2), 1u);
  __syncthreads();
  unsigned int* out = ws.llrec + (size_t)(b * cpi + j0 % cpi) * C * 16;     // zeroed by the host before the launch
  for (int i = tid; i < C * 16; i += 256) {
    const int t = i & 15;
    const unsigned int val = t == 0 ? hist[i] : (t < 13 ? hist[i & ~15] - hist[i] : 0u);
    if (val) atomicAdd(out + i, val);
  }
}

inline size_t prep_smem(int C, int nf) {
  return (size_t)C * 16 * 4 + (size_t)256 * 4 + (size_t)3 * (PTH2 + 2) * LPITCH + 16;
}

// tile walk of a persistent CTA without integer divisions: tile index advances by cpi per step
struct TileWalk {
  int tyi, txi, tiles_x, step_y, step_x;
  __device__ __forceinline__ void init(int first, int cpi, int tx_count) {
    tiles_x = tx_count;
    tyi = first / tx_count; txi = first - tyi * tx_count;
    step_y = cpi / tx_count; step_x = cpi - step_y * tx_count;
  }
  __device__ __forceinline__ void next() {
    tyi += step_y; txi += step_x;
    if (txi >= tiles_x) { txi -= tiles_x; ++tyi; }
  }
  __device__ __forceinline__ int y0() const { return tyi * TH; }
  __device__ __forceinline__ int x0() const { return txi * TW; }
};

constexpr int PMW = 2 * 3 * 32;   // presence words per tile: [level][block] x {info, hash}

inline size_t pass1_smem(int C, int nf, int nbuf) {
  size_t s = (size_t)nbuf * NR * PLANE * 4;          // planes
  s += (size_t)XD * NPROD * 16;                      // cp.async staging of the producers
  s += (size_t)HD * NCONS * PPC * HSLOT;             // cp.async staging of the consumers' plane halos
  s += (size_t)3 * TH * TW;                          // label tile
  s += (size_t)PMW * 4;                              // presence words
  s += (size_t)C * kFastRec * 8;                     // fp64 totals
  s += (size_t)C * 8;                                // channel byte offsets
  s += (size_t)(C + 2 * nf) * 4 + 128 * 4;           // tables + reduction scratch
  s += 8 * 8 + 16;                                   // mbarriers (full / empty per round buffer)
  return (s + 15) & ~(size_t)15;
}

// 13 product taps of one 4 x BR block: anchors = centre rows (window rows 2..BR+1), taps forward.
// The row loop stays rolled (3 rows per trip, the period of the rotating register window): fully unrolled the
// consumer code overflows the instruction cache it shares with the producers.
template <bool BORDER>
__device__ __forceinline__ void pp_taps(const float* pl, unsigned int rowI, unsigned int colI, float (&acc)[16]) {
  static_assert(BR % 3 == 0, "row loop is unrolled by the window period");
  float w[3][8];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(pl + (2 + q) * PW);
    const float4 c4 = *reinterpret_cast<const float4*>(pl + (2 + q) * PW + 4);
    w[q][0] = a.x; w[q][1] = a.y; w[q][2] = a.z; w[q][3] = a.w;
    w[q][4] = c4.x; w[q][5] = c4.y; w[q][6] = c4.z; w[q][7] = c4.w;
  }
#pragma unroll 1
  for (int ib = 0; ib < BR; ib += 3) {
#pragma unroll
    for (int ii = 0; ii < 3; ++ii) {
      const int i = ib + ii;
      float(&r0)[8] = w[ii % 3];
      float(&r1)[8] = w[(ii + 1) % 3];
      float(&r2)[8] = w[(ii + 2) % 3];
      {
        const float4 a = *reinterpret_cast<const float4*>(pl + (4 + i) * PW);
        const float4 c4 = *reinterpret_cast<const float4*>(pl + (4 + i) * PW + 4);
        r2[0] = a.x; r2[1] = a.y; r2[2] = a.z; r2[3] = a.w;
        r2[4] = c4.x; r2[5] = c4.y; r2[6] = c4.z; r2[7] = c4.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float p = r0[k + 2];
        float a = p;
        if (BORDER) a = (((rowI >> (i + 2)) & 1u) && ((colI >> (k + 2)) & 1u)) ? p : 0.f;
        acc[0] = fmaf(a, p, acc[0]);
        acc[1] = fmaf(a, r0[k + 3], acc[1]);
        acc[2] = fmaf(a, r0[k + 4], acc[2]);
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
          acc[3 + dx] = fmaf(a, r1[k + dx], acc[3 + dx]);
          acc[8 + dx] = fmaf(a, r2[k + dx], acc[8 + dx]);
        }
      }
    }
  }
}

template <bool BORDER>
__device__ __forceinline__ void load_row8(const float* pl, int prow, unsigned int rowI, unsigned int colI, float (&dst)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(pl + prow * PW);
  const float4 c4 = *reinterpret_cast<const float4*>(pl + prow * PW + 4);
  dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
  dst[4] = c4.x; dst[5] = c4.y; dst[6] = c4.z; dst[7] = c4.w;
  if (BORDER) {
    const bool rok = (rowI >> prow) & 1u;
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = (rok && ((colI >> q) & 1u)) ? dst[q] : 0.f;
  }
}

// la_pr taps of one 4 x BR block, label-anchored:  lp[d] = sum_{q in block, L(q) = cl} PI(q - d).
//   uniform block of class cl: box sums (column sums over the BR rows, slid down 4 times, then 4-wide row sums);
//   a run of whole rows of class cl inside a mixed block: the same box sums over that run;
//   anything else: per matching pixel, 25 adds from the 5-row window (rolled loop, the window shifts through registers).
template <bool BORDER>
__device__ __forceinline__ void lp_taps(const float* pl, const unsigned char* ltrow, bool uniform, unsigned int pat,
                                        unsigned int rowI, unsigned int colI, float (&al)[32]) {
  // rows of the block that are entirely of this class; a single run of such rows (and no partly matching row) is a
  // uniform sub-block and takes the box sums -- the usual case on a label boundary that crosses the block horizontally
  int r0 = 0, r1 = BR;
  bool box = uniform;
  if (!uniform) {
    unsigned int full = 0u, part = 0u;
#pragma unroll
    for (int i = 0; i < BR; ++i) {
      const unsigned int z = *reinterpret_cast<const unsigned int*>(ltrow + i * TW) ^ pat;
      full |= (z == 0u ? 1u : 0u) << i;
      part |= ((z != 0u && has_zero_byte(z)) ? 1u : 0u) << i;
    }
    if (part == 0u) {
      if (full == 0u) return;                       // the class is not in the block (hash collision)
      r0 = __ffs(full) - 1;
      r1 = 32 - __clz(full);
      box = __popc(full) == r1 - r0;
    }
  }
  if (box) {
    float S[8], t[8];
    load_row8<BORDER>(pl, r0, rowI, colI, S);
#pragma unroll 1
    for (int r = r0 + 1; r < r1; ++r) {
      load_row8<BORDER>(pl, r, rowI, colI, t);
#pragma unroll
      for (int q = 0; q < 8; ++q) S[q] += t[q];
    }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      if (a > 0) {
        load_row8<BORDER>(pl, a - 1 + r0, rowI, colI, t);
#pragma unroll
        for (int q = 0; q < 8; ++q) S[q] -= t[q];
        load_row8<BORDER>(pl, a + r1 - 1, rowI, colI, t);
#pragma unroll
        for (int q = 0; q < 8; ++q) S[q] += t[q];
      }
#pragma unroll
      for (int bq = 0; bq < 5; ++bq) al[24 - (a * 5 + bq)] += (S[bq] + S[bq + 1]) + (S[bq + 2] + S[bq + 3]);
    }
  } else {
    // scattered pixels of the class (noisy labels): the 5-row window of a block row is fetched only when that row has one
#pragma unroll 1
    for (int i = 0; i < BR; ++i) {
      const unsigned int z = *reinterpret_cast<const unsigned int*>(ltrow + i * TW) ^ pat;
      if (!has_zero_byte(z)) continue;
      float w[5][8];
#pragma unroll
      for (int a = 0; a < 5; ++a) load_row8<BORDER>(pl, i + a, rowI, colI, w[a]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (((z >> (8 * k)) & 0xffu) == 0u) {
#pragma unroll
          for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int bq = 0; bq < 5; ++bq)      // window row a <-> dy = 2 - a ; col k+bq <-> dx = 2 - bq
              al[24 - (a * 5 + bq)] += w[a][k + bq];
        }
    }
  }
}

// Per-lane list of the 2-pixel ring of a plane: 2 strips of the plane rows 0,1,TH+2,TH+3 (64 strips) and up to 2 of
// the 2*PR column pairs (PR rows x {cols 0,1 | cols 66,67}).  Byte offsets into a channel plane (-1 = outside the image).
template <typename T>
__device__ __forceinline__ void halo_offsets(int lane, int ty0, int tx0, int H, int W, int e, long& vo, long& so) {
  const int hv = lane + 32 * e;
  const int vrow = (hv >> 4) < 2 ? (hv >> 4) : TH + (hv >> 4), vstrip = hv & 15;
  const int vy = ty0 - 2 + vrow, vx = tx0 + 4 * vstrip;
  vo = (vy >= 0 && vy < H && vx < W) ? ((long)vy * W + vx) * (long)sizeof(T) : -1;
  const bool hs = hv < 2 * PR;
  const int srow = hs ? hv % PR : 0, sside = hs ? hv / PR : 0;
  const int sy = ty0 - 2 + srow, sx = sside ? tx0 + TW : tx0 - 2;
  so = (hs && sy >= 0 && sy < H && sx >= 0 && sx < W) ? ((long)sy * W + sx) * (long)sizeof(T) : -1;
}

// -------------------------------------------------------------------------------------------------
// k3f_pass1
// -------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NTHREADS, 1)
k3f_pass1(const T* __restrict__ x, int B, int H, int W, FastHier hg, Ws3 ws, float eps, int cpi, int nbuf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = hg.nf + hg.nm + hg.nh;
  float* planes = reinterpret_cast<float*>(smem_raw);                                   // [NBUF][NR][PLANE]
  uint4* xstage = reinterpret_cast<uint4*>(planes + nbuf * NR * PLANE);                  // [XD][NPROD]
  unsigned char* hstage = reinterpret_cast<unsigned char*>(xstage + XD * NPROD);         // [HD][NCONS][PPC][HSLOT]
  unsigned char* LT = hstage + HD * NCONS * PPC * HSLOT;                                 // [3][TH][TW]
  unsigned int* PM = reinterpret_cast<unsigned int*>(LT + 3 * TH * TW);                  // [3][32][2]
  double* tot = reinterpret_cast<double*>(PM + PMW);                                     // [C][kFastRec]
  long long* s_chb = reinterpret_cast<long long*>(tot + (size_t)C * kFastRec);           // [C] channel byte offsets (order index)
  unsigned int* s_order = reinterpret_cast<unsigned int*>(s_chb + C);                    // [C]
  int* s_f2m = reinterpret_cast<int*>(s_order + C);                                      // [nf]
  int* s_f2h = s_f2m + hg.nf;                                                            // [nf]
  float* s_red = reinterpret_cast<float*>(s_f2h + hg.nf);                                // [128]
  unsigned long long* s_mbar = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(s_red + 128) + 7) & ~(uintptr_t)7);                  // [4] full, [4] empty
  const unsigned int mb_full = (unsigned int)__cvta_generic_to_shared(s_mbar);
  const unsigned int mb_empty = mb_full + 4 * 8;

  const int tid = threadIdx.x;
  const long HW = (long)H * W;
  for (int i = tid; i < C; i += NTHREADS) {
    const unsigned int oe = hg.order[i];
    s_order[i] = oe;
    s_chb[i] = (long long)(oe >> 24) * HW * (long long)sizeof(T);
  }
  for (int i = tid; i < hg.nf; i += NTHREADS) { s_f2m[i] = hg.f2m[i]; s_f2h[i] = hg.f2h[i]; }
  for (int i = tid; i < C * kFastRec; i += NTHREADS) tot[i] = 0.0;
  if (tid == 0) {
    for (int q = 0; q < nbuf; ++q) { mbar_init(mb_full + 8 * q, NPROD); mbar_init(mb_empty + 8 * q, NCONS); }
  }
  __syncthreads();

  const int b = blockIdx.x / cpi, j0 = blockIdx.x - b * cpi;
  const int tiles_x = (W + TW - 1) / TW, ntiles = tiles_x * ((H + TH - 1) / TH);
  const int nt = j0 < ntiles ? (ntiles - j0 + cpi - 1) / cpi : 0;     // tiles of this CTA: j0, j0+cpi, ...
  const int RPT = (C + NR - 1) / NR;
  const int total_rounds = nt * RPT;
  const unsigned char* lab8 = ws.lab8 + (long)b * HW;
  const char* xbb = reinterpret_cast<const char*>(x + (long)b * C * HW);
  const long BHW = (long)B * HW;

  if (tid < NPROD) {
    // ============================== producers ==============================
    const int ty = tid >> 4, tx = (tid & 15) << 2;
    float lacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // [0..2] log2 units (BCE fine/mid/high), [3..5] natural (CE)
    // prefetch stream: runs XD-2 channels ahead of the consumer side of the ring, across tile boundaries
    const unsigned int xs_base = (unsigned int)__cvta_generic_to_shared(xstage + tid);
    const char* xs_gen = reinterpret_cast<const char*>(xstage + tid);
    TileWalk pw;
    pw.init(j0, cpi, tiles_x);
    int pf_it = 0, pf_ci = 0;
    unsigned int pf_seq = 0;
    long pf_offb = 0;
    auto pf_tile = [&]() {
      const int y = pw.y0() + ty, xg = pw.x0() + tx;
      const bool in = pf_it < nt && y < H && xg < W;
      pf_offb = in ? ((long)y * W + xg) * (long)sizeof(T) : 0;     // outside the image: any valid address will do
    };
    auto pf_issue = [&]() {
      const char* g = xbb + s_chb[pf_ci] + pf_offb;
      const unsigned int dst = xs_base + (pf_seq & (XD - 1)) * (NPROD * 16);
      if (sizeof(T) == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g));
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(g));
      cp_async_commit();
      ++pf_seq;
      if (++pf_ci == C) { pf_ci = 0; ++pf_it; pw.next(); pf_tile(); }
    };
    pf_tile();
#pragma unroll 1
    for (int q = 0; q < XD - 2; ++q) pf_issue();

    unsigned int seq = 0;
    TileWalk tw;
    tw.init(j0, cpi, tiles_x);
#pragma unroll 1
    for (int it = 0; it < nt; ++it, tw.next()) {
      const int y = tw.y0() + ty, xg = tw.x0() + tx;
      const bool inimg = y < H && xg < W;
      const long off = (long)y * W + xg;
      const unsigned int tf4 = inimg ? *reinterpret_cast<const unsigned int*>(lab8 + off) : 0xffffffffu;
      unsigned char* hold_px = ws.hold + (long)b * HW + off;     // + plane * B*HW
      unsigned int tm4 = 0xffffffffu, th4 = 0xffffffffu;
      float vf[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned int t = (tf4 >> (8 * k)) & 0xffu;
        vf[k] = t != SH_IGNORE ? 1.f : 0.f;
        if (t != SH_IGNORE) {
          tm4 = (tm4 & ~(0xffu << (8 * k))) | ((unsigned int)s_f2m[t] << (8 * k));
          th4 = (th4 & ~(0xffu << (8 * k))) | ((unsigned int)s_f2h[t] << (8 * k));
        }
      }
      float sumF[4], sumM[4], sumH[4], prodF[4], prodM[4], prodH[4], rmax[4], rmaxH[4], a_t[4], b_t[4];
      unsigned int rhold[4], rholdH[4], hpf = 0, hpm = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        sumF[k] = sumM[k] = sumH[k] = 0.f;
        prodF[k] = prodM[k] = prodH[k] = 1.f;
        rmax[k] = rmaxH[k] = -1.f;
        a_t[k] = b_t[k] = 1.f;
        rhold[k] = rholdH[k] = 0u;
      }
      // tree BCE / CE bookkeeping of one channel (s = sigmoid, E = e^x of the thread's 4 pixels)
      auto update = [&](unsigned int oe, const float (&xv)[4], const float (&s)[4], const float (&E)[4]) {
        const int kind = oe & 3, cl = (oe >> 8) & 0xff, fl = (oe >> 16) & 0xff;
        const unsigned int ch = oe >> 24;
        const unsigned int pat = (unsigned int)cl * 0x01010101u;
        if (kind == 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            sumF[k] += E[k];
            prodF[k] *= (1.0f - s[k]) + eps;
            if (s[k] > rmax[k]) { rmax[k] = s[k]; rhold[k] = ch; }   // lowest fine id wins ties (rmi...py:386)
          }
          const unsigned int z = tf4 ^ pat;
          if (has_zero_byte(z)) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (((z >> (8 * k)) & 0xffu) == 0u) {
                a_t[k] = s[k];
                lacc[3] -= xv[k];
                lacc[0] -= lg2((1.0f - s[k]) + eps);     // the target's own factor does not belong to the product
              }
          }
          if (fl & 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { lacc[0] = fmaf(vf[k], lg2(prodF[k]), lacc[0]); prodF[k] = 1.f; }
          }
        } else if (kind == 1) {
          unsigned int hd = 0;
          float cur[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            sumM[k] += E[k];
            cur[k] = rmax[k];
            unsigned int hk = rhold[k];
            if (s[k] > cur[k]) { cur[k] = s[k]; hk = ch; }                // fine max wins ties (rmi...py:386-387)
            hd |= hk << (8 * k);
            prodM[k] *= (1.0f - cur[k]) + eps;
            if (cur[k] > rmaxH[k]) { rmaxH[k] = cur[k]; rholdH[k] = hk; }  // lower mid id wins ties (rmi...py:408)
            rmax[k] = -1.f; rhold[k] = 0u;
          }
          if (inimg) *reinterpret_cast<unsigned int*>(hold_px + (long)cl * BHW) = hd;
          const unsigned int z = tm4 ^ pat;
          if (has_zero_byte(z)) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (((z >> (8 * k)) & 0xffu) == 0u) {
                b_t[k] = s[k];
                lacc[4] -= xv[k];
                lacc[1] -= lg2((1.0f - cur[k]) + eps);
                const bool a_holds = a_t[k] <= b_t[k];                     // fine wins ties (rmi...py:421-425)
                lacc[0] += lg2((a_holds ? a_t[k] : b_t[k]) + eps);
                hpf |= (a_holds ? ((tf4 >> (8 * k)) & 0xffu) : ch) << (8 * k);
              }
          }
          if (fl & 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { lacc[1] = fmaf(vf[k], lg2(prodM[k]), lacc[1]); prodM[k] = 1.f; }
          }
        } else {
          unsigned int hd = 0;
          float cur[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            sumH[k] += E[k];
            cur[k] = rmaxH[k];
            unsigned int hk = rholdH[k];
            if (s[k] > cur[k]) { cur[k] = s[k]; hk = ch; }                // mid max wins ties (rmi...py:408-409)
            hd |= hk << (8 * k);
            prodH[k] *= (1.0f - cur[k]) + eps;
            rmaxH[k] = -1.f; rholdH[k] = 0u;
          }
          if (inimg) *reinterpret_cast<unsigned int*>(hold_px + (long)(hg.nm + cl) * BHW) = hd;
          const unsigned int z = th4 ^ pat;
          if (has_zero_byte(z)) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (((z >> (8 * k)) & 0xffu) == 0u) {
                lacc[5] -= xv[k];
                lacc[2] -= lg2((1.0f - cur[k]) + eps);
                const bool c_holds = s[k] <= b_t[k];                       // high wins ties (rmi...py:439-440)
                lacc[1] += lg2((c_holds ? s[k] : b_t[k]) + eps);
                lacc[2] += lg2(s[k] + eps);
                hpm |= (c_holds ? ch : (unsigned int)hg.nf + ((tm4 >> (8 * k)) & 0xffu)) << (8 * k);
              }
          }
          if (fl & 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { lacc[2] = fmaf(vf[k], lg2(prodH[k]), lacc[2]); prodH[k] = 1.f; }
          }
        }
      };
#pragma unroll 1
      for (int r = 0; r < RPT; ++r) {
        const int R = it * RPT + r, use = R / nbuf, buf = R - use * nbuf;
        if (use > 0) mbar_wait(mb_empty + 8 * buf, (use - 1) & 1);     // the consumers are done with the buffer's previous round
        const int cend = min(C, (r + 1) * NR);
        float* prow = planes + (buf * NR) * PLANE + (ty + 2) * PW + tx + 2;
#pragma unroll 1
        for (int ci = r * NR; ci < cend; ci += 2, prow += 2 * PLANE) {
          // two channels per step: their sigmoid chains interleave (ILP 8 instead of 4)
          const bool two = ci + 1 < cend;
          pf_issue();
          if (two) pf_issue();
          if (two) cp_async_wait<XD - 2>(); else cp_async_wait<XD - 3>();
          float xa[4], xc[4];
          staged_vec4<T>(xs_gen + (seq & (XD - 1)) * (NPROD * 16), xa);
          staged_vec4<T>(xs_gen + ((seq + 1) & (XD - 1)) * (NPROD * 16), xc);
          seq += two ? 2 : 1;
          float sa[4], Ea[4], sc[4], Ec[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) { sig_exp3(xa[k], sa[k], Ea[k]); sig_exp3(xc[k], sc[k], Ec[k]); }
          // literally probs * valid + 1e-6 (rmi...py:487)
          *reinterpret_cast<float2*>(prow) = make_float2(fmaf(sa[0], vf[0], 1e-6f), fmaf(sa[1], vf[1], 1e-6f));
          *reinterpret_cast<float2*>(prow + 2) = make_float2(fmaf(sa[2], vf[2], 1e-6f), fmaf(sa[3], vf[3], 1e-6f));
          update(s_order[ci], xa, sa, Ea);
          if (two) {
            *reinterpret_cast<float2*>(prow + PLANE) = make_float2(fmaf(sc[0], vf[0], 1e-6f), fmaf(sc[1], vf[1], 1e-6f));
            *reinterpret_cast<float2*>(prow + PLANE + 2) = make_float2(fmaf(sc[2], vf[2], 1e-6f), fmaf(sc[3], vf[3], 1e-6f));
            update(s_order[ci + 1], xc, sc, Ec);
          }
        }
        mbar_arrive(mb_full + 8 * buf);
      }
      // ---- per-pixel epilogue of the tile -------------------------------------------------------
      float iv[3][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        lacc[0] = fmaf(vf[k], lg2(prodF[k]), lacc[0]);
        lacc[1] = fmaf(vf[k], lg2(prodM[k]), lacc[1]);
        lacc[2] = fmaf(vf[k], lg2(prodH[k]), lacc[2]);
        lacc[3] = fmaf(vf[k] * kLn2, lg2(sumF[k]), lacc[3]);
        lacc[4] = fmaf(vf[k] * kLn2, lg2(sumM[k]), lacc[4]);
        lacc[5] = fmaf(vf[k] * kLn2, lg2(sumH[k]), lacc[5]);
        // 1 / sum e^x, zero on void pixels (their CE gradient is zero; the fast backward pass relies on it)
        iv[0][k] = rcp(sumF[k]) * vf[k]; iv[1][k] = rcp(sumM[k]) * vf[k]; iv[2][k] = rcp(sumH[k]) * vf[k];
      }
      if (inimg) {
        *reinterpret_cast<unsigned int*>(hold_px + (long)(hg.nm + hg.nh) * BHW) = hpf;
        *reinterpret_cast<unsigned int*>(hold_px + (long)(hg.nm + hg.nh + 1) * BHW) = hpm;
        float* ivp = ws.inv + (long)b * HW + off;
#pragma unroll
        for (int l = 0; l < 3; ++l)
          *reinterpret_cast<float4*>(ivp + (long)l * BHW) = make_float4(iv[l][0], iv[l][1], iv[l][2], iv[l][3]);
      }
    }
    cp_async_wait<0>();
    // ---- CTA totals of the six scalar sums (producer warps only) ---------------------------------
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float v = warp_sum(lacc[k]);
      if (lane == 0) s_red[k * PWARPS + warp] = v;
    }
    bar_sync(BAR_PROD, NPROD);
    if (tid < 8) {
      float r = 0.f;
      if (tid < 6) {
#pragma unroll
        for (int w = 0; w < PWARPS; ++w) r += s_red[tid * PWARPS + w];
        if (tid < 3) r *= -kLn2;
      }
      ws.bce2[(size_t)blockIdx.x * 8 + tid] = r;
    }
  } else {
    // ============================== consumers ==============================
    const int ct = tid - NPROD, cw = ct >> 5, lane = ct & 31;
    const int hb = lane >> 4, sb = lane & 15, i0 = hb * BR;
    // The 2-pixel ring of the warp's own planes (sigmoid only); logits come through a cp.async ring that runs
    // HD-1 rounds ahead (across tile boundaries).
    unsigned char* hs_gen = hstage + (size_t)ct * PPC * HSLOT;
    const unsigned int hs_base = (unsigned int)__cvta_generic_to_shared(hs_gen);
    int v_pl[2], s_pl[2];
    bool has_s[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int hv = lane + 32 * e;
      const int vrow = (hv >> 4) < 2 ? (hv >> 4) : TH + (hv >> 4);
      v_pl[e] = vrow * PW + 2 + 4 * (hv & 15);
      has_s[e] = hv < 2 * PR;
      s_pl[e] = has_s[e] ? (hv % PR) * PW + (hv / PR) * (TW + 2) : 0;
    }
    TileWalk pw;
    pw.init(j0, cpi, tiles_x);
    int pf_it = 0, pf_r = 0;
    unsigned int pf_seq = 0, seq = 0;
    long pv_offb[2] = {0, 0}, ps_offb[2] = {0, 0};
    auto pf_tile = [&]() {
      if (pf_it < nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          halo_offsets<T>(lane, pw.y0(), pw.x0(), H, W, e, pv_offb[e], ps_offb[e]);
          if (pv_offb[e] < 0) pv_offb[e] = 0;          // outside the image: any valid address will do
          if (ps_offb[e] < 0) ps_offb[e] = 0;
        }
      }
    };
    auto pf_issue = [&]() {     // one commit group per round: the halo logits of the warp's PPC planes
      if (pf_it < nt) {
        const unsigned int dst0 = hs_base + (pf_seq & (HD - 1)) * (NCONS * PPC * HSLOT);
#pragma unroll
        for (int pp = 0; pp < PPC; ++pp) {
          const int pci = pf_r * NR + cw * PPC + pp;
          if (pci < C) {
            const char* g = xbb + s_chb[pci];
            const unsigned int dst = dst0 + pp * HSLOT;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (sizeof(T) == 4) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * e), "l"(g + pv_offb[e]));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 32 + 8 * e), "l"(g + ps_offb[e]));
              } else {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 16 * e), "l"(g + pv_offb[e]));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 32 + 8 * e), "l"(g + ps_offb[e]));
              }
            }
          }
        }
      }
      cp_async_commit();
      ++pf_seq;
      if (++pf_r == RPT) { pf_r = 0; ++pf_it; pw.next(); pf_tile(); }
    };
    pf_tile();
#pragma unroll 1
    for (int q = 0; q < HD - 1; ++q) pf_issue();
    TileWalk tw;
    tw.init(j0, cpi, tiles_x);
#pragma unroll 1
    for (int it = 0; it < nt; ++it, tw.next()) {
      const int ty0 = tw.y0(), tx0 = tw.x0();
      const bool border = ty0 < 2 || ty0 + TH > H - 2 || tx0 < 2 || tx0 + TW > W - 2;
      // interior masks (only used by border tiles): plane rows i0..i0+BR+3 and window cols 0..7
      unsigned int rowI = 0, colI = 0;
      if (border) {
#pragma unroll
        for (int q = 0; q < BR + 4; ++q) { const int yy = ty0 - 2 + i0 + q; if (yy >= 2 && yy < H - 2) rowI |= 1u << q; }
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int xx = tx0 - 2 + 4 * sb + q; if (xx >= 2 && xx < W - 2) colI |= 1u << q; }
      }
      // validity of this lane's halo pixels (P = s * valid + 1e-6 inside the image, exact 0 outside)
      float vv[2][4], sv[2][2], vz[2], sz[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        long vo, so;
        halo_offsets<T>(lane, ty0, tx0, H, W, e, vo, so);
        vz[e] = vo >= 0 ? 1e-6f : 0.f;
        sz[e] = so >= 0 ? 1e-6f : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) vv[e][k] = 0.f;
        sv[e][0] = sv[e][1] = 0.f;
        if (vo >= 0) {
          const unsigned int t4 = *reinterpret_cast<const unsigned int*>(lab8 + vo / (long)sizeof(T));
#pragma unroll
          for (int k = 0; k < 4; ++k) vv[e][k] = ((t4 >> (8 * k)) & 0xffu) != SH_IGNORE ? 1.f : 0.f;
        }
        if (so >= 0) {
          const unsigned short t2 = *reinterpret_cast<const unsigned short*>(lab8 + so / (long)sizeof(T));
          sv[e][0] = (t2 & 0xffu) != SH_IGNORE ? 1.f : 0.f;
          sv[e][1] = (t2 >> 8) != SH_IGNORE ? 1.f : 0.f;
        }
      }
      // ---- label bytes of the tile per level (RMI labels: void -> class 0, outside the image -> 0xff) and the
      //      per-block presence words, built by the consumer warps together ----
      bar_sync(BAR_CONS, NCONS);                       // every consumer is done with the previous tile's labels
#pragma unroll 1
      for (int wi = ct; wi < TH * 16; wi += NCONS) {
        const int row = wi >> 4, st = wi & 15;
        const int yy = ty0 + row, xx = tx0 + 4 * st;
        unsigned int f4 = 0xffffffffu, m4 = 0xffffffffu, g4 = 0xffffffffu;
        if (yy < H && xx < W) {
          const unsigned int t4 = *reinterpret_cast<const unsigned int*>(lab8 + (long)yy * W + xx);
          f4 = m4 = g4 = 0u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const unsigned int t = (t4 >> (8 * k)) & 0xffu;
            if (t != SH_IGNORE) {
              f4 |= t << (8 * k);
              m4 |= (unsigned int)s_f2m[t] << (8 * k);
              g4 |= (unsigned int)s_f2h[t] << (8 * k);
            }
          }
        }
        *reinterpret_cast<unsigned int*>(LT + (0 * TH + row) * TW + 4 * st) = f4;
        *reinterpret_cast<unsigned int*>(LT + (1 * TH + row) * TW + 4 * st) = m4;
        *reinterpret_cast<unsigned int*>(LT + (2 * TH + row) * TW + 4 * st) = g4;
      }
      bar_sync(BAR_CONS, NCONS);
      // per (level, 4 x BR block): info = class of the first pixel | uniform << 8 ; hash = classes present (bit c & 31)
      if (ct < 96) {
        const int l = ct >> 5, blk = ct & 31, bi0 = (blk >> 4) * BR, bsb = blk & 15;
        const unsigned int first = *reinterpret_cast<const unsigned int*>(LT + (l * TH + bi0) * TW + 4 * bsb);
        const unsigned int pat0 = (first & 0xffu) * 0x01010101u;
        unsigned int diff = 0, hash = 0;
#pragma unroll
        for (int i = 0; i < BR; ++i) {
          const unsigned int wd = *reinterpret_cast<const unsigned int*>(LT + (l * TH + bi0 + i) * TW + 4 * bsb);
          diff |= wd ^ pat0;
          hash |= (1u << (wd & 31)) | (1u << ((wd >> 8) & 31)) | (1u << ((wd >> 16) & 31)) | (1u << ((wd >> 24) & 31));
        }
        PM[2 * ct] = (first & 0xffu) | (diff == 0u ? 0x100u : 0u);
        PM[2 * ct + 1] = hash;
      }
      bar_sync(BAR_CONS, NCONS);
      unsigned int pinfo[3], phash[3];
#pragma unroll
      for (int l = 0; l < 3; ++l) { pinfo[l] = PM[l * 64 + 2 * lane]; phash[l] = PM[l * 64 + 2 * lane + 1]; }
      const unsigned char* lt = LT + i0 * TW + 4 * sb;
#pragma unroll 1
      for (int r = 0; r < RPT; ++r) {
        const int R = it * RPT + r, use = R / nbuf, buf = R - use * nbuf;
        pf_issue();
        cp_async_wait<HD - 1>();
        // halos of this warp's planes: independent of the producers, done while they finish the round
        const unsigned char* st0 = hs_gen + (seq & (HD - 1)) * (NCONS * PPC * HSLOT);
        ++seq;
#pragma unroll 1
        for (int pp = 0; pp < PPC; ++pp) {
          const int ci = r * NR + cw * PPC + pp;
          if (ci >= C) break;
          const unsigned char* st = st0 + pp * HSLOT;
          float* plw = planes + (buf * NR + cw * PPC + pp) * PLANE;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float xv[4], xs[2];
            staged_vec4<T>(st + 16 * e, xv);
            if (sizeof(T) == 4) {
              const float2 t2 = *reinterpret_cast<const float2*>(st + 32 + 8 * e);
              xs[0] = t2.x; xs[1] = t2.y;
            } else {
              xs[0] = staged_elem<T>(st + 32 + 8 * e, 0);
              xs[1] = staged_elem<T>(st + 32 + 8 * e, 1);
            }
            float* pr = plw + v_pl[e];
            *reinterpret_cast<float2*>(pr) = make_float2(fmaf(sig_only(xv[0]), vv[e][0], vz[e]), fmaf(sig_only(xv[1]), vv[e][1], vz[e]));
            *reinterpret_cast<float2*>(pr + 2) = make_float2(fmaf(sig_only(xv[2]), vv[e][2], vz[e]), fmaf(sig_only(xv[3]), vv[e][3], vz[e]));
            if (has_s[e])
              *reinterpret_cast<float2*>(plw + s_pl[e]) =
                  make_float2(fmaf(sig_only(xs[0]), sv[e][0], sz[e]), fmaf(sig_only(xs[1]), sv[e][1], sz[e]));
          }
        }
        mbar_wait(mb_full + 8 * buf, use & 1);                         // the producers have filled this round's planes
#pragma unroll 1
        for (int pp = 0; pp < PPC; ++pp) {
          const int ci = r * NR + cw * PPC + pp;
          if (ci >= C) break;
          const unsigned int oe = s_order[ci];
          const int lvl = oe & 3, cl = (oe >> 8) & 0xff;
          const unsigned int ch = oe >> 24;
          const float* pl = planes + (buf * NR + cw * PPC + pp) * PLANE + i0 * PW + 4 * sb;   // window row 0 = plane row i0
          double* trow = tot + (size_t)ch * kFastRec;
          // ---- pr_cov: 13 product taps of the interior anchors ----
          {
            float acc[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = 0.f;
            if (border) pp_taps<true>(pl, rowI, colI, acc); else pp_taps<false>(pl, rowI, colI, acc);
            const float t = warp_reduce16(acc, lane);
            if ((lane & 1) == 0) {
              const int slot = reduce16_slot(lane);
              if (slot < 13) trow[slot] += (double)t;
            }
          }
          // ---- la_pr: label-anchored taps  lp[d] = sum_{q : L(q) = cl} PI(q - d),  PI = P on interior anchors, else 0 ----
          const unsigned int info = lvl == 0 ? pinfo[0] : (lvl == 1 ? pinfo[1] : pinfo[2]);
          const unsigned int hash = lvl == 0 ? phash[0] : (lvl == 1 ? phash[1] : phash[2]);
          const bool uniform = (info & 0x100u) != 0u;
          const bool hit = uniform ? (info & 0xffu) == (unsigned int)cl : ((hash >> (cl & 31)) & 1u) != 0u;
          if (__any_sync(0xffffffffu, hit)) {
            float al[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) al[q] = 0.f;
            if (hit) {
              const unsigned int pat = (unsigned int)cl * 0x01010101u;
              const unsigned char* ltrow = lt + lvl * TH * TW;
              if (border) lp_taps<true>(pl, ltrow, uniform, pat, rowI, colI, al);
              else lp_taps<false>(pl, ltrow, uniform, pat, rowI, colI, al);
            }
            const float t0 = warp_reduce16(*reinterpret_cast<float(*)[16]>(al), lane);
            const float t1 = warp_reduce16(*reinterpret_cast<float(*)[16]>(al + 16), lane);
            if ((lane & 1) == 0) {
              const int slot = reduce16_slot(lane);
              trow[13 + slot] += (double)t0;
              if (slot < 9) trow[29 + slot] += (double)t1;
            }
          }
          __syncwarp();
        }
        mbar_arrive(mb_empty + 8 * buf);
      }
    }
    cp_async_wait<0>();
    // ---- records of the channels this warp owned ------------------------------------------------
    __syncwarp();
    for (int ci = 0; ci < C; ++ci) {
      if (((ci % NR) / PPC) != cw) continue;
      const unsigned int ch = s_order[ci] >> 24;
      double* rec = ws.rec2 + ((size_t)blockIdx.x * C + ch) * kFastRec;
      for (int q = lane; q < kFastRec; q += 32) rec[q] = tot[(size_t)ch * kFastRec + q];
    }
  }
}

}  // namespace fast
}  // namespace sh
