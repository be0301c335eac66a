// Shared definitions for the three-level (BCE + RMI + CE) kernels.
//
// RMI decomposition (mirrors oracle/rmi_taps.py, which is checked on the CPU
// against the reference's unfold definition, rmi_hiera_triplet_loss.py:292-311,
// 493-517): with the P-side pixel r as anchor and d = d_i - d_j,
//     S_xy[i,j] = sum_{r in R_j} Y[r] X[r+d]
// Pixels are split by (row class, col class), class = which window rows/cols j are
// valid for them: 0,1 = first two rows, 2 = middle, 3,4 = last two rows.  The
// (2,2) class (the interior) is accumulated by the streaming kernels as 25/13 "taps";
// the 24 border classes are accumulated by the small frame kernels.
#pragma once
#include "common.cuh"

namespace sh {

constexpr int kTW = 64;           // tile width (pixels); one thread owns 4 consecutive pixels
constexpr int kStrips = kTW / 4;  // 16 strips per tile row
constexpr int kPitch = kTW + 4;   // plane pitch: cols x0-2 .. x0+TW+1
constexpr int kLabPitch = kTW + 8;  // label tile pitch (bytes), cols x0-2 .. (8-byte windows stay in bounds)
constexpr int kNPart = 56;        // per (tile, channel) partial: see PartIdx
constexpr int kNTap = 25;
constexpr int kNHalf = 13;

// layout of one per-(tile,channel) partial record (floats)
enum PartIdx {
  kPP = 0,        // [13] interior taps of P*P, half plane
  kLPFull = 13,   // sum of P over label-uniform interior anchors of this class (all 25 lp taps)
  kLLFull = 14,   // count of label-uniform interior anchors of this class   (all ll taps)
  kLPS = 15,      // [25] lp taps from non-uniform interior anchors
  kLLS = 40,      // [13] ll taps from non-uniform interior anchors (half plane)
};

// flags byte per pixel (from k3_prep)
constexpr int kFlagInterior = 1;
constexpr int kFlagUniF = 2;
constexpr int kFlagUniM = 4;
constexpr int kFlagUniH = 8;

struct Hier3 {
  int nf, nm, nh;
  const int* f2m;             // [nf]
  const int* f2h;             // [nf]
  const int* mh_ptr;          // [nm+1]  CSR: highs h with m in Ms(h)
  const int* mh_idx;
  const unsigned int* hsmask; // [nm] bit h set iff h in Hs(m)
};

__host__ __device__ inline int half_tap_index(int dy, int dx) {
  // order: (0,0),(0,1),(0,2),(1,-2..2),(2,-2..2)
  return dy == 0 ? dx : 3 + (dy - 1) * 5 + (dx + 2);
}
__host__ __device__ inline int tap_index(int dy, int dx) { return (dy + 2) * 5 + (dx + 2); }
__host__ __device__ inline bool in_half_plane(int dy, int dx) { return dy > 0 || (dy == 0 && dx >= 0); }

// row/col class of coordinate v in an axis of length n (n >= 5)
__host__ __device__ inline int axis_class(int v, int n) { return v < 2 ? v : (v >= n - 2 ? 3 + (v - (n - 2)) : 2); }
// is window offset o (0..2) valid for class k?   J0={0} J1={0,1} J2={0,1,2} J3={1,2} J4={2}
__host__ __device__ inline bool offset_valid(int k, int o) {
  const int lo = k <= 2 ? 0 : k - 2, hi = k >= 2 ? 2 : k;
  return o >= lo && o <= hi;
}

struct Ws3 {
  unsigned long long* counts;  // [4]: 0 = #valid, 2 = error flag
  unsigned char* lab8;         // [B*HW]
  unsigned char* flags;        // [B*HW]
  unsigned char* hold;         // [(nm+nh+2)][B*HW]
  float* inv;                  // [3][B*HW]   1/sum_c e^x per level
  float* part1;                // [B*ntiles][C][kNPart]
  float* bcepart;              // [B*ntiles][8]
  double* sums;                // [8]
  float* frameT;               // [nseg][B*C][25 classes][75]
  double* rbc;                 // [B*C]
  float* wts;                  // [B*C][64]: W1[25], W2[25], W2full at 50
  float* fwts;                 // [B*C][25 classes][50]
  size_t bytes;
  int tiles_x, tiles_y, th, nseg;
};

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Tile height: as tall as the per-pixel streaming state of pass 1 allows (see k3_pass1).
inline int pick_tile_rows(int nm, int nh) {
  for (int th = 16; th >= 2; th >>= 1) {
    size_t px = (size_t)th * kTW;
    size_t state = (size_t)(nm + nh) * px * 5;
    if (state <= 96 * 1024) return th;
  }
  return 2;
}

// frame runs are split into segments so that long image edges spread over more CTAs
inline int frame_segments(int H, int W) {
  int n = ((H > W ? H : W) + 511) / 512;
  return n < 1 ? 1 : (n > 8 ? 8 : n);
}

inline Ws3 ws3_layout(void* base, int B, int H, int W, int nf, int nm, int nh) {
  Ws3 w;
  const size_t n = (size_t)B * H * W;
  const int C = nf + nm + nh;
  w.th = pick_tile_rows(nm, nh);
  w.tiles_x = (W + kTW - 1) / kTW;
  w.tiles_y = (H + w.th - 1) / w.th;
  const size_t ntiles = (size_t)w.tiles_x * w.tiles_y * B;
  size_t off = 0;
  unsigned char* p = (unsigned char*)base;
  auto take = [&](size_t bytes) { size_t o = off; off = align256(off + bytes); return p + o; };
  w.counts = (unsigned long long*)take(4 * 8);
  w.sums = (double*)take(8 * 8);
  w.lab8 = take(n);
  w.flags = take(n);
  w.hold = take((size_t)(nm + nh + 2) * n);
  w.inv = (float*)take(3 * n * 4);
  w.part1 = (float*)take(ntiles * C * kNPart * 4);
  w.bcepart = (float*)take(ntiles * 8 * 4);
  w.nseg = frame_segments(H, W);
  w.frameT = (float*)take((size_t)w.nseg * B * C * 25 * 75 * 4);
  w.rbc = (double*)take((size_t)B * C * 8);
  w.wts = (float*)take((size_t)B * C * 64 * 4);
  w.fwts = (float*)take((size_t)B * C * 25 * 50 * 4);
  w.bytes = off;
  return w;
}

inline Hier3 hier3_from_tab(const int* tab, int nf, int nm, int nh, int n_mh) {
  Hier3 h;
  h.nf = nf; h.nm = nm; h.nh = nh;
  h.f2m = tab;
  h.f2h = h.f2m + nf;
  h.mh_ptr = h.f2h + nf;
  h.mh_idx = h.mh_ptr + nm + 1;
  h.hsmask = (const unsigned int*)(h.mh_idx + n_mh);
  return h;
}

// byte k of a 64-bit little-endian window
__device__ __forceinline__ unsigned int byte_of(unsigned long long w, int k) { return (unsigned int)(w >> (8 * k)) & 0xffu; }

}  // namespace sh
