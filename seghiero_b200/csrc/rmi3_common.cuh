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
//
// Precision: the reference accumulates the moments in float64.  Here every tap is
// split into a large common part and a small difference part,
//     P[r] P[r+d] = P[r]^2 + P[r] (P[r+d] - P[r]),
// the common parts (sum P^2, sum P over a class region) are carried in float64 from
// the warp level upwards, the difference parts in float32 (their rounding error
// scales with the LOCAL variation of P, which is what the small eigenvalues of the
// 9x9 matrices are made of).  All 9x9 algebra is float64.
#pragma once
#include "common.cuh"

namespace sh {

constexpr int kTW = 64;             // tile width (pixels)
constexpr int kTH = 16;             // tile height of the streaming kernels
constexpr int kStrips = kTW / 4;    // 16 four-pixel strips per tile row
constexpr int kPitch = kTW + 4;     // plane pitch: cols x0-2 .. x0+TW+1
constexpr int kLabPitch = kTW + 8;  // label tile pitch (bytes); 8-byte windows stay in bounds
constexpr int kNR = 4;              // channels per round
constexpr int kThreads = 16 * kTH;  // CTA size of k3_pass1 / k3_pass2 (one thread per 4-pixel strip)
constexpr int kGroup = kThreads / kNR;      // threads that share one channel plane in phase B of pass 1
constexpr int kGroupWarps = kGroup / 32;
constexpr int kRec = 64;            // floats per (tile, channel) partial record

// layout of one per-(tile,channel) partial record (float slots)
enum RecIdx {
  kD = 0,        // [12] difference taps  sum a P[r] (P[r+d] - P[r]),  d in half plane \ {0}  (order: half_tap_index-1)
  kLLFull = 12,  // count of label-uniform interior anchors of this class
  kT0 = 16,      // double: sum a P[r]^2
  kLPFull = 18,  // double: sum of P over label-uniform interior anchors of this class
  kLPS = 20,     // [25] lp taps from non-uniform interior anchors
  kLLS = 45,     // [13] ll taps from non-uniform interior anchors (half plane)
};

// per-class record of the frame kernels (float slots, kFrameRec per class)
constexpr int kFrameRec = 80;
enum FrameIdx {
  kFT0 = 0,     // double: sum P[r]^2 over the class
  kFLP0 = 2,    // double: sum P[r] L[r]
  kFD = 4,      // [25] sum P[r] (P[r+d] - P[r])
  kFDL = 29,    // [25] sum P[r] (L[r+d] - L[r])
  kFLL = 54,    // [25] sum L[r] L[r+d]
};

// flags byte per pixel (from k3_prep)
constexpr int kFlagInterior = 1;
constexpr int kFlagUniF = 2;

// channel processing order entry (pass 1): kind 0/1/2 = fine/mid/high
struct OrderEntry {
  unsigned char kind, cl, flags, pad;   // flags: bit0 reset running max, bit1 flush the level's log product
};

struct Hier3 {
  int nf, nm, nh;
  const int* f2m;             // [nf]
  const int* f2h;             // [nf]
  const int* mh_ptr;          // [nm+1]  CSR: highs h with m in Ms(h)
  const int* mh_idx;
  const unsigned int* hsmask; // [nm] bit h set iff h in Hs(m)
  const unsigned int* order;  // [C] packed OrderEntry: kind | cl<<8 | flags<<16
  const int* tab;             // the whole int32 blob the pointers above point into
  int n_mh, tab_len;
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
  float* part1;                // [B*ntiles][C][kRec]
  float* bcepart;              // [B*ntiles][8]
  double* sums;                // [8]
  float* frameT;               // [nseg][B*C][25 classes][kFrameRec]
  double* rbc;                 // [B*C]
  float* wts;                  // [B*C][64]: W1[25], W2[25], W2full at 50
  float* fwts;                 // [B*C][25 classes][50]
  // fast path (rmi3_fast.cuh): per-(persistent CTA, channel) records, cpi CTAs per image
  double* rec2;                // [B*cpi][C][kFastRec]: pp product taps [13], lp taps [25]
  unsigned int* llrec;         // [B*cpi][C][16]: label-label half-plane counts [13]
  float* bce2;                 // [B*cpi][8]
  unsigned int* strips;        // [B][2]: label strips with a mixed 3x8 window (fine level), all strips (k3f_prep)
  unsigned char* labB;         // [B][3 levels][8 W + 8 H]: RMI labels (void = 0) of the 4-pixel border bands, laid out
                               // like bandR / bandC (k3_band writes them, the frame kernels stage them)
  size_t bytes;
  int tiles_x, tiles_y, nseg, cpi;
};

constexpr int kFastRec = 40;   // doubles per (CTA, channel) record of the fast forward pass
// persistent CTAs per image of the fast path: one CTA per SM over the whole batch
inline int fast_ctas_per_image(int B) { int c = SH_NUM_SMS / (B > 0 ? B : 1); return c < 1 ? 1 : c; }

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// frame runs are split into segments so that long image edges spread over more CTAs
constexpr int kSegMax = 256;        // pixels of a band segment staged in shared memory at a time
inline int frame_segments(int H, int W) {
  int n = ((H > W ? H : W) + kSegMax - 1) / kSegMax;
  return n < 1 ? 1 : (n > 32 ? 32 : n);
}

// A segment of one of the four 4-pixel border bands, staged in shared memory: P and the indicator of the
// channel's class (1.0 where the RMI label equals cl).  Band coordinates: u along the image edge, v = 0..3 across
// it; side 0 = rows 0..3, 1 = rows H-4..H-1, 2 = cols 0..3, 3 = cols W-4..W-1.
struct BandSeg {
  float P[4][kSegMax + 4];
  float L[4][kSegMax + 4];
};
// Returns whether this thread staged a pixel of class cl (callers skip the label taps of segments without one).
// `lb` = the (image, level) block of Ws3::labB.
__device__ __forceinline__ bool stage_band(BandSeg& s, int side, int u0, int n, const float* bandR, const float* bandC,
                                           const unsigned char* lb, int cl, int H, int W, int tid, int nthreads) {
  const int N = side < 2 ? W : H;
  const float* pb = side < 2 ? bandR : bandC;
  const unsigned char* lbb = side < 2 ? lb : lb + 8 * W;
  bool saw = false;
#pragma unroll 4
  for (int e = tid; e < 4 * (n + 4); e += nthreads) {
    const int v = e / (n + 4), i = e - v * (n + 4), u = u0 - 2 + i;
    float p = 0.f, l = 0.f;
    if (u >= 0 && u < N) {
      const size_t at = (size_t)((side & 1) ? 4 + v : v) * N + u;
      p = pb[at];
      const bool is = lbb[at] == cl;
      l = is ? 1.f : 0.f;
      saw |= is;
    }
    s.P[v][i] = p;
    s.L[v][i] = l;
  }
  return saw;
}

inline Ws3 ws3_layout(void* base, int B, int H, int W, int nf, int nm, int nh) {
  Ws3 w;
  const size_t n = (size_t)B * H * W;
  const int C = nf + nm + nh;
  w.tiles_x = (W + kTW - 1) / kTW;
  w.tiles_y = (H + kTH - 1) / kTH;
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
  w.part1 = (float*)take(ntiles * C * kRec * 4);
  w.bcepart = (float*)take(ntiles * 8 * 4);
  w.nseg = frame_segments(H, W);
  w.frameT = (float*)take((size_t)w.nseg * B * C * 25 * kFrameRec * 4);
  w.rbc = (double*)take((size_t)B * C * 8);
  w.wts = (float*)take((size_t)B * C * 64 * 4);
  w.fwts = (float*)take((size_t)B * C * 25 * 50 * 4);
  w.cpi = fast_ctas_per_image(B);
  w.rec2 = (double*)take((size_t)B * w.cpi * C * kFastRec * 8);
  w.llrec = (unsigned int*)take((size_t)B * w.cpi * C * 16 * 4);
  w.bce2 = (float*)take((size_t)B * w.cpi * 8 * 4);
  w.strips = (unsigned int*)take((size_t)B * 2 * 4);
  w.labB = (unsigned char*)take((size_t)B * 3 * 8 * ((size_t)W + H));
  w.bytes = off;
  return w;
}

// hier_tab (device int32): [f2m nf][f2h nf][mh_ptr nm+1][mh_idx n_mh][hsmask nm][order C]
__host__ __device__ inline Hier3 hier3_from_tab(const int* tab, int nf, int nm, int nh, int n_mh) {
  Hier3 h;
  h.nf = nf; h.nm = nm; h.nh = nh;
  h.tab = tab; h.n_mh = n_mh;
  h.tab_len = 2 * nf + (nm + 1) + n_mh + nm + (nf + nm + nh);
  h.f2m = tab;
  h.f2h = h.f2m + nf;
  h.mh_ptr = h.f2h + nf;
  h.mh_idx = h.mh_ptr + nm + 1;
  h.hsmask = (const unsigned int*)(h.mh_idx + n_mh);
  h.order = h.hsmask + nm;
  return h;
}

// byte k of a 64-bit little-endian window
__device__ __forceinline__ unsigned int byte_of(unsigned long long w, int k) { return (unsigned int)(w >> (8 * k)) & 0xffu; }

__device__ __forceinline__ unsigned int load4_u8(const unsigned char* p, int nvalid, bool aligned) {
  if (aligned && nvalid == 4) return *reinterpret_cast<const unsigned int*>(p);
  unsigned int r = 0;
  for (int k = 0; k < nvalid; ++k) r |= (unsigned int)p[k] << (8 * k);
  return r;
}
__device__ __forceinline__ void store4_u8(unsigned char* p, unsigned int v, int nvalid, bool aligned) {
  if (aligned && nvalid == 4) {
    *reinterpret_cast<unsigned int*>(p) = v;
  } else {
    for (int k = 0; k < nvalid; ++k) p[k] = (unsigned char)(v >> (8 * k));
  }
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

// P and RMI label of a pixel that is known to lie in a 4-pixel border band (see k3_band)
struct BandView {
  const float* bandR; const float* bandC; const unsigned char* lab8; const int* lmap; int H, W;
  __device__ __forceinline__ float P(int yy, int xx) const {
    if (yy < 4) return bandR[(size_t)yy * W + xx];
    if (yy >= H - 4) return bandR[(size_t)(yy - (H - 8)) * W + xx];
    if (xx < 4) return bandC[(size_t)xx * H + yy];
    return bandC[(size_t)(xx - (W - 8)) * H + yy];
  }
  __device__ __forceinline__ int L(int yy, int xx) const {
    const int t = lab8[(long)yy * W + xx];
    return t == SH_IGNORE ? 0 : (lmap ? lmap[t] : t);
  }
};

// Copy the hierarchy tables into shared memory (they are read once per channel by every thread) and
// return a view that points into the copy.  Caller must __syncthreads() before using it.
__device__ __forceinline__ Hier3 stage_hier(const Hier3& h, int* dst, int tid, int nthreads) {
  for (int i = tid; i < h.tab_len; i += nthreads) dst[i] = h.tab[i];
  return hier3_from_tab(dst, h.nf, h.nm, h.nh, h.n_mh);
}

// label tile (3 levels, halo 2) from the uint8 labels: rows y0-2 .. y0+rows+1, cols x0-2 .. x0+65
__device__ __forceinline__ void load_label_tile(unsigned char* labt, int rows, const unsigned char* lab8, int H, int W,
                                                int y0, int x0, const Hier3& h, int tid, int nthreads) {
  for (int e = tid; e < (rows + 4) * kPitch; e += nthreads) {
    const int r = e / kPitch, j = e - r * kPitch;
    const int yy = y0 - 2 + r, xx = x0 - 2 + j;
    unsigned char f = 0xff, m = 0xff, g = 0xff;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const int t = lab8[(long)yy * W + xx];
      f = m = g = 0;   // void pixels are one-hot of class 0 at every level inside RMI
      if (t != SH_IGNORE) { f = (unsigned char)t; m = (unsigned char)h.f2m[t]; g = (unsigned char)h.f2h[t]; }
    }
    labt[(0 * (rows + 4) + r) * kLabPitch + j] = f;
    labt[(1 * (rows + 4) + r) * kLabPitch + j] = m;
    labt[(2 * (rows + 4) + r) * kLabPitch + j] = g;
  }
}

}  // namespace sh
