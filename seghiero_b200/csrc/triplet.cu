// TreeTripletLoss: deterministic anchor/positive/negative selection ("first <= max_triplet rows in
// raster order") and PAIRED row dots -- there is no pairwise distance matrix in the reference
// (models/loss/tree_triplet_loss.py:15-65, models/loss/rmi_tree_triplet_loss.py:14-70), so no GEMM here.
// The work is tiny (<= ncls*3*200 rows of D floats) and latency bound: no host syncs, four launches.
//   mode 0: hierarchy flavour, pos = same bucket \ {c}, neg = outside bucket (255 is a negative)
//   mode 1: id-list flavour,  pos = same list \ {c},   neg = other list; classes 0 and 255 skipped
#include "common.cuh"

namespace sh {

__device__ __forceinline__ int nearest_src(int dst, float scale, int n_in) {
  const int s = (int)floorf((float)dst * scale);   // F.interpolate(mode='nearest') source index
  return s < n_in - 1 ? s : n_in - 1;
}

template <typename L>
__global__ void __launch_bounds__(256) k_trip_labels(const L* __restrict__ label, int B, int H, int W, int h,
                                                     int w, int mode, const int* __restrict__ tab, int ncls,
                                                     int* __restrict__ lab_ds, int* __restrict__ status) {
  const float sy = (float)H / (float)h, sx = (float)W / (float)w;
  const long R = (long)B * h * w;
  bool bad = false;
  for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < R; r += (long)gridDim.x * blockDim.x) {
    const int b = (int)(r / (h * w)), rem = (int)(r - (long)b * h * w);
    const int i = rem / w, j = rem - i * w;
    const long long t = lab_ld(label, ((long)b * H + nearest_src(i, sy, H)) * W + nearest_src(j, sx, W));
    int v = (t >= 0 && t < 0x7fffffff) ? (int)t : -1;
    if (mode == 0) { if (v != SH_IGNORE && (v < 0 || v >= ncls)) bad = true; }
    else { if (v != SH_IGNORE && v != 0 && (v < 0 || v >= 256 || tab[v] < 0)) bad = true; }
    lab_ds[r] = v;
  }
  if (bad) atomicOr(status + 1, 1);
}

__device__ __forceinline__ void classify(int lab, int c, int mode, int lo, int hi, int grp_c,
                                         const int* __restrict__ tab, bool& fa, bool& fp, bool& fn) {
  fa = lab == c;
  if (mode == 0) {
    const bool in = lab >= lo && lab < hi;
    fp = in && !fa;
    fn = !in;
  } else {
    const int g = (lab >= 0 && lab < 256) ? tab[lab] : -1;
    fp = g >= 0 && g == grp_c && !fa;
    fn = g >= 0 && g == 1 - grp_c;
  }
}

// one CTA per class: ordered compaction of the first `maxT` anchors / positives / negatives.
// Rows are scanned 8192 at a time: warp w owns rows [r0 + 1024 w, r0 + 1024 (w + 1)) in raster order (32 steps of 32
// consecutive rows), counts its three lists with ballots, the eight warp totals are prefix-summed once, and a second
// walk over the same (cached) labels writes the row indices: two barriers per 8192 rows.
__global__ void __launch_bounds__(256) k_trip_select(const int* __restrict__ lab_ds, long R, int mode,
                                                     const int* __restrict__ tab, int ncls, int maxT,
                                                     int* __restrict__ sel, int* __restrict__ kcount) {
  constexpr int STEPS = 32, WROWS = 32 * STEPS, CHUNK = 8 * WROWS;
  __shared__ int wtot[3][8];
  __shared__ int base[3];
  const int c = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int lo = 0, hi = 0, grp_c = -1;
  bool skip = false;
  if (mode == 0) { lo = tab[c]; hi = tab[ncls + c]; }
  else { grp_c = tab[c]; skip = (c == 0 || c == SH_IGNORE || grp_c < 0); }
  if (c == SH_IGNORE) skip = true;
  if (skip) { if (threadIdx.x == 0) kcount[c] = 0; return; }
  if (threadIdx.x < 3) base[threadIdx.x] = 0;
  __syncthreads();
  int* out = sel + (size_t)c * 3 * maxT;
  const unsigned int lt_mask = (1u << lane) - 1u;
  for (long r0 = 0; r0 < R; r0 += CHUNK) {
    const long wbase = r0 + (long)warp * WROWS;
    int cnt[3] = {0, 0, 0};
    for (int t = 0; t < STEPS; ++t) {
      const long r = wbase + t * 32 + lane;
      bool f[3] = {false, false, false};
      if (r < R) classify(lab_ds[r], c, mode, lo, hi, grp_c, tab, f[0], f[1], f[2]);
#pragma unroll
      for (int s = 0; s < 3; ++s) cnt[s] += __popc(__ballot_sync(0xffffffffu, f[s]));
    }
    if (lane < 3) wtot[lane][warp] = lane == 0 ? cnt[0] : (lane == 1 ? cnt[1] : cnt[2]);
    __syncthreads();
    int off[3], tot[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      off[s] = base[s];
      tot[s] = 0;
#pragma unroll
      for (int w2 = 0; w2 < 8; ++w2) {
        const int n = wtot[s][w2];
        if (w2 < warp) off[s] += n;
        tot[s] += n;
      }
    }
    // second walk: positions (only while the list still has room)
    if (off[0] < maxT || off[1] < maxT || off[2] < maxT) {
      for (int t = 0; t < STEPS; ++t) {
        const long r = wbase + t * 32 + lane;
        bool f[3] = {false, false, false};
        if (r < R) classify(lab_ds[r], c, mode, lo, hi, grp_c, tab, f[0], f[1], f[2]);
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const unsigned int bal = __ballot_sync(0xffffffffu, f[s]);
          const int pos = off[s] + __popc(bal & lt_mask);
          if (f[s] && pos < maxT) out[s * maxT + pos] = (int)r;
          off[s] += __popc(bal);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < 3) base[threadIdx.x] += threadIdx.x == 0 ? tot[0] : (threadIdx.x == 1 ? tot[1] : tot[2]);
    __syncthreads();
    if (base[0] >= maxT && base[1] >= maxT && base[2] >= maxT) break;
  }
  if (threadIdx.x == 0) kcount[c] = min(min(base[0], base[1]), min(base[2], maxT));
}

template <typename T>
__device__ __forceinline__ float feat_at(const T* __restrict__ f, int D, long hw, long r, int d) {
  const long b = r / hw, p = r - b * hw;
  return to_f32<T>(f[(b * D + d) * hw + p]);
}

// grid (ncls, ceil(maxT/8)), block 256: one warp per triplet
template <typename T>
__global__ void __launch_bounds__(256) k_trip_hinge(const T* __restrict__ feats, int D, long hw, int maxT,
                                                    const int* __restrict__ sel, const int* __restrict__ kcount,
                                                    float* __restrict__ tl) {
  const int c = blockIdx.x, t = blockIdx.y * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (t >= kcount[c]) return;
  const int* s = sel + (size_t)c * 3 * maxT;
  const long ra = s[t], rp = s[maxT + t], rn = s[2 * maxT + t];
  float ap = 0.f, an = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float a = feat_at<T>(feats, D, hw, ra, d);
    ap = fmaf(a, feat_at<T>(feats, D, hw, rp, d), ap);
    an = fmaf(a, feat_at<T>(feats, D, hw, rn, d), an);
  }
  ap = warp_sum(ap);
  an = warp_sum(an);
  if (lane == 0) tl[(size_t)c * maxT + t] = fmaxf((1.0f - ap) - (1.0f - an) + 0.6f, 0.f);
}

// trip[0] = mean over contributing classes of the per-class mean hinge, trip[1] = #classes; status[0] = ready
__global__ void __launch_bounds__(256) k_trip_reduce(int ncls, int maxT, const int* __restrict__ kcount,
                                                     const float* __restrict__ tl, float* __restrict__ trip,
                                                     int* __restrict__ status) {
  __shared__ float csum[256];
  __shared__ int ccnt[256];
  float s = 0.f;
  int n = 0;
  for (int c = threadIdx.x; c < ncls; c += 256) {
    const int k = kcount[c];
    if (k > 0) {
      float a = 0.f;
      for (int t = 0; t < k; ++t) a += tl[(size_t)c * maxT + t];
      s += a / (float)k;
      n++;
    }
  }
  csum[threadIdx.x] = s; ccnt[threadIdx.x] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f; int cnt = 0;
    for (int i = 0; i < 256; ++i) { tot += csum[i]; cnt += ccnt[i]; }
    trip[0] = cnt > 0 ? tot / (float)cnt : 0.f;
    trip[1] = (float)cnt;
    status[0] = cnt > 0 ? 1 : 0;
  }
}

// grad wrt feats, one warp per active triplet, scattered with INTEGER atomics: every contribution is quantised to
// fixed point (2^-40 of the unit "feature / (k_c * #classes)", i.e. ~4e-9 relative to a typical term) before it is
// added, so the sum does not depend on the order the atomics land in and embedding.grad is bitwise reproducible (a
// float-atomic scatter is not; a barrier-ordered plain read-modify-write walk was measured 8x slower: 0.35 vs 0.044 ms).
// The common factor *tscale * *gscale is applied by k_trip_finish, so the fixed-point range does not depend on loss
// weights or AMP gradient scales: |feature| < 2^22 / 57 cannot overflow 63 bits.
constexpr float kTripFix = 1099511627776.0f;      // 2^40
template <typename T>
__global__ void __launch_bounds__(256) k_trip_bwd(const T* __restrict__ feats, int D, long hw, int maxT,
                                                  const int* __restrict__ sel, const int* __restrict__ kcount,
                                                  const float* __restrict__ tl, const float* __restrict__ trip,
                                                  unsigned long long* __restrict__ acc) {
  const int c = blockIdx.x, t = blockIdx.y * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int k = kcount[c];
  if (t >= k || !(tl[(size_t)c * maxT + t] > 0.f)) return;
  const float u = kTripFix / ((float)k * trip[1]);
  const int* s = sel + (size_t)c * 3 * maxT;
  const long ra = s[t], rp = s[maxT + t], rn = s[2 * maxT + t];
  const long ba = ra / hw, pa = ra - ba * hw, bp = rp / hw, pp = rp - bp * hw, bn = rn / hw, pn = rn - bn * hw;
  for (int d = lane; d < D; d += 32) {
    const float a = to_f32<T>(feats[(ba * D + d) * hw + pa]);
    const float p = to_f32<T>(feats[(bp * D + d) * hw + pp]);
    const float n = to_f32<T>(feats[(bn * D + d) * hw + pn]);
    const long long qa = __float2ll_rn(u * a);
    atomicAdd(acc + (ba * D + d) * hw + pa, (unsigned long long)__float2ll_rn(u * (n - p)));
    atomicAdd(acc + (bp * D + d) * hw + pp, (unsigned long long)(-qa));
    atomicAdd(acc + (bn * D + d) * hw + pn, (unsigned long long)qa);
  }
}
__global__ void __launch_bounds__(256) k_trip_finish(const unsigned long long* __restrict__ acc, long n,
                                                     const float* __restrict__ tscale, const float* __restrict__ gscale,
                                                     float* __restrict__ gfeat) {
  const double w0 = (double)(*tscale) * (gscale ? (double)*gscale : 1.0) / (double)kTripFix;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    gfeat[i] = (float)((double)(long long)acc[i] * w0);
}

}  // namespace sh

extern "C" {

// tab (device int32): mode 0 -> [bucket_lo ncls][bucket_hi ncls]; mode 1 -> group id per label [256] (0/1, -1 = none)
// status: [0] ready flag (#classes>0), [1] error flag (label outside the tables; the reference raises)
int sh_triplet_forward(const void* feats, int dtype, const void* label, int label_dtype, int B, int D, int h, int w, int H, int W,
                       int mode, const int* tab, int ncls, int max_triplet, int* lab_ds, int* sel, int* kcount,
                       float* tl, float* trip, int* status, void* stream) {
  if (B <= 0 || D <= 0 || h <= 0 || w <= 0 || ncls <= 0 || max_triplet <= 0) return SH_ERR_BAD_ARG;
  if (mode == 1 && ncls != 256) return SH_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long R = (long)B * h * w;
  cudaError_t e = cudaMemsetAsync(status, 0, 2 * sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  long blocks = (R + 255) / 256;
  if (blocks > SH_NUM_SMS * 4L) blocks = SH_NUM_SMS * 4L;
  SH_LABEL_SWITCH(label_dtype, L, {
    sh::k_trip_labels<L><<<(unsigned)blocks, 256, 0, st>>>((const L*)label, B, H, W, h, w, mode, tab, ncls, lab_ds, status);
  })
  SH_CHECK_LAUNCH();
  sh::k_trip_select<<<ncls, 256, 0, st>>>(lab_ds, R, mode, tab, ncls, max_triplet, sel, kcount);
  SH_CHECK_LAUNCH();
  dim3 g(ncls, (max_triplet + 7) / 8);
  const long hw = (long)h * w;
  switch (dtype) {
    case SH_DT_F32: sh::k_trip_hinge<float><<<g, 256, 0, st>>>((const float*)feats, D, hw, max_triplet, sel, kcount, tl); break;
    case SH_DT_BF16: sh::k_trip_hinge<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)feats, D, hw, max_triplet, sel, kcount, tl); break;
    case SH_DT_F16: sh::k_trip_hinge<__half><<<g, 256, 0, st>>>((const __half*)feats, D, hw, max_triplet, sel, kcount, tl); break;
    default: return SH_ERR_UNSUPPORTED;
  }
  SH_CHECK_LAUNCH();
  sh::k_trip_reduce<<<1, 256, 0, st>>>(ncls, max_triplet, kcount, tl, trip, status);
  SH_CHECK_LAUNCH();
  return SH_OK;
}

int sh_triplet_backward(const void* feats, int dtype, int B, int D, int h, int w, int ncls, int max_triplet,
                        const int* sel, const int* kcount, const float* tl, const float* trip, const float* tscale,
                        const float* gscale, float* gfeat, void* scratch, void* stream) {
  if (scratch == nullptr || gfeat == nullptr) return SH_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long hw = (long)h * w, n = (long)B * D * hw;
  unsigned long long* acc = (unsigned long long*)scratch;
  cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)n * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return (int)e;
  dim3 g(ncls, (max_triplet + 7) / 8);
  switch (dtype) {
    case SH_DT_F32: sh::k_trip_bwd<float><<<g, 256, 0, st>>>((const float*)feats, D, hw, max_triplet, sel, kcount, tl, trip, acc); break;
    case SH_DT_BF16: sh::k_trip_bwd<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)feats, D, hw, max_triplet, sel, kcount, tl, trip, acc); break;
    case SH_DT_F16: sh::k_trip_bwd<__half><<<g, 256, 0, st>>>((const __half*)feats, D, hw, max_triplet, sel, kcount, tl, trip, acc); break;
    default: return SH_ERR_UNSUPPORTED;
  }
  SH_CHECK_LAUNCH();
  long blocks = (n + 255) / 256;
  if (blocks > SH_NUM_SMS * 8L) blocks = SH_NUM_SMS * 8L;
  sh::k_trip_finish<<<(unsigned)blocks, 256, 0, st>>>(acc, n, tscale, gscale, gfeat);
  SH_CHECK_LAUNCH();
  return SH_OK;
}

}  // extern "C"
