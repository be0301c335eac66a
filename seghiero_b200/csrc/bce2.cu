// Two-level hierarchical loss: fused target-derive + sigmoid tree-min/max BCE +
// per-level softmax CE, forward AND input gradient in ONE pass over the logits.
//
// Reference arithmetic: models/loss/hiera_triplet_loss.py:11-107, 183-188 and
// models/loss/cross_entropy_loss.py:7-30 (CE mean over ALL pixels).
//
// Data flow per CTA (128 threads, VEC pixels per thread, PX = 128*VEC pixels):
//   sweep A : each logit is loaded from HBM exactly once (128-bit coalesced per
//             channel plane), e^x and sigmoid come from one ex2 + one rcp, both
//             are parked in thread-private shared-memory columns [C][PX];
//   tree    : per-bucket max / holder ids, positive terms, CE log-sum-exp;
//   sweep B : gradient assembled per channel from the parked values and
//             streamed out once.
// Algorithmic HBM bytes per pixel: 2*C*sizeof(T) + 1 (uint8 label from k_prep2).
#include "common.cuh"

namespace sh {

struct Hier2 {
  int nf, nc;
  const int* bstart;   // [nc]
  const int* bend;     // [nc] exclusive
  const int* owner;    // [nf] last bucket containing f, or -1
  const int* fb_ptr;   // [nf+1] CSR: buckets containing f
  const int* fb_idx;
  const int* lut;      // [lut_size] fine -> coarse target (255 = none)
  int lut_size;
};

// Label pre-pass: int64 -> uint8 labels, valid counts per level, range check.
// counts[0]=#fine-valid, counts[1]=#coarse-valid, counts[2]=error flag.
// 8 labels per thread and trip: 16-byte loads (int64: four, int32: two, uint8: one 8-byte load), coarse-validity table in
// shared memory, one 8-byte store.  (One label per thread with a global LUT lookup each ran at 1.1 TB/s of label traffic.)
template <typename L>
__device__ __forceinline__ void prep2_load8(const L* __restrict__ p, bool vec, long long (&t)[8]) {
  if (vec && sizeof(L) == 8) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(p) + q);
      t[2 * q] = v.x; t[2 * q + 1] = v.y;
    }
  } else if (vec && sizeof(L) == 4) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(p) + q);
      t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
    }
  } else if (vec && sizeof(L) == 1) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
#pragma unroll
    for (int k = 0; k < 4; ++k) { t[k] = (v.x >> (8 * k)) & 0xffu; t[4 + k] = (v.y >> (8 * k)) & 0xffu; }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = lab_ld(p, k);
  }
}
template <typename L>
__global__ void __launch_bounds__(256) k_prep2(const L* __restrict__ label, unsigned char* __restrict__ lab8,
                                               long n, int nf, const int* __restrict__ lut, int lut_size,
                                               unsigned long long* __restrict__ counts) {
  __shared__ unsigned char s_cv[256];       // fine label -> has a coarse target
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_cv[i] = (i < nf && i < lut_size && lut[i] != SH_IGNORE) ? 1 : 0;
  __syncthreads();
  long long nvf = 0, nvc = 0;
  bool bad = false;
  const bool vec = ((uintptr_t)label % 16 == 0) && ((uintptr_t)lab8 % 8 == 0);
  const long groups = (n + 7) / 8;
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < groups; g += (long)gridDim.x * blockDim.x) {
    const long base = 8 * g;
    const int m = (int)min(8L, n - base);
    long long t[8];
    if (m == 8) prep2_load8<L>(label + base, vec, t);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = k < m ? lab_ld(label, base + k) : (long long)SH_IGNORE;
    }
    unsigned int o[2] = {0u, 0u};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned int ob = SH_IGNORE;
      if (t[k] != SH_IGNORE) {
        nvf++;
        if (t[k] >= 0 && t[k] < nf) { ob = (unsigned int)t[k]; nvc += s_cv[ob]; }
        else bad = true;                     // F.one_hot would raise in the reference
      }
      o[k >> 2] |= ob << (8 * (k & 3));
    }
    if (m == 8 && vec) *reinterpret_cast<uint2*>(lab8 + base) = make_uint2(o[0], o[1]);
    else
      for (int k = 0; k < m; ++k) lab8[base + k] = (unsigned char)(o[k >> 2] >> (8 * (k & 3)));
  }
  nvf = warp_sum(nvf);
  nvc = warp_sum(nvc);
  if ((threadIdx.x & 31) == 0) {
    if (nvf) atomicAdd(counts, (unsigned long long)nvf);
    if (nvc) atomicAdd(counts + 1, (unsigned long long)nvc);
  }
  if (bad) atomicOr((unsigned int*)(counts + 2), 1u);
}

template <typename T, int VEC, bool GRAD>
__global__ void __launch_bounds__(128) k_bce2_fused(const T* __restrict__ x, const unsigned char* __restrict__ lab8,
                                                    T* __restrict__ grad, int B, long HW, Hier2 h, float eps,
                                                    float loss_weight, const unsigned long long* __restrict__ counts,
                                                    float* __restrict__ partials, int vec_ok) {
  constexpr int PX = 128 * VEC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = h.nf + h.nc;
  float* S = reinterpret_cast<float*>(smem_raw);   // [C][PX] sigmoid
  float* V = S + (size_t)C * PX;                   // [C][PX] e^x
  unsigned char* HOLD = reinterpret_cast<unsigned char*>(V + (size_t)C * PX);  // [nc][PX]

  const int col0 = threadIdx.x * VEC;
  const long chunks = (HW + PX - 1) / PX;
  const long items = chunks * B;
  const float nvf = fmaxf((float)counts[0], 1.0f), nvc = fmaxf((float)counts[1], 1.0f);
  const float wF = 5.0f * loss_weight / (nvf * (float)h.nf);
  const float wC = 5.0f * loss_weight / (nvc * (float)h.nc);
  const float wCE = loss_weight / ((float)B * (float)HW);

  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // sum BCE fine, BCE coarse, CE fine, CE coarse (unnormalised)

  for (long item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / chunks);
    const long p0 = (item - (long)b * chunks) * PX + col0;
    const T* xb = x + (long)b * C * HW;
    int tf[VEC], tc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      int t = (p0 + v < HW) ? (int)lab8[(long)b * HW + p0 + v] : SH_IGNORE;
      tf[v] = t;
      tc[v] = (t != SH_IGNORE && t < h.lut_size) ? h.lut[t] : SH_IGNORE;
    }
    float sumv_f[VEC], sumv_c[VEC], prod[VEC], lf[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { sumv_f[v] = 0.f; sumv_c[v] = 0.f; prod[v] = 1.f; lf[v] = 0.f; }

    // ---- sweep A: the only HBM read of the logits -------------------------------------------
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      float xv[VEC];
      load_n<T, VEC>(xb + (long)c * HW, p0, HW, vec_ok != 0, xv);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        SigExp se = sig_exp(xv[v]);
        S[(size_t)c * PX + col0 + v] = se.s;
        V[(size_t)c * PX + col0 + v] = se.v;
        if (c < h.nf) {
          sumv_f[v] += se.v;
          if (c != tf[v]) prod[v] *= (1.0f - se.s) + eps;   // literal fp32 order of the reference
        } else {
          sumv_c[v] += se.v;
        }
      }
      if ((c & 3) == 3 || c == h.nf - 1) {   // one log per <=4 factors: each factor >= eps = 1e-8
#pragma unroll
        for (int v = 0; v < VEC; ++v) { lf[v] -= fast_log(prod[v]); prod[v] = 1.f; }
      }
    }

    // ---- tree logic on the parked sigmoids ---------------------------------------------------
    bool hold_pos_a[VEC];
    float inv_f[VEC], inv_c[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int col = col0 + v;
      const bool vf = tf[v] != SH_IGNORE, vc = tc[v] != SH_IGNORE;
      float lc = 0.f, pc = 1.f, pos_c = 1.f;
      for (int i = 0; i < h.nc; ++i) {
        float best = -1.f;
        int hold = 0;
        for (int f = h.bstart[i]; f < h.bend[i]; ++f) {
          float a = S[(size_t)f * PX + col];
          if (a > best) { best = a; hold = f; }
        }
        const float bi = S[(size_t)(h.nf + i) * PX + col];
        if (bi > best) { best = bi; hold = h.nf + i; }
        HOLD[(size_t)i * PX + col] = (unsigned char)hold;
        if (i != tc[v]) pc *= (1.0f - best) + eps; else pos_c = bi;
        if ((i & 3) == 3) { lc -= fast_log(pc); pc = 1.f; }
      }
      lc -= fast_log(pc);
      hold_pos_a[v] = true;
      if (vf) {
        const float a = S[(size_t)tf[v] * PX + col];
        float m = a;
        const int o = h.owner[tf[v]];
        if (o >= 0) {
          const float bo = S[(size_t)(h.nf + o) * PX + col];
          if (!(a <= bo)) { m = bo; hold_pos_a[v] = false; }
        }
        acc[0] += lf[v] - fast_log(m + eps);
        acc[2] += fast_log(sumv_f[v]) - fast_log(V[(size_t)tf[v] * PX + col]);
      }
      if (vc) {
        acc[1] += lc - fast_log(pos_c + eps);
        acc[3] += fast_log(sumv_c[v]) - fast_log(V[(size_t)(h.nf + tc[v]) * PX + col]);
      }
      inv_f[v] = rcp(sumv_f[v]);
      inv_c[v] = rcp(sumv_c[v]);
    }

    // ---- sweep B: gradient, written once ------------------------------------------------------
    if (GRAD) {
      T* gb = grad + (long)b * C * HW;
      for (int c = 0; c < C; ++c) {
        float g[VEC];
        const bool fine = c < h.nf;
        const int k0 = fine ? h.fb_ptr[c] : 0, k1 = fine ? h.fb_ptr[c + 1] : 0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const int col = col0 + v;
          const float s = S[(size_t)c * PX + col], ev = V[(size_t)c * PX + col];
          const bool vf = tf[v] != SH_IGNORE, vc = tc[v] != SH_IGNORE;
          const float q = 1.0f - s;
          float ds = 0.f, ce = 0.f;
          if (fine) {
            if (vf) {
              if (c == tf[v]) { if (hold_pos_a[v]) ds -= wF * rcp(s + eps); }
              else ds += wF * rcp(q + eps);
              ce = wCE * (ev * inv_f[v] - (c == tf[v] ? 1.f : 0.f));
            }
            if (vc) {
              for (int k = k0; k < k1; ++k) {
                const int i = h.fb_idx[k];
                if (i != tc[v] && HOLD[(size_t)i * PX + col] == c) ds += wC * rcp(q + eps);
              }
            }
          } else {
            const int i = c - h.nf;
            if (vc) {
              if (i == tc[v]) ds -= wC * rcp(s + eps);
              else if (HOLD[(size_t)i * PX + col] == c) ds += wC * rcp(q + eps);
              ce = wCE * (ev * inv_c[v] - (i == tc[v] ? 1.f : 0.f));
            }
            if (vf && !hold_pos_a[v] && h.owner[tf[v]] == i) ds -= wF * rcp(s + eps);
          }
          g[v] = ds * (q * s) + ce;
        }
        store_n<T, VEC>(gb + (long)c * HW, p0, HW, vec_ok != 0, g);
      }
    }
  }

  __shared__ float red[4 * 4];
  float r = block_sum_k<4>(acc, red);
  if (threadIdx.x < 4) partials[(size_t)blockIdx.x * 4 + threadIdx.x] = r;
}

// ---------------------------------------------------------------------------------------------
// Fast path: buckets are disjoint ranges (every fine class sits in at most one bucket), HW % 4 == 0, C <= 64.
// Thread = 4 consecutive pixels.  Channels are walked in tree order (a bucket's fine children, then its coarse
// channel; orphan fines last) in BOTH sweeps, so the bucket max / its holder live in registers in sweep A and the
// holder bytes are decoded once per bucket in sweep B.  Logits arrive through a thread-private cp.async ring that
// runs XD-2 channels ahead (also across work items).  Only w = e^-x is parked in shared memory (16 bytes per
// thread and channel); sweep B rebuilds sigmoid and e^x from it with the same instructions, i.e. the same bits.
// Label logic is SIMD over the 4 label bytes; the terms of a pixel's own target classes take a separate path
// that only runs for channels that are a target somewhere in the strip.
// Same arithmetic as k_bce2_fused (hiera_triplet_loss.py:41-107, cross_entropy_loss.py:7-30).
// ---------------------------------------------------------------------------------------------
constexpr int F2_NT = 128, F2_VEC = 4, F2_PX = F2_NT * F2_VEC, F2_XD = 8;

__device__ __forceinline__ bool b0(unsigned int z, int k) { return ((z >> (8 * k)) & 0xffu) == 0u; }

// w = e^-x (|x| clamped), s = sigmoid with torch's rounding of 1 + w, E = e^x
__device__ __forceinline__ float exp_neg(float x) { return ex2(clamp_nan(x * (-kLog2e), -115.0f, 115.0f)); }
__device__ __forceinline__ void sig_from_w(float w, float& s, float& E) {
  const float y = 1.0f + w;
  const float q0 = rcp(y);
  s = fmaf(q0, fmaf(-y, q0, 1.0f), q0);
  E = rcp(w);
}

template <typename T, bool GRAD>
__global__ void __launch_bounds__(F2_NT, 3) k_bce2_fast(const T* __restrict__ x, const unsigned char* __restrict__ lab8,
                                                        T* __restrict__ grad, int B, long HW, Hier2 h, float eps,
                                                        float loss_weight, const unsigned long long* __restrict__ counts,
                                                        float* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = h.nf + h.nc;
  float4* Wp = reinterpret_cast<float4*>(smem_raw);                                    // [C][NT] e^-x of the 4 pixels
  uint4* xst = reinterpret_cast<uint4*>(Wp + (size_t)C * F2_NT);                       // [XD][NT] 16-byte slots
  unsigned int* HOLD = reinterpret_cast<unsigned int*>(xst + F2_XD * F2_NT);           // [nc][NT] holder channel of the 4 pixels
  unsigned int* s_order = HOLD + (size_t)h.nc * F2_NT;                                 // [C] channel | kind << 8 | first << 10 | bucket << 16
  int* s_lut = reinterpret_cast<int*>(s_order + C);                                    // [lut_size]

  const int tid = threadIdx.x;
  for (int i = tid; i < h.lut_size; i += F2_NT) s_lut[i] = h.lut[i];
  if (tid == 0) {
    int n = 0;
    for (int i = 0; i < h.nc; ++i) {
      bool first = true;
      for (int f = h.bstart[i]; f < h.bend[i]; ++f) { s_order[n++] = (unsigned)f | (first ? 1u << 10 : 0u) | ((unsigned)i << 16); first = false; }
      s_order[n++] = (unsigned)(h.nf + i) | (1u << 8) | (first ? 1u << 10 : 0u) | ((unsigned)i << 16);
    }
    for (int f = 0; f < h.nf; ++f)
      if (h.owner[f] < 0) s_order[n++] = (unsigned)f | (3u << 8) | (0xffu << 16);
  }
  __syncthreads();

  const long chunks = (HW + F2_PX - 1) / F2_PX;
  const long items = chunks * B;
  const float nvf = fmaxf((float)counts[0], 1.0f), nvc = fmaxf((float)counts[1], 1.0f);
  const float wF = 5.0f * loss_weight / (nvf * (float)h.nf);
  const float wC = 5.0f * loss_weight / (nvc * (float)h.nc);
  const float wCE = loss_weight / ((float)B * (float)HW);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // sum BCE fine, BCE coarse, CE fine, CE coarse (unnormalised)

  // prefetch stream over (item, order index)
  const unsigned int xs_base = (unsigned int)__cvta_generic_to_shared(xst + tid);
  const unsigned char* xs_gen = reinterpret_cast<const unsigned char*>(xst + tid);
  long pf_item = blockIdx.x;
  int pf_ci = 0;
  unsigned int pf_seq = 0;
  const char* pf_ptr = reinterpret_cast<const char*>(x);
  auto pf_setup = [&]() {
    if (pf_item < items) {
      const int b = (int)(pf_item / chunks);
      long p0 = (pf_item - (long)b * chunks) * F2_PX + tid * F2_VEC;
      if (p0 >= HW) p0 = 0;                                  // ragged last chunk: any valid address will do
      pf_ptr = reinterpret_cast<const char*>(x + (long)b * C * HW + p0);
    }
  };
  auto pf_issue = [&]() {
    if (pf_item < items) {
      const char* g = pf_ptr + (long)(s_order[pf_ci] & 0xffu) * HW * (long)sizeof(T);
      const unsigned int dst = xs_base + (pf_seq & (F2_XD - 1)) * (F2_NT * 16);
      if (sizeof(T) == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g));
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(g));
    }
    cp_async_commit();
    ++pf_seq;
    if (++pf_ci == C) { pf_ci = 0; pf_item += gridDim.x; pf_setup(); }
  };
  pf_setup();
#pragma unroll 1
  for (int q = 0; q < F2_XD - 2; ++q) pf_issue();
  unsigned int seq = 0;

#pragma unroll 1
  for (long item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / chunks);
    const long p0 = (item - (long)b * chunks) * F2_PX + tid * F2_VEC;
    const bool inb = p0 < HW;
    // ---- labels of the strip: fine target channel bytes, coarse target channel bytes (0xff = none) ----
    const unsigned int tf4 = inb ? *reinterpret_cast<const unsigned int*>(lab8 + (long)b * HW + p0) : 0xffffffffu;
    unsigned int tb4 = 0xffffffffu;       // coarse target bucket per pixel
    unsigned int tc4 = 0xffffffffu;       // coarse target channel per pixel
    unsigned long long present = 0ull;    // channels that are a target of some pixel of the strip
    float vf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned int t = (tf4 >> (8 * k)) & 0xffu;
      vf[k] = t != SH_IGNORE ? 1.f : 0.f;
      if (t != SH_IGNORE) {
        present |= 1ull << t;
        const unsigned int i = t < (unsigned)h.lut_size ? (unsigned)s_lut[t] : 0xffu;
        if (i != 0xffu) {
          tb4 = (tb4 & ~(0xffu << (8 * k))) | (i << (8 * k));
          tc4 = (tc4 & ~(0xffu << (8 * k))) | (((unsigned)h.nf + i) << (8 * k));
          present |= 1ull << (h.nf + i);
        }
      }
    }
    float sumF[4], sumC[4], prodF[4], prodC[4], Lf[4], Lc[4], rmax[4], a_t[4], b_t[4], xt_f[4], xt_c[4];
    unsigned int rhold[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sumF[k] = sumC[k] = 0.f; prodF[k] = prodC[k] = 1.f; Lf[k] = Lc[k] = 0.f;
      rmax[k] = -1.f; a_t[k] = b_t[k] = 1.f; xt_f[k] = xt_c[k] = 0.f; rhold[k] = 0u;
    }
    unsigned int hpa4 = 0xffffffffu;      // byte k = 0xff: the fine channel holds min(A_t, B_c(t)) (fine wins ties, :91-92)
    int nF = 0, nC = 0;

    // ---- sweep A: the only HBM read of the logits -------------------------------------------
#pragma unroll 1
    for (int ci = 0; ci < C; ++ci) {
      pf_issue();
      cp_async_wait<F2_XD - 2>();
      float xv[4];
      staged_vec4<T>(xs_gen + (seq & (F2_XD - 1)) * (F2_NT * 16), xv);
      ++seq;
      const unsigned int oe = s_order[ci];
      const unsigned int ch = oe & 0xffu, kind = (oe >> 8) & 3u, bucket = oe >> 16;
      const unsigned int cc = ch * 0x01010101u;
      float w[4], s[4], E[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { w[k] = exp_neg(xv[k]); sig_from_w(w[k], s[k], E[k]); }
      Wp[(size_t)ch * F2_NT + tid] = make_float4(w[0], w[1], w[2], w[3]);
      if (kind != 1u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sumF[k] += E[k];
          prodF[k] *= (1.0f - s[k]) + eps;                    // literal fp32 order of the reference
          if (kind == 0u) {
            if (oe & (1u << 10)) { rmax[k] = -1.f; rhold[k] = 0u; }
            if (s[k] > rmax[k]) { rmax[k] = s[k]; rhold[k] = ch; }               // lowest fine id wins ties (:84-85)
          }
        }
        if ((present >> ch) & 1ull) {
          const unsigned int z = tf4 ^ cc;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (b0(z, k)) { a_t[k] = s[k]; xt_f[k] = xv[k]; Lf[k] -= lg2((1.0f - s[k]) + eps); }   // the target's own factor is not in the product
        }
        if (((++nF) & 3) == 0) {   // one log per <= 4 factors: each factor >= eps = 1e-8
#pragma unroll
          for (int k = 0; k < 4; ++k) { Lf[k] += lg2(prodF[k]); prodF[k] = 1.f; }
        }
      } else {
        unsigned int hd = 0;
        float best[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sumC[k] += E[k];
          best[k] = (oe & (1u << 10)) ? -1.f : rmax[k];
          unsigned int hold = (oe & (1u << 10)) ? 0u : rhold[k];
          if (s[k] > best[k]) { best[k] = s[k]; hold = ch; }                     // the coarse logit comes last in the cat
          hd |= hold << (8 * k);
          prodC[k] *= (1.0f - best[k]) + eps;
        }
        HOLD[(size_t)bucket * F2_NT + tid] = hd;
        if ((present >> ch) & 1ull) {
          const unsigned int z = tb4 ^ (bucket * 0x01010101u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (b0(z, k)) {
              b_t[k] = s[k]; xt_c[k] = xv[k];
              Lc[k] -= lg2((1.0f - best[k]) + eps);
              if (!(a_t[k] <= s[k])) hpa4 &= ~(0xffu << (8 * k));
            }
        }
        if (((++nC) & 3) == 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { Lc[k] += lg2(prodC[k]); prodC[k] = 1.f; }
        }
      }
    }
    float inv_f[4], inv_c[4];
    unsigned int fpos4 = 0xffffffffu;     // channel that holds the fine target's positive term, per pixel
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned int t = (tf4 >> (8 * k)) & 0xffu, i = (tb4 >> (8 * k)) & 0xffu;
      const bool vfk = t != SH_IGNORE, vck = i != 0xffu;
      const bool fine_holds = !vck || ((hpa4 >> (8 * k)) & 1u);
      if (vfk) {
        const float m = fine_holds ? a_t[k] : b_t[k];
        acc[0] += -kLn2 * (Lf[k] + lg2(prodF[k])) - fast_log(m + eps);
        acc[2] += fast_log(sumF[k]) - xt_f[k];
        fpos4 = (fpos4 & ~(0xffu << (8 * k))) | ((fine_holds ? t : (unsigned)h.nf + i) << (8 * k));
      }
      if (vck) {
        acc[1] += -kLn2 * (Lc[k] + lg2(prodC[k])) - fast_log(b_t[k] + eps);
        acc[3] += fast_log(sumC[k]) - xt_c[k];
      }
      inv_f[k] = vfk ? rcp(sumF[k]) : 0.f;      // zero on pixels the level ignores: their CE gradient vanishes
      inv_c[k] = vck ? rcp(sumC[k]) : 0.f;
    }

    // ---- sweep B: gradient, written once ------------------------------------------------------
    if (GRAD) {
      T* gb = grad + (long)b * C * HW + p0;
      const unsigned int novc = __vcmpeq4(tb4, 0xffffffffu);
      unsigned int hdN = 0xffffffffu;
#pragma unroll 1
      for (int ci = 0; ci < C; ++ci) {
        const unsigned int oe = s_order[ci];
        const unsigned int ch = oe & 0xffu, kind = (oe >> 8) & 3u, bucket = oe >> 16;
        const unsigned int cc = ch * 0x01010101u;
        if (oe & (1u << 10))          // bucket start: who holds the bucket max, where that term counts (not at the bucket's own targets)
          hdN = HOLD[(size_t)bucket * F2_NT + tid] | __vcmpeq4(tb4, bucket * 0x01010101u) | novc;
        if (kind == 3u) hdN = 0xffffffffu;
        const float4 w4 = Wp[(size_t)ch * F2_NT + tid];
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
        const float wbase = kind == 1u ? 0.f : wF;
        const float* invp = kind == 1u ? inv_c : inv_f;
        const unsigned int zH = hdN ^ cc;
        const bool pos = (present >> ch) & 1ull;
        const unsigned int zT = (kind == 1u ? tc4 : tf4) ^ cc, zP = fpos4 ^ cc;
        float g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float s, E;
          sig_from_w(w[k], s, E);
          const float t = 1.0f - s;
          const float r = rcp(t + eps);
          float ds = wbase * r;
          if (b0(zH, k)) ds = fmaf(wC, r, ds);
          float oh = 0.f;
          if (pos) {
            float Bp = b0(zP, k) ? wF : 0.f;
            if (b0(zT, k)) {
              oh = 1.f;
              ds = fmaf(-wbase, r, ds);                        // the target has no own (1 - s) term
              if (kind == 1u) Bp += wC;
            }
            ds = fmaf(-Bp, rcp(s + eps), ds);
          }
          const float qv = s * t * vf[k];
          g[k] = fmaf(ds, qv, wCE * fmaf(E, invp[k], -oh));
        }
        if (inb) VecIO<T, 4>::store(gb + (long)ch * HW, g);
      }
    }
  }
  cp_async_wait<0>();

  __shared__ float red[4 * 4];
  float r = block_sum_k<4>(acc, red);
  if (threadIdx.x < 4) partials[(size_t)blockIdx.x * 4 + threadIdx.x] = r;
}

static size_t bce2_fast_smem(int nf, int nc, int lut_size) {
  const int C = nf + nc;
  size_t s = (size_t)C * F2_NT * 16 + (size_t)F2_XD * F2_NT * 16 + (size_t)nc * F2_NT * 4;
  s += (size_t)C * 4 + (size_t)lut_size * 4 + 32;
  return (s + 15) & ~(size_t)15;
}

// Sum per-CTA fp32 partials [n][K] into double out[K] in a fixed order (deterministic).
__global__ void __launch_bounds__(256) k_reduce_partials(const float* __restrict__ part, int n, int K,
                                                         double* __restrict__ out) {
  __shared__ double sm[256];
  for (int k = 0; k < K; ++k) {
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) a += (double)part[(size_t)i * K + k];
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

// cosine schedule of hiera_triplet_loss.py:204-208 / rmi_hiera_triplet_loss.py:539-543, evaluated on device
__device__ __forceinline__ double schedule_factor(double step, double total) {
  if (step < total) return 0.25 * (1.0 + cos((step - total) / total * 3.141592653589793));
  return 0.5;
}

// loss = (5*(Lf/(NvF*nf) + Lc/(NvC*nc)) + CEf/Npx + CEc/Npx + ready*factor*triplet) * loss_weight
// out[0] = loss, out[1] = factor*ready*loss_weight (the scale the triplet backward needs)
__global__ void k_loss2_final(const double* __restrict__ sums, const unsigned long long* __restrict__ counts, int nf,
                              int nc, double npx, const double* __restrict__ step, double total_steps,
                              const float* __restrict__ trip /* [0]=mean over classes, [1]=class count */,
                              const int* __restrict__ ready, float loss_weight, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double nvf = fmax((double)counts[0], 1.0), nvc = fmax((double)counts[1], 1.0);
  double loss = 5.0 * (sums[0] / (nvf * nf) + sums[1] / (nvc * nc)) + sums[2] / npx + sums[3] / npx;
  double tscale = 0.0;
  if (trip != nullptr && ready != nullptr && *ready > 0 && trip[1] > 0.f) {
    const double f = schedule_factor(*step, total_steps);
    loss += f * (double)trip[0];
    tscale = f * loss_weight;
  }
  loss *= loss_weight;
  if (counts[2] != 0) loss = __longlong_as_double(0x7ff8000000000000LL);  // out-of-range labels poison the loss
  // a label the triplet tables do not know (the reference raises IndexError from hiera_map[ii]) poisons it too
  if (ready != nullptr && ready[1] != 0) loss = __longlong_as_double(0x7ff8000000000000LL);
  out[0] = (float)loss;
  out[1] = (float)tscale;
}

// grad *= scale[0] unless scale[0] == 1 (used when autograd hands a non-unit grad_output)
template <typename T>
__global__ void __launch_bounds__(256) k_scale_inplace(T* __restrict__ g, long n, const float* __restrict__ scale) {
  const float s = *scale;
  if (s == 1.0f) return;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    g[i] = from_f32<T>(to_f32<T>(g[i]) * s);
}

template <typename T, int VEC>
static int launch_bce2_generic(const void* x, const unsigned char* lab8, void* grad, int B, long HW, const Hier2& h, float eps,
                               float lw, const unsigned long long* counts, float* partials, int grid, cudaStream_t st);

template <typename T>
static int launch_bce2(const void* x, const unsigned char* lab8, void* grad, int B, long HW, const Hier2& h, float eps,
                       float lw, const unsigned long long* counts, float* partials, int grid, bool tree, cudaStream_t st) {
  const size_t fsm = bce2_fast_smem(h.nf, h.nc, h.lut_size);
  if (tree && HW % 4 == 0 && h.nf + h.nc <= 64 && (uintptr_t)x % (4 * sizeof(T)) == 0 &&
      (uintptr_t)grad % (4 * sizeof(T)) == 0 && fsm <= 113 * 1024) {
    if (grad) {
      auto kern = k_bce2_fast<T, true>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
      kern<<<grid, F2_NT, fsm, st>>>((const T*)x, lab8, (T*)grad, B, HW, h, eps, lw, counts, partials);
    } else {
      auto kern = k_bce2_fast<T, false>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
      kern<<<grid, F2_NT, fsm, st>>>((const T*)x, lab8, nullptr, B, HW, h, eps, lw, counts, partials);
    }
    SH_CHECK_LAUNCH();
    return SH_OK;
  }
  const int C = h.nf + h.nc;
  // sigmoid and e^x of every channel are parked in shared memory: 256 pixels per CTA (2 per thread), or 128 (1 per
  // thread) for hierarchies whose 256-pixel tile does not fit (e.g. 150 fine + 30 coarse classes)
  if ((size_t)C * 256 * 8 + (size_t)h.nc * 256 <= 227 * 1024) return launch_bce2_generic<T, 2>(x, lab8, grad, B, HW, h, eps, lw, counts, partials, grid, st);
  if ((size_t)C * 128 * 8 + (size_t)h.nc * 128 <= 227 * 1024) return launch_bce2_generic<T, 1>(x, lab8, grad, B, HW, h, eps, lw, counts, partials, grid, st);
  return SH_ERR_UNSUPPORTED;
}

template <typename T, int VEC>
static int launch_bce2_generic(const void* x, const unsigned char* lab8, void* grad, int B, long HW, const Hier2& h, float eps,
                               float lw, const unsigned long long* counts, float* partials, int grid, cudaStream_t st) {
  constexpr int PX = 128 * VEC;
  const int C = h.nf + h.nc;
  const size_t smem = (size_t)C * PX * 8 + (size_t)h.nc * PX;
  bool vec_ok = (HW % VEC == 0) && ((uintptr_t)x % (VEC * sizeof(T)) == 0) &&
                (grad == nullptr || (uintptr_t)grad % (VEC * sizeof(T)) == 0);
  if (grad) {
    auto kern = k_bce2_fused<T, VEC, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 128, smem, st>>>((const T*)x, lab8, (T*)grad, B, HW, h, eps, lw, counts, partials, vec_ok);
  } else {
    auto kern = k_bce2_fused<T, VEC, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 128, smem, st>>>((const T*)x, lab8, nullptr, B, HW, h, eps, lw, counts, partials, vec_ok);
  }
  SH_CHECK_LAUNCH();
  return SH_OK;
}

}  // namespace sh

extern "C" {

// Number of CTAs sh_bce2_fwdbwd launches (= rows of the `partials` workspace, 4 floats each).
int sh_bce2_grid(int B, long HW, int C, int n_coarse) {
  const int PX = 256;
  long items = ((HW + PX - 1) / PX) * (long)B;
  // shared memory of the larger of the two kernels (the fast one adds its cp.async ring and tables)
  size_t smem = (size_t)C * PX * 8 + (size_t)n_coarse * PX + (size_t)sh::F2_XD * sh::F2_NT * 16 + 2048;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 16) per_sm = 16;
  long grid = (long)SH_NUM_SMS * per_sm;
  if (grid > items) grid = items;
  if (grid < 1) grid = 1;
  return (int)grid;
}

// hier_tab: device int32 blob laid out as
//   [bstart nc][bend nc][owner nf][fb_ptr nf+1][fb_idx n_fb][lut lut_size]
int sh_bce2_fwdbwd(const void* logits, int dtype, const void* label, int label_dtype, void* grad /* nullable */, int B, long HW,
                   int n_fine, int n_coarse, const int* hier_tab, int n_fb, int lut_size, float eps, float loss_weight,
                   unsigned char* lab8 /* [B*HW] */, unsigned long long* counts /* [4], zeroed by callee */,
                   float* partials /* [grid*4] */, double* sums /* [4] */, int stages, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B <= 0 || HW <= 0 || n_fine <= 0 || n_coarse <= 0 || n_fine + n_coarse > 255) return SH_ERR_BAD_ARG;
  sh::Hier2 h;
  h.nf = n_fine; h.nc = n_coarse;
  h.bstart = hier_tab;
  h.bend = h.bstart + n_coarse;
  h.owner = h.bend + n_coarse;
  h.fb_ptr = h.owner + n_fine;
  h.fb_idx = h.fb_ptr + n_fine + 1;
  h.lut = h.fb_idx + n_fb;
  h.lut_size = lut_size;
  if (stages & 1) {
    cudaError_t e = cudaMemsetAsync(counts, 0, 4 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    const long n = (long)B * HW;
    long pb = (n + 255) / 256;
    if (pb > SH_NUM_SMS * 8L) pb = SH_NUM_SMS * 8L;
    SH_LABEL_SWITCH(label_dtype, L, {
      sh::k_prep2<L><<<(unsigned)pb, 256, 0, st>>>((const L*)label, lab8, n, n_fine, h.lut, lut_size, counts);
    })
    SH_CHECK_LAUNCH();
  }
  const int grid = sh_bce2_grid(B, HW, n_fine + n_coarse, n_coarse);
  int rc = SH_OK;
  if (stages & 2) switch (dtype) {
    case SH_DT_F32: rc = sh::launch_bce2<float>(logits, lab8, grad, B, HW, h, eps, loss_weight, counts, partials, grid, (stages & 256) != 0, st); break;
    case SH_DT_BF16: rc = sh::launch_bce2<__nv_bfloat16>(logits, lab8, grad, B, HW, h, eps, loss_weight, counts, partials, grid, (stages & 256) != 0, st); break;
    case SH_DT_F16: rc = sh::launch_bce2<__half>(logits, lab8, grad, B, HW, h, eps, loss_weight, counts, partials, grid, (stages & 256) != 0, st); break;
    default: return SH_ERR_UNSUPPORTED;
  }
  if (rc != SH_OK) return rc;
  if (stages & 4) {
    sh::k_reduce_partials<<<1, 256, 0, st>>>(partials, grid, 4, sums);
    SH_CHECK_LAUNCH();
  }
  return SH_OK;
}

int sh_loss2_final(const double* sums, const unsigned long long* counts, int n_fine, int n_coarse, double npx,
                   const double* step, double total_steps, const float* trip, const int* ready, float loss_weight,
                   float* out, void* stream) {
  sh::k_loss2_final<<<1, 32, 0, (cudaStream_t)stream>>>(sums, counts, n_fine, n_coarse, npx, step, total_steps, trip,
                                                        ready, loss_weight, out);
  SH_CHECK_LAUNCH();
  return SH_OK;
}

int sh_scale_inplace(void* grad, int dtype, long n, const float* scale, void* stream) {
  if (n <= 0) return SH_OK;
  long blocks = (n + 255) / 256;
  if (blocks > SH_NUM_SMS * 8L) blocks = SH_NUM_SMS * 8L;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case SH_DT_F32: sh::k_scale_inplace<float><<<(unsigned)blocks, 256, 0, st>>>((float*)grad, n, scale); break;
    case SH_DT_BF16: sh::k_scale_inplace<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((__nv_bfloat16*)grad, n, scale); break;
    case SH_DT_F16: sh::k_scale_inplace<__half><<<(unsigned)blocks, 256, 0, st>>>((__half*)grad, n, scale); break;
    default: return SH_ERR_UNSUPPORTED;
  }
  SH_CHECK_LAUNCH();
  return SH_OK;
}

}  // extern "C"
