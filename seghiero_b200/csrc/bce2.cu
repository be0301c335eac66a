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
__global__ void __launch_bounds__(256) k_prep2(const long long* __restrict__ label, unsigned char* __restrict__ lab8,
                                               long n, int nf, const int* __restrict__ lut, int lut_size,
                                               unsigned long long* __restrict__ counts) {
  long long nvf = 0, nvc = 0;
  bool bad = false;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    long long t = label[i];
    unsigned char o = SH_IGNORE;
    if (t != SH_IGNORE) {
      if (t >= 0 && t < nf) {
        o = (unsigned char)t;
        nvf++;
        if (t < lut_size && lut[t] != SH_IGNORE) nvc++;
      } else {
        bad = true;  // F.one_hot would raise in the reference
        nvf++;
      }
    }
    lab8[i] = o;
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
// Fast path: buckets are disjoint ranges (every fine class sits in at most one bucket).  Channels are walked in
// tree order (a bucket's fine children, then its coarse channel; orphan fines last), so the bucket max and its
// holder live in registers; logits arrive through a thread-private cp.async ring that runs XD-2 channels ahead
// (also across work items); sigmoid / e^x are parked as float2 in shared memory for the gradient sweep.
// Same arithmetic as k_bce2_fused (hiera_triplet_loss.py:41-107, cross_entropy_loss.py:7-30).
// ---------------------------------------------------------------------------------------------
constexpr int F2_NT = 128, F2_VEC = 2, F2_PX = F2_NT * F2_VEC, F2_XD = 8;

template <typename T, bool GRAD>
__global__ void __launch_bounds__(F2_NT, 3) k_bce2_fast(const T* __restrict__ x, const unsigned char* __restrict__ lab8,
                                                        T* __restrict__ grad, int B, long HW, Hier2 h, float eps,
                                                        float loss_weight, const unsigned long long* __restrict__ counts,
                                                        float* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = h.nf + h.nc;
  float2* S = reinterpret_cast<float2*>(smem_raw);                                     // [C][NT] sigmoid of the 2 pixels
  float2* V = S + (size_t)C * F2_NT;                                                   // [C][NT] e^x
  unsigned long long* xst = reinterpret_cast<unsigned long long*>(V + (size_t)C * F2_NT);   // [XD][NT] 8-byte slots
  unsigned int* s_order = reinterpret_cast<unsigned int*>(xst + F2_XD * F2_NT);        // [C] channel | kind << 8 | first << 10 | bucket << 16
  int* s_owner = reinterpret_cast<int*>(s_order + C);                                  // [nf]
  int* s_lut = s_owner + h.nf;                                                         // [lut_size]
  unsigned short* HOLD = reinterpret_cast<unsigned short*>(s_lut + h.lut_size);        // [nc][NT] holder channel of the 2 pixels
  __shared__ int s_n;

  const int tid = threadIdx.x;
  for (int i = tid; i < h.nf; i += F2_NT) s_owner[i] = h.owner[i];
  for (int i = tid; i < h.lut_size; i += F2_NT) s_lut[i] = h.lut[i];
  if (tid == 0) {
    int n = 0;
    for (int i = 0; i < h.nc; ++i) {
      bool first = true;
      for (int f = h.bstart[i]; f < h.bend[i]; ++f) { s_order[n++] = (unsigned)f | (first ? 1u << 10 : 0u) | ((unsigned)i << 16); first = false; }
      s_order[n++] = (unsigned)(h.nf + i) | (1u << 8) | (first ? 1u << 10 : 0u) | ((unsigned)i << 16);
    }
    for (int f = 0; f < h.nf; ++f)
      if (h.owner[f] < 0) s_order[n++] = (unsigned)f | (3u << 8);
    s_n = n;
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
  const char* pf_ptr = nullptr;
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
      const unsigned int dst = xs_base + (pf_seq & (F2_XD - 1)) * (F2_NT * 8);
      if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(g));
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(g));
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
    int tf[2], tc[2];
    {
      const unsigned int t2 = inb ? *reinterpret_cast<const unsigned short*>(lab8 + (long)b * HW + p0) : 0xffffu;
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int t = (t2 >> (8 * v)) & 0xff;
        tf[v] = t;
        tc[v] = (t != SH_IGNORE && t < h.lut_size) ? s_lut[t] : SH_IGNORE;
      }
    }
    float sumF[2] = {0.f, 0.f}, sumC[2] = {0.f, 0.f}, prodF[2] = {1.f, 1.f}, prodC[2] = {1.f, 1.f};
    float lf[2] = {0.f, 0.f}, lc[2] = {0.f, 0.f}, rmax[2] = {-1.f, -1.f}, a_t[2] = {1.f, 1.f}, b_t[2] = {1.f, 1.f};
    float xt_f[2] = {0.f, 0.f}, xt_c[2] = {0.f, 0.f};
    unsigned int rhold[2] = {0u, 0u};
    bool hold_pos_a[2] = {true, true};
    int nF = 0, nC = 0;

    // ---- sweep A: the only HBM read of the logits -------------------------------------------
#pragma unroll 1
    for (int ci = 0; ci < C; ++ci) {
      pf_issue();
      cp_async_wait<F2_XD - 2>();
      float xv[2];
      {
        const unsigned char* slot = xs_gen + (seq & (F2_XD - 1)) * (F2_NT * 8);
        ++seq;
        if (sizeof(T) == 4) { const float2 t2 = *reinterpret_cast<const float2*>(slot); xv[0] = t2.x; xv[1] = t2.y; }
        else { xv[0] = staged_elem<T>(slot, 0); xv[1] = staged_elem<T>(slot, 1); }
      }
      const unsigned int oe = s_order[ci];
      const int ch = oe & 0xff, kind = (oe >> 8) & 3, bucket = oe >> 16;
      float s[2], E[2];
#pragma unroll
      for (int v = 0; v < 2; ++v) sig_exp3(xv[v], s[v], E[v]);
      S[(size_t)ch * F2_NT + tid] = make_float2(s[0], s[1]);
      V[(size_t)ch * F2_NT + tid] = make_float2(E[0], E[1]);
      if (kind != 1) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          sumF[v] += E[v];
          if (ch != tf[v]) prodF[v] *= (1.0f - s[v]) + eps;   // literal fp32 order of the reference
          else { a_t[v] = s[v]; xt_f[v] = xv[v]; }
          if (kind == 0) {
            if (oe & (1u << 10)) { rmax[v] = -1.f; rhold[v] = 0u; }
            if (s[v] > rmax[v]) { rmax[v] = s[v]; rhold[v] = (unsigned)ch; }     // lowest fine id wins ties (:84-85)
          }
        }
        if (((++nF) & 3) == 0) {   // one log per <= 4 factors: each factor >= eps = 1e-8
#pragma unroll
          for (int v = 0; v < 2; ++v) { lf[v] -= fast_log(prodF[v]); prodF[v] = 1.f; }
        }
      } else {
        unsigned int hd = 0;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          sumC[v] += E[v];
          float best = (oe & (1u << 10)) ? -1.f : rmax[v];
          unsigned int hold = (oe & (1u << 10)) ? 0u : rhold[v];
          if (s[v] > best) { best = s[v]; hold = (unsigned)ch; }                 // the coarse logit comes last in the cat
          hd |= hold << (8 * v);
          if (bucket != tc[v]) prodC[v] *= (1.0f - best) + eps;
          else {
            b_t[v] = s[v]; xt_c[v] = xv[v];
            if (tf[v] != SH_IGNORE && !(a_t[v] <= s[v])) hold_pos_a[v] = false;  // fine wins ties (:91-92)
          }
        }
        HOLD[(size_t)bucket * F2_NT + tid] = (unsigned short)hd;
        if (((++nC) & 3) == 0) {
#pragma unroll
          for (int v = 0; v < 2; ++v) { lc[v] -= fast_log(prodC[v]); prodC[v] = 1.f; }
        }
      }
    }
    float inv_f[2], inv_c[2];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      lf[v] -= fast_log(prodF[v]);
      lc[v] -= fast_log(prodC[v]);
      const bool vf = tf[v] != SH_IGNORE, vc = tc[v] != SH_IGNORE;
      if (vf) {
        const float m = (vc && !hold_pos_a[v]) ? b_t[v] : a_t[v];
        acc[0] += lf[v] - fast_log(m + eps);
        acc[2] += fast_log(sumF[v]) - xt_f[v];
      }
      if (vc) {
        acc[1] += lc[v] - fast_log(b_t[v] + eps);
        acc[3] += fast_log(sumC[v]) - xt_c[v];
      }
      inv_f[v] = rcp(sumF[v]);
      inv_c[v] = rcp(sumC[v]);
    }

    // ---- sweep B: gradient, written once ------------------------------------------------------
    if (GRAD) {
      T* gb = grad + (long)b * C * HW + p0;
#pragma unroll 2
      for (int c = 0; c < C; ++c) {
        const float2 s2 = S[(size_t)c * F2_NT + tid], e2 = V[(size_t)c * F2_NT + tid];
        const float s[2] = {s2.x, s2.y}, ev[2] = {e2.x, e2.y};
        const bool fine = c < h.nf;
        const int i = fine ? s_owner[c] : c - h.nf;
        const unsigned int hd = i >= 0 ? (unsigned int)HOLD[(size_t)i * F2_NT + tid] : 0xffffu;
        float g[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const bool vf = tf[v] != SH_IGNORE, vc = tc[v] != SH_IGNORE;
          const float q = 1.0f - s[v];
          const float rneg = rcp(q + eps);
          const bool holds = ((hd >> (8 * v)) & 0xffu) == (unsigned)c;
          float ds = 0.f, ce = 0.f;
          if (fine) {
            const bool isT = c == tf[v];
            if (vf && !isT) ds = wF * rneg;
            if (vc && i >= 0 && i != tc[v] && holds) ds = fmaf(wC, rneg, ds);
            if (vf && isT && !(vc && !hold_pos_a[v])) ds -= wF * rcp(s[v] + eps);
            if (vf) ce = wCE * (ev[v] * inv_f[v] - (isT ? 1.f : 0.f));
          } else {
            const bool isT = i == tc[v];
            if (vc) {
              if (isT) ds = -wC * rcp(s[v] + eps);
              else if (holds) ds = wC * rneg;
              ce = wCE * (ev[v] * inv_c[v] - (isT ? 1.f : 0.f));
            }
            if (vf && vc && isT && !hold_pos_a[v]) ds -= wF * rcp(s[v] + eps);
          }
          g[v] = fmaf(ds, q * s[v], ce);
        }
        if (inb) VecIO<T, 2>::store(gb + (long)c * HW, g);
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
  size_t s = (size_t)C * F2_NT * 8 * 2 + (size_t)nc * F2_NT * 2 + 16;
  s += (size_t)F2_XD * F2_NT * 8 + (size_t)C * 4 + (size_t)nf * 4 + (size_t)lut_size * 4 + 32;
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

template <typename T>
static int launch_bce2(const void* x, const unsigned char* lab8, void* grad, int B, long HW, const Hier2& h, float eps,
                       float lw, const unsigned long long* counts, float* partials, int grid, bool tree, cudaStream_t st) {
  const size_t fsm = bce2_fast_smem(h.nf, h.nc, h.lut_size);
  if (tree && HW % 2 == 0 && (uintptr_t)x % (2 * sizeof(T)) == 0 && (uintptr_t)grad % (2 * sizeof(T)) == 0 &&
      fsm <= 113 * 1024) {
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
  constexpr int VEC = 2;
  constexpr int PX = 128 * VEC;
  const int C = h.nf + h.nc;
  size_t smem = (size_t)C * PX * 8 + (size_t)h.nc * PX;
  if (smem > 227 * 1024) return SH_ERR_UNSUPPORTED;
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
  size_t smem = (size_t)C * PX * 8 + (size_t)n_coarse * PX + (size_t)sh::F2_XD * sh::F2_NT * 8 + 2048;
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
int sh_bce2_fwdbwd(const void* logits, int dtype, const long long* label, void* grad /* nullable */, int B, long HW,
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
    sh::k_prep2<<<(unsigned)pb, 256, 0, st>>>(label, lab8, n, n_fine, h.lut, lut_size, counts);
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
