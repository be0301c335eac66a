// Target builders (integer gathers) and the hierarchical argmax decode.
// HBM-bound streaming kernels; all results are bit-exact integers.
#include "common.cuh"

namespace sh {

// ---------------------------------------------------------------------------
// Target builders.  Reference: models/loss/hiera_triplet_loss.py:11-38 (range
// test, later bucket wins -> precomputed LUT), models/loss/rmi_hiera_triplet_loss.py
// :21-63 (gather with literal-255 passthrough), dataset/dataloader.py:166-177
// (plain gather).  torch advanced indexing wraps negative indices, so do we.
// ---------------------------------------------------------------------------
// One thread = 16 bytes of labels (2 int64 / 4 int32 / 16 uint8) in, the same out; the scalar form takes tensors
// that are not 16-byte aligned (e.g. a batch slice of an odd-sized image) and the tail.
template <typename L, bool VEC>
__global__ void __launch_bounds__(256) k_targets_two_level(const L* __restrict__ label, L* __restrict__ coarse, long n,
                                                           const int* __restrict__ lut, int lut_size) {
  constexpr int N = VEC ? 16 / (int)sizeof(L) : 1;
  long i = (blockIdx.x * (long)blockDim.x + threadIdx.x) * N;
  const long stride = (long)gridDim.x * blockDim.x * N;
  for (; i < n; i += stride) {
    if (VEC && i + N <= n) {
      long long t[16 / sizeof(L)];
      lab_ld16<L>(label + i, t);
#pragma unroll
      for (int k = 0; k < N; ++k) t[k] = (t[k] >= 0 && t[k] < lut_size) ? lut[t[k]] : SH_IGNORE;
      lab_st16<L>(coarse + i, t);
    } else {
      for (long j = i; j < min(i + N, n); ++j) {
        const long long t = lab_ld(label, j);
        coarse[j] = (L)((t >= 0 && t < lut_size) ? lut[t] : SH_IGNORE);
      }
    }
  }
}

// mode 0: three-level builder (255 passes through, two maps); mode 1: dataloader gather (one map, no passthrough).
// Same vector form; the maps (a few hundred int64 at most) are copied to shared memory first.
template <typename L, bool VEC>
__global__ void __launch_bounds__(256) k_targets_gather(const L* __restrict__ label, L* __restrict__ out_a,
                                                        L* __restrict__ out_b, long n,
                                                        const long long* __restrict__ map_a,
                                                        const long long* __restrict__ map_b, int map_size,
                                                        int passthrough_255, int* __restrict__ err) {
  extern __shared__ long long s_map[];          // [map_size] a, then [map_size] b
  for (int i = threadIdx.x; i < map_size; i += blockDim.x) {
    s_map[i] = map_a[i];
    if (map_b) s_map[map_size + i] = map_b[i];
  }
  __syncthreads();
  constexpr int N = VEC ? 16 / (int)sizeof(L) : 1;
  long i = (blockIdx.x * (long)blockDim.x + threadIdx.x) * N;
  const long stride = (long)gridDim.x * blockDim.x * N;
  bool bad = false;
  auto one = [&](long long t, long long& a, long long& b) {
    a = SH_IGNORE; b = SH_IGNORE;
    if (!(passthrough_255 && t == SH_IGNORE)) {
      const long long u = t < 0 ? t + map_size : t;       // torch advanced indexing wraps negative indices
      if (u >= 0 && u < map_size) {
        a = s_map[u];
        if (map_b) b = s_map[map_size + u];
      } else {
        bad = true;
      }
    }
  };
  for (; i < n; i += stride) {
    if (VEC && i + N <= n) {
      long long t[16 / sizeof(L)], a[16 / sizeof(L)], b[16 / sizeof(L)];
      lab_ld16<L>(label + i, t);
#pragma unroll
      for (int k = 0; k < N; ++k) one(t[k], a[k], b[k]);
      lab_st16<L>(out_a + i, a);
      if (out_b) lab_st16<L>(out_b + i, b);
    } else {
      for (long j = i; j < min(i + N, n); ++j) {
        long long a, b;
        one(lab_ld(label, j), a, b);
        out_a[j] = (L)a;
        if (out_b) out_b[j] = (L)b;
      }
    }
  }
  if (bad) atomicOr(err, 1);
}

// Class-id mask -> RGB image (infer.py:117-131, a per-pixel Python loop in the reference): negative ids are black,
// ids >= n_colors raise IndexError there and set *err here.  One thread = 4 pixels = 12 output bytes (3 words).
template <typename L>
__global__ void __launch_bounds__(256) k_colorize(const L* __restrict__ mask, long n, const unsigned char* __restrict__ pal,
                                                  int n_colors, unsigned char* __restrict__ rgb, int* __restrict__ err) {
  extern __shared__ unsigned int s_pal[];       // packed r | g << 8 | b << 16
  for (int i = threadIdx.x; i < n_colors; i += blockDim.x)
    s_pal[i] = (unsigned int)pal[3 * i] | ((unsigned int)pal[3 * i + 1] << 8) | ((unsigned int)pal[3 * i + 2] << 16);
  __syncthreads();
  bool bad = false;
  const bool al = ((uintptr_t)rgb & 3) == 0;
  for (long q = blockIdx.x * (long)blockDim.x + threadIdx.x; 4 * q < n; q += (long)gridDim.x * blockDim.x) {
    unsigned int c[4] = {0u, 0u, 0u, 0u};
    const int m = (int)min(4L, n - 4 * q);
    for (int k = 0; k < m; ++k) {
      const long long t = lab_ld(mask, 4 * q + k);
      if (t >= n_colors) bad = true;
      else if (t >= 0) c[k] = s_pal[t];
    }
    if (m == 4 && al) {
      unsigned int* o = reinterpret_cast<unsigned int*>(rgb + 12 * q);
      o[0] = c[0] | (c[1] << 24);
      o[1] = (c[1] >> 8) | (c[2] << 16);
      o[2] = (c[2] >> 16) | (c[3] << 8);
    } else {
      for (int k = 0; k < m; ++k) {
        rgb[12 * q + 3 * k] = (unsigned char)c[k];
        rgb[12 * q + 3 * k + 1] = (unsigned char)(c[k] >> 8);
        rgb[12 * q + 3 * k + 2] = (unsigned char)(c[k] >> 16);
      }
    }
  }
  if (bad) atomicOr(err, 1);
}

// ---------------------------------------------------------------------------
// Decode: independent per-level argmax over channel slices (infer.py:303-312),
// first max wins, NaN counts as max (torch.argmax).  Optional fine-level pixel
// accuracy counts (train.py:37-49, 382-385).  Logits are read exactly once.
// ---------------------------------------------------------------------------
template <typename T, int VEC, typename OutT, typename L>
__global__ void __launch_bounds__(256) k_decode(const T* __restrict__ x, int B, int C, long HW, int n0, int n1, int n2,
                                                OutT* __restrict__ o0, OutT* __restrict__ o1, OutT* __restrict__ o2,
                                                const L* __restrict__ label,
                                                unsigned long long* __restrict__ counts, int vec_ok) {
  const long groups_per_img = (HW + VEC - 1) / VEC;
  const long total = groups_per_img * B;
  long long correct = 0, valid = 0;
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
    const int b = (int)(g / groups_per_img);
    const long p = (g - (long)b * groups_per_img) * VEC;
    const T* base = x + (long)b * C * HW;
    int c0 = 0;
#pragma unroll
    for (int lvl = 0; lvl < 3; ++lvl) {
      const int k = lvl == 0 ? n0 : (lvl == 1 ? n1 : n2);
      OutT* out = lvl == 0 ? o0 : (lvl == 1 ? o1 : o2);
      if (k <= 0 || out == nullptr) { c0 += max(k, 0); continue; }
      float best[VEC];
      int arg[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) { best[v] = 0.f; arg[v] = -1; }
#pragma unroll 4
      for (int c = 0; c < k; ++c) {
        float val[VEC];
        load_n<T, VEC>(base + (long)(c0 + c) * HW, p, HW, vec_ok != 0, val);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          // take if first, strictly greater, or first NaN
          bool take = (arg[v] < 0) || (val[v] > best[v]) || (val[v] != val[v] && best[v] == best[v]);
          if (take) { best[v] = val[v]; arg[v] = c; }
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (p + v < HW) {
          out[(long)b * HW + p + v] = (OutT)arg[v];
          if (lvl == 0 && label != nullptr) {
            const long long t = lab_ld(label, (long)b * HW + p + v);
            if (t != SH_IGNORE) { valid++; correct += (t == arg[v]); }
          }
        }
      }
      c0 += k;
    }
  }
  if (label != nullptr && counts != nullptr) {
    correct = warp_sum(correct);
    valid = warp_sum(valid);
    if ((threadIdx.x & 31) == 0 && valid) {
      atomicAdd(counts, (unsigned long long)correct);
      atomicAdd(counts + 1, (unsigned long long)valid);
    }
  }
}

// Vector path (HW % VEC == 0, 16-byte aligned logits): the channels are walked in batches of DEPTH raw 128-bit loads
// that are all in flight before the first compare (the kernel is bound by load latency, not by arithmetic), level
// boundaries are handled inside the walk, int64 results leave as 16-byte stores.
template <typename T, int VEC>
__device__ __forceinline__ void raw_to_f32(const uint4& r, float (&o)[VEC]);
template <>
__device__ __forceinline__ void raw_to_f32<float, 4>(const uint4& r, float (&o)[4]) {
  o[0] = __uint_as_float(r.x); o[1] = __uint_as_float(r.y); o[2] = __uint_as_float(r.z); o[3] = __uint_as_float(r.w);
}
template <>
__device__ __forceinline__ void raw_to_f32<__nv_bfloat16, 8>(const uint4& r, float (&o)[8]) {
  const unsigned int w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) { o[2 * k] = __uint_as_float(w[k] << 16); o[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u); }
}
template <>
__device__ __forceinline__ void raw_to_f32<__half, 8>(const uint4& r, float (&o)[8]) {
  const unsigned int w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
    o[2 * k] = f.x; o[2 * k + 1] = f.y;
  }
}

template <int VEC, typename OutT>
__device__ __forceinline__ void store_args(OutT* dst, const int (&arg)[VEC]);
template <>
__device__ __forceinline__ void store_args<4, long long>(long long* dst, const int (&arg)[4]) {
  __stcs(reinterpret_cast<longlong2*>(dst), make_longlong2(arg[0], arg[1]));
  __stcs(reinterpret_cast<longlong2*>(dst) + 1, make_longlong2(arg[2], arg[3]));
}
template <>
__device__ __forceinline__ void store_args<8, long long>(long long* dst, const int (&arg)[8]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) __stcs(reinterpret_cast<longlong2*>(dst) + q, make_longlong2(arg[2 * q], arg[2 * q + 1]));
}
template <>
__device__ __forceinline__ void store_args<4, unsigned char>(unsigned char* dst, const int (&arg)[4]) {
  *reinterpret_cast<unsigned int*>(dst) = (unsigned)arg[0] | ((unsigned)arg[1] << 8) | ((unsigned)arg[2] << 16) | ((unsigned)arg[3] << 24);
}
template <>
__device__ __forceinline__ void store_args<8, unsigned char>(unsigned char* dst, const int (&arg)[8]) {
  uint2 v;
  v.x = (unsigned)arg[0] | ((unsigned)arg[1] << 8) | ((unsigned)arg[2] << 16) | ((unsigned)arg[3] << 24);
  v.y = (unsigned)arg[4] | ((unsigned)arg[5] << 8) | ((unsigned)arg[6] << 16) | ((unsigned)arg[7] << 24);
  *reinterpret_cast<uint2*>(dst) = v;
}

// 16-bit logits: two pixels per 32-bit word, compared as packed pairs (HSET2 masks + bitwise selects): 5 instructions per
// pair where the fp32 path spends ~8 per pixel.  Same decision as argmax over torch's tensors: take the new channel when
// the running best is not NaN and NOT (value <= best) -- strictly greater, or the first NaN.
template <typename T> __device__ __forceinline__ unsigned int le2_mask(unsigned int a, unsigned int b);
template <> __device__ __forceinline__ unsigned int le2_mask<__nv_bfloat16>(unsigned int a, unsigned int b) {
  return __hle2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
}
template <> __device__ __forceinline__ unsigned int le2_mask<__half>(unsigned int a, unsigned int b) {
  return __hle2_mask(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
}
template <> __device__ __forceinline__ unsigned int le2_mask<float>(unsigned int, unsigned int) { return 0u; }
template <typename T> __device__ __forceinline__ unsigned int ord2_mask(unsigned int a);      // 0xffff where the half is not NaN
template <> __device__ __forceinline__ unsigned int ord2_mask<__nv_bfloat16>(unsigned int a) {
  return __heq2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&a));
}
template <> __device__ __forceinline__ unsigned int ord2_mask<__half>(unsigned int a) {
  return __heq2_mask(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&a));
}
template <> __device__ __forceinline__ unsigned int ord2_mask<float>(unsigned int) { return 0u; }

template <typename T, int VEC, typename OutT, typename L>
__global__ void __launch_bounds__(256, 3) k_decode_vec(const T* __restrict__ x, int B, int C, long HW, int n0, int n1, int n2,
                                                    OutT* __restrict__ o0, OutT* __restrict__ o1, OutT* __restrict__ o2,
                                                    const L* __restrict__ label,
                                                    unsigned long long* __restrict__ counts) {
  constexpr int DEPTH = 8;
  const long groups_per_img = HW / VEC;
  const long total = groups_per_img * B;
  const int e0 = n0, e1 = n0 + max(n1, 0), e2 = e1 + max(n2, 0);     // level ends in channel units
  long long correct = 0, valid = 0;
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
    const int b = (int)(g / groups_per_img);
    const long p = (g - (long)b * groups_per_img) * VEC;
    const char* base = reinterpret_cast<const char*>(x + (long)b * C * HW + p);
    const long cstride = HW * (long)sizeof(T);
    float best[VEC];
    int arg[VEC];
    unsigned int best2[4], arg2[4];      // 16-bit logits: packed pairs (value bits, channel index per half)
    int lvl = 0, cbeg = 0;
#pragma unroll 1
    for (int cb = 0; cb < e2; cb += DEPTH) {
      uint4 raw[DEPTH];
#pragma unroll
      for (int j = 0; j < DEPTH; ++j)
        if (cb + j < e2) raw[j] = __ldcs(reinterpret_cast<const uint4*>(base + (long)(cb + j) * cstride));
#pragma unroll
      for (int j = 0; j < DEPTH; ++j) {
        const int c = cb + j;
        if (c >= e2) break;
        if (sizeof(T) == 2) {
          const unsigned int w[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
          if (c == cbeg) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { best2[q] = w[q]; arg2[q] = 0u; }
          } else {
            const unsigned int cc = (unsigned int)(c - cbeg) * 0x00010001u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const unsigned int take = ~le2_mask<T>(w[q], best2[q]) & ord2_mask<T>(best2[q]);
              best2[q] = (best2[q] & ~take) | (w[q] & take);
              arg2[q] = (arg2[q] & ~take) | (cc & take);
            }
          }
        } else {
          float val[VEC];
          raw_to_f32<T, VEC>(raw[j], val);
          if (c == cbeg) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { best[v] = val[v]; arg[v] = 0; }
          } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              // strictly greater, or first NaN (torch.argmax treats NaN as the maximum)
              const bool take = (val[v] > best[v]) || (val[v] != val[v] && best[v] == best[v]);
              if (take) { best[v] = val[v]; arg[v] = c - cbeg; }
            }
          }
        }
        const int lend = lvl == 0 ? e0 : (lvl == 1 ? e1 : e2);
        if (c + 1 == lend) {            // last channel of the level: results out, next level starts
          if (sizeof(T) == 2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { arg[(2 * q) % VEC] = (int)(arg2[q] & 0xffffu); arg[(2 * q + 1) % VEC] = (int)(arg2[q] >> 16); }
          }
          OutT* out = lvl == 0 ? o0 : (lvl == 1 ? o1 : o2);
          if (out != nullptr) store_args<VEC, OutT>(out + (long)b * HW + p, arg);
          if (lvl == 0 && label != nullptr) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              const long long t = lab_ld(label, (long)b * HW + p + v);
              if (t != SH_IGNORE) { valid++; correct += (t == arg[v]); }
            }
          }
          cbeg = lend;
          ++lvl;
          while (lvl < 3 && (lvl == 1 ? e1 : e2) == cbeg) ++lvl;     // empty levels
        }
      }
    }
  }
  if (label != nullptr && counts != nullptr) {
    correct = warp_sum(correct);
    valid = warp_sum(valid);
    if ((threadIdx.x & 31) == 0 && valid) {
      atomicAdd(counts, (unsigned long long)correct);
      atomicAdd(counts + 1, (unsigned long long)valid);
    }
  }
}

template <typename T, typename OutT, typename L>
static int launch_decode(const void* x, int B, int C, long HW, int n0, int n1, int n2, void* o0, void* o1, void* o2,
                         const L* label, unsigned long long* counts, cudaStream_t st) {
  constexpr int VEC = sizeof(T) == 4 ? 4 : 8;  // 128-bit loads
  bool vec_ok = (HW % VEC == 0) && ((uintptr_t)x % (VEC * sizeof(T)) == 0);
  long groups = ((HW + VEC - 1) / VEC) * B;
  long blocks = (groups + 255) / 256;
  if (blocks > SH_NUM_SMS * 16L) blocks = SH_NUM_SMS * 16L;
  if (blocks < 1) blocks = 1;
  const bool out_al = ((uintptr_t)o0 | (uintptr_t)o1 | (uintptr_t)o2) % 16 == 0;
  if (vec_ok && out_al && n0 > 0) {
    k_decode_vec<T, VEC, OutT, L><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, B, C, HW, n0, n1, n2, (OutT*)o0, (OutT*)o1,
                                                                  (OutT*)o2, label, counts);
    SH_CHECK_LAUNCH();
    return SH_OK;
  }
  k_decode<T, VEC, OutT, L><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, B, C, HW, n0, n1, n2, (OutT*)o0, (OutT*)o1,
                                                            (OutT*)o2, label, counts, vec_ok ? 1 : 0);
  SH_CHECK_LAUNCH();
  return SH_OK;
}

}  // namespace sh

extern "C" {

static long sh_stream_blocks(long items) {
  long blocks = (items + 255) / 256;
  if (blocks > SH_NUM_SMS * 16L) blocks = SH_NUM_SMS * 16L;
  return blocks < 1 ? 1 : blocks;
}

int sh_targets_two_level(const void* label, int label_dtype, void* coarse, long n, const int* lut, int lut_size,
                         void* stream) {
  if (n <= 0) return SH_OK;
  const bool al = (((uintptr_t)label | (uintptr_t)coarse) % 16) == 0;
  SH_LABEL_SWITCH(label_dtype, L, {
    constexpr int N = 16 / (int)sizeof(L);
    if (al)
      sh::k_targets_two_level<L, true><<<(unsigned)sh_stream_blocks((n + N - 1) / N), 256, 0, (cudaStream_t)stream>>>(
          (const L*)label, (L*)coarse, n, lut, lut_size);
    else
      sh::k_targets_two_level<L, false><<<(unsigned)sh_stream_blocks(n), 256, 0, (cudaStream_t)stream>>>(
          (const L*)label, (L*)coarse, n, lut, lut_size);
  })
  SH_CHECK_LAUNCH();
  return SH_OK;
}

static int sh_gather_launch(const void* label, int label_dtype, void* out_a, void* out_b, long n, const long long* map_a,
                            const long long* map_b, int map_size, int passthrough, int* err_flag, cudaStream_t st) {
  if (map_size <= 0 || map_size > 2048) return SH_ERR_BAD_ARG;
  const bool al = (((uintptr_t)label | (uintptr_t)out_a | (uintptr_t)out_b) % 16) == 0;
  const size_t smem = (size_t)map_size * 8 * (map_b ? 2 : 1);
  SH_LABEL_SWITCH(label_dtype, L, {
    constexpr int N = 16 / (int)sizeof(L);
    if (al)
      sh::k_targets_gather<L, true><<<(unsigned)sh_stream_blocks((n + N - 1) / N), 256, smem, st>>>(
          (const L*)label, (L*)out_a, (L*)out_b, n, map_a, map_b, map_size, passthrough, err_flag);
    else
      sh::k_targets_gather<L, false><<<(unsigned)sh_stream_blocks(n), 256, smem, st>>>(
          (const L*)label, (L*)out_a, (L*)out_b, n, map_a, map_b, map_size, passthrough, err_flag);
  })
  SH_CHECK_LAUNCH();
  return SH_OK;
}

int sh_targets_three_level(const void* label, int label_dtype, void* mid, void* high, long n, const long long* f2m,
                           const long long* f2h, int n_fine, int* err_flag, void* stream) {
  if (n <= 0) return SH_OK;
  return sh_gather_launch(label, label_dtype, mid, high, n, f2m, f2h, n_fine, 1, err_flag, (cudaStream_t)stream);
}

int sh_targets_gather(const void* label, int label_dtype, void* out, long n, const long long* map, int map_size,
                      int* err_flag, void* stream) {
  if (n <= 0) return SH_OK;
  return sh_gather_launch(label, label_dtype, out, nullptr, n, map, nullptr, map_size, 0, err_flag, (cudaStream_t)stream);
}

int sh_colorize(const void* mask, int label_dtype, long n, const unsigned char* palette, int n_colors,
                unsigned char* rgb, int* err_flag, void* stream) {
  if (n <= 0) return SH_OK;
  if (n_colors <= 0 || n_colors > 4096) return SH_ERR_BAD_ARG;
  SH_LABEL_SWITCH(label_dtype, L, {
    sh::k_colorize<L><<<(unsigned)sh_stream_blocks((n + 3) / 4), 256, (size_t)n_colors * 4, (cudaStream_t)stream>>>(
        (const L*)mask, n, palette, n_colors, rgb, err_flag);
  })
  SH_CHECK_LAUNCH();
  return SH_OK;
}

int sh_decode(const void* logits, int dtype, int B, int C, long HW, int n0, int n1, int n2, void* out0, void* out1,
              void* out2, int out_is_u8, const void* label, int label_dtype, unsigned long long* counts, void* stream) {
  if (B <= 0 || HW <= 0) return SH_OK;
  if (n0 + (n1 > 0 ? n1 : 0) + (n2 > 0 ? n2 : 0) > C) return SH_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
#define SH_DECODE_DISPATCH(T)                                                                                        \
  SH_LABEL_SWITCH(label_dtype, L, {                                                                                  \
    return out_is_u8 ? sh::launch_decode<T, unsigned char, L>(logits, B, C, HW, n0, n1, n2, out0, out1, out2,        \
                                                              (const L*)label, counts, st)                           \
                     : sh::launch_decode<T, long long, L>(logits, B, C, HW, n0, n1, n2, out0, out1, out2,            \
                                                          (const L*)label, counts, st);                              \
  })                                                                                                                 \
  break
  switch (dtype) {
    case SH_DT_F32: SH_DECODE_DISPATCH(float);
    case SH_DT_BF16: SH_DECODE_DISPATCH(__nv_bfloat16);
    case SH_DT_F16: SH_DECODE_DISPATCH(__half);
  }
#undef SH_DECODE_DISPATCH
  return SH_ERR_UNSUPPORTED;
}

}  // extern "C"
