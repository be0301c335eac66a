/* seghiero_b200 -- C ABI of the B200-native hierarchical-loss path.
 *
 * Drop-in boundary for SegHiero's `models/loss` (Python nn.Modules; the reference has no FFI of
 * its own).  Every entry point takes plain device pointers, sizes and a
 * cudaStream_t passed as void*; nothing here names a torch type.  All pointers are caller-owned
 * device memory (PyTorch's caching allocator in the shipped host code); the library keeps no global
 * state, allocates nothing and launches only on the given stream.  `stages` arguments are bit masks
 * selecting which kernels of a multi-kernel entry point run (-1 = all; bench.py brackets single kernels
 * with CUDA events this way).  Return value: 0 on success,
 * a negative SH_ERR_* code for bad arguments, or a positive cudaError_t from the launch.
 *
 * dtype codes: 0 = float32, 1 = bfloat16, 2 = float16 (logits / embeddings / gradients).
 * label_dtype codes: 0 = int64 (what the reference passes), 1 = int32, 2 = uint8 (1 byte per pixel instead of 8 from
 * the dataloader to the loss: SURVEY section 8f row N4); 255 is the ignore label in every type.
 * Paths below are relative to the reference root.
 */
#ifndef SEGHIERO_B200_H
#define SEGHIERO_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- target builders (bit-exact integer gathers) ------------------------------------------- */

/* Replaces _prepare_targets_two_level, models/loss/hiera_triplet_loss.py:11-38.
 * lut[t] = last bucket [start,end) containing t, else 255; labels outside [0,lut_size) -> 255.
 * Outputs of the three builders have the element type of the input labels. */
int sh_targets_two_level(const void* label, int label_dtype, void* coarse, long n, const int* lut, int lut_size,
                         void* stream);

/* Replaces _prepare_targets_three_level, models/loss/rmi_hiera_triplet_loss.py:21-63.
 * mid = f2m[t], high = f2h[t] where t != 255 (negative t wraps like torch indexing); out-of-range
 * labels set *err_flag (the reference raises IndexError). */
int sh_targets_three_level(const void* label, int label_dtype, void* mid, void* high, long n, const long long* f2m,
                           const long long* f2h, int n_fine, int* err_flag, void* stream);

/* Replaces the dataloader gather `map[fine_mask]`, dataset/dataloader.py:166-177 (no ignore handling). */
int sh_targets_gather(const void* label, int label_dtype, void* out, long n, const long long* map, int map_size,
                      int* err_flag, void* stream);

/* Replaces mask_to_color_image, infer.py:117-131 (a per-pixel interpreted loop there): class-id mask [n] ->
 * RGB bytes [n][3].  palette = n_colors x 3 bytes; negative ids are black; ids >= n_colors (IndexError in the
 * reference) set *err_flag. */
int sh_colorize(const void* mask, int label_dtype, long n, const unsigned char* palette, int n_colors,
                unsigned char* rgb, int* err_flag, void* stream);

/* ---- decode --------------------------------------------------------------------------------- */

/* Replaces the per-level argmax of infer.py:303-312 and the pixel-accuracy counts of
 * train.py:37-49, 382-385.  logits [B,C,HW]; levels are the channel slices [0,n0), [n0,n0+n1),
 * [n0+n1,n0+n1+n2) (n1/n2 may be 0).  Outputs int64 (or uint8 when out_is_u8) [B,HW]; first max
 * wins, NaN counts as max.  If label != NULL, counts[0] += #correct fine, counts[1] += #valid. */
int sh_decode(const void* logits, int dtype, int B, int C, long HW, int n0, int n1, int n2, void* out0, void* out1,
              void* out2, int out_is_u8, const void* label, int label_dtype, unsigned long long* counts, void* stream);

/* ---- two-level loss: HieraTripletLoss.forward, models/loss/hiera_triplet_loss.py:152-211 ---- */

/* CTAs launched by sh_bce2_fwdbwd = rows (4 floats each) of its `partials` workspace. */
int sh_bce2_grid(int B, long HW, int C, int n_coarse);

/* Fused target-derive + tree BCE (hiera_triplet_loss.py:41-107) + per-level softmax CE
 * (cross_entropy_loss.py:7-30), forward and d/dlogits in one pass.  grad may be NULL (forward only);
 * it is the gradient for grad_output == 1, already scaled by loss_weight.
 * hier_tab (device int32): [bucket_start nc][bucket_end nc][owner nf][fb_ptr nf+1][fb_idx n_fb][lut lut_size].
 * Outputs: sums[0..3] = un-normalised BCE fine, BCE coarse, CE fine, CE coarse; counts[0..2] = #valid fine,
 * #valid coarse, label-range error flag.
 * stages: bit 0 = label prep, bit 1 = fused loss kernel, bit 2 = reduction of the per-CTA partials; bit 8 (256)
 * is a hint from the host table builder that the buckets are disjoint ranges (no fine class in two buckets,
 * hierarchy.py::two_level_is_tree): the tree-order kernel k_bce2_fast runs then, else the any-bucket kernel. */
int sh_bce2_fwdbwd(const void* logits, int dtype, const void* label, int label_dtype, void* grad, int B, long HW, int n_fine,
                   int n_coarse, const int* hier_tab, int n_fb, int lut_size, float eps, float loss_weight,
                   unsigned char* lab8, unsigned long long* counts, float* partials, double* sums, int stages,
                   void* stream);

/* Scalar assembly incl. the cosine schedule and ready gate (hiera_triplet_loss.py:188-211), on device.
 * out[0] = loss, out[1] = ready*factor*loss_weight (scale for the triplet backward). */
int sh_loss2_final(const double* sums, const unsigned long long* counts, int n_fine, int n_coarse, double npx,
                   const double* step, double total_steps, const float* trip, const int* ready, float loss_weight,
                   float* out, void* stream);

/* grad *= *scale unless *scale == 1 (non-unit grad_output from autograd). */
int sh_scale_inplace(void* grad, int dtype, long n, const float* scale, void* stream);

/* ---- three-level loss: RMIHieraTripletLoss.forward, models/loss/rmi_hiera_triplet_loss.py:323-546 */

size_t sh_rmi3_workspace_bytes(int B, int H, int W, int nf, int nm, int nh);

/* Byte offsets of the workspace fields, for diagnostics and tests: out[0..11] = counts, sums, lab8, flags,
 * hold, inv, part1, bcepart, frameT, rbc, wts, fwts. */
int sh_rmi3_workspace_offsets(int B, int H, int W, int nf, int nm, int nh, size_t* out);

/* 1 when the warp-specialised kernels (csrc/rmi3_fast.cuh) will run this problem: tree-shaped maps
 * (fast_tab_ok from the host table builder), W % 4 == 0, 16-byte aligned tensors, 7 < C <= 254.  Everything
 * else runs the generic kernels; results agree to fp32 rounding. */
int sh_rmi3_fast_path(const void* logits, const void* grad, int dtype, int H, int W, int nf, int nm, int nh,
                      int fast_tab_ok);

/* Pass 1 (+ label prep, frame taps, per-(b,c) 9x9 algebra): tree BCE (rmi...py:352-470), CE (:523-526),
 * RMI lower bound (:479-517) without materialising the unfolds.  Leaves everything the backward pass and
 * sh_loss3_final need in `workspace`.
 * hier_tab (device int32): [f2m nf][f2h nf][mh_ptr nm+1][mh_idx n_mh][hsmask nm][order C][fast_order C][fast_aux C]
 * (seghiero_b200/hierarchy.py::three_level_tables). */
int sh_rmi3_forward(const void* logits, int dtype, const void* label, int label_dtype, int B, int H, int W, int nf, int nm, int nh,
                    const int* hier_tab, int n_mh, int fast_tab_ok, float lam, float loss_weight, void* workspace,
                    int stages, void* stream);

/* out[0] = loss, out[1] = ready*factor*loss_weight, out[2] = rmi term. */
int sh_loss3_final(int B, int H, int W, int nf, int nm, int nh, void* workspace, float lam, const double* step,
                   double total_steps, const float* trip, const int* ready, float loss_weight, float* out,
                   void* stream);

/* Which pass-2 kernel sh_rmi3_backward runs for this problem (host arithmetic only): 0 = generic, 1 = tiled kernel with
 * per-thread cp.async staging (csrc/rmi3_fast_bwd.cuh), 2 = tiled kernel with TMA box loads (csrc/rmi3_fast_bwd2.cuh:
 * needs W % 16 == 0 for the holder-byte tensor map). */
int sh_rmi3_pass2_kind(const void* logits, const void* grad, int dtype, int H, int W, int nf, int nm, int nh,
                       int fast_tab_ok);

/* Pass 2: d loss / d logits written once (grad_out = device scalar handed over by autograd). */
int sh_rmi3_backward(const void* logits, int dtype, void* grad, int B, int H, int W, int nf, int nm, int nh,
                     const int* hier_tab, int n_mh, int fast_tab_ok, float loss_weight, void* workspace,
                     const float* grad_out, int stages, void* stream);

/* ---- triplet: TreeTripletLoss.forward, models/loss/tree_triplet_loss.py:15-65 (mode 0) and
 *      models/loss/rmi_tree_triplet_loss.py:14-70 (mode 1) ------------------------------------- */

/* feats [B,D,h,w]; label [B,H,W] is nearest-downsampled to (h,w) inside.
 * tab: mode 0 -> [bucket_lo ncls][bucket_hi ncls]; mode 1 -> group per label value [256] (0/1/-1), ncls = 256.
 * Workspaces: lab_ds [B*h*w] int32, sel [ncls*3*max_triplet] int32, kcount [ncls] int32, tl [ncls*max_triplet] f32.
 * Outputs: trip[0] = loss, trip[1] = #contributing classes; status[0] = ready, status[1] = label error. */
int sh_triplet_forward(const void* feats, int dtype, const void* label, int label_dtype, int B, int D, int h, int w, int H, int W,
                       int mode, const int* tab, int ncls, int max_triplet, int* lab_ds, int* sel, int* kcount,
                       float* tl, float* trip, int* status, void* stream);

/* gfeat: fp32 [B,D,h,w], fully written; weight = *tscale * *gscale / (k_c * #classes).  scratch: B*D*h*w*8 bytes
 * (zeroed here): the scatter accumulates in 64-bit fixed point with integer atomics, so the result is bitwise
 * reproducible. */
int sh_triplet_backward(const void* feats, int dtype, int B, int D, int h, int w, int ncls, int max_triplet,
                        const int* sel, const int* kcount, const float* tl, const float* trip, const float* tscale,
                        const float* gscale, float* gfeat, void* scratch, void* stream);

/* ---- upsample-fused variants: the head's logits stay at their own resolution ---- */

/* F.interpolate(x, size=(H, W), mode="bilinear", align_corners=False) of train.py:282-284 for `planes` = B*C planes
 * [h, w] -> [H, W], same dtype; feeds the loss kernels when the caller hands the head's H/4 logits to the loss. */
int sh_upsample_bilinear(const void* in, int dtype, void* out, long planes, int h, int w, int H, int W, void* stream);

/* Exact adjoint of the above (deterministic gather): full-resolution gradient [planes, H, W] -> [planes, h, w]. */
int sh_upsample_bilinear_adjoint(const void* gout, int dtype, void* gin, long planes, int h, int w, int H, int W,
                                 void* stream);

/* Aux-head loss of train.py:309-313: nn.CrossEntropyLoss(ignore_index=255)(F.interpolate(aux_logits, (H, W)), label)
 * computed from the LOW-resolution aux logits [B, C, h, w] (any H, W), forward value and (grad != NULL) gradient w.r.t.
 * the low-resolution logits in one fused kernel; no [B, C, H, W] tensor exists.  out_loss[0] = mean over valid pixels.
 * grad_out: device scalar multiplied into the gradient (NULL = 1).  workspace: sh_aux_ce_workspace_bytes. */
size_t sh_aux_ce_workspace_bytes(int B, int C, int h, int w);
int sh_aux_ce_fwdbwd(const void* logits, int dtype, const void* label, int label_dtype, int B, int C, int h, int w, int H,
                     int W, void* grad, const float* grad_out, float* out_loss, void* workspace, void* stream);

/* Validation / inference decode of train.py:350-385 and infer.py:296-312 without the full-resolution logits: per-level
 * argmax of the bilinearly upsampled logits + fine pixel-accuracy counts straight from [B, C, h, w], for H = 4h, W = 4w
 * (the head's geometry); other sizes return SH_ERR_UNSUPPORTED (-2) and the host runs sh_upsample_bilinear + sh_decode. */
int sh_decode_upsampled(const void* logits, int dtype, int B, int C, int h, int w, int H, int W, int n0, int n1, int n2,
                        void* out0, void* out1, void* out2, int out_is_u8, const void* label, int label_dtype,
                        unsigned long long* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGHIERO_B200_H */
