/*
 * cosa_b200 — C-ABI of the B200 (sm_100a) CAM -> pseudo-label refinement path of CoSA.
 *
 * Plain pointers and sizes only; no torch types.  Unless a name ends in `_host`, every pointer is a
 * DEVICE pointer, every call is asynchronous on `stream` (a cudaStream_t passed as void*) and the
 * caller owns all buffers, including the scratch `ws` whose size the matching `*_ws_bytes` returns.
 * All tensors are contiguous, float32 NCHW unless stated.  Return value: 0 = ok, < 0 = COSA_E_*,
 * > 0 = a cudaError_t raised by a launch.
 *
 * Each entry point cites the reference interface (file:line under the CoSA tree) it replaces; the
 * reference-side binding a maintainer would add is shown in INTEGRATION.md.
 */
#ifndef COSA_B200_H_
#define COSA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COSA_B200_ABI_VERSION 2

enum {
  COSA_OK = 0,
  COSA_E_ARG = -1,        /* bad shape / null pointer / unsupported parameter            */
  COSA_E_WORKSPACE = -2,  /* ws_bytes smaller than *_ws_bytes() for the same arguments  */
  COSA_E_KEYRANGE = -3    /* lattice coordinate outside the packed-key range: |coordinate| <= 6143 here,
                             the reference stores `short`; the filter output and the energy loss are NaN when this
                             happens, and cosa_bilateral_stats / the _host form return this code */
};

int cosa_abi_version(void);
/* Human-readable text for a return code (static storage). */
const char *cosa_strerror(int code);
/* Number of kernel launches issued through this library since load (bench.py's gpu_launches). */
unsigned long long cosa_launch_count(void);
/* Instrumentation for bench.py's roofline: between begin and end every kernel launch of the library is
 * bracketed by CUDA events on its own stream; end synchronises the device and writes one text line per
 * kernel, "name count total_ms\n", into buf.  Returns the number of distinct kernels. */
void cosa_profile_begin(void);
int cosa_profile_end(char *buf, size_t buf_len);

/* ------------------------------------------------------------------------------------------------
 * PAR — pixel-adaptive refinement.            replaces models/PAR.py:64-91 (PAR.forward)
 *   imgs   [B,3,h,w]; masks_in [B,C,hm,wm] (bilinearly resized to h,w with align_corners=True when
 *   hm,wm differ, PAR.py:66); masks_out [B,C,h,w].
 *   dilations[n_dil] is a HOST array (ctor argument, PAR.py:28); the 8*n_dil neighbours follow
 *   get_kernel() order (PAR.py:10-24) with replicate borders (PAR.py:44).  w1 = 0.3, w2 = 0.01.
 * ---------------------------------------------------------------------------------------------- */
size_t cosa_par_ws_bytes(int B, int C, int h, int w, int n_dil);
int cosa_par_forward(const float *imgs, const float *masks_in, float *masks_out, int B, int C, int h, int w,
                     int hm, int wm, const int *dilations, int n_dil, int num_iter, void *ws, size_t ws_bytes,
                     void *stream);
/* Selects the propagation kernel used by cosa_par_forward / cosa_cam2mask from now on (process-wide, meant for A/B
 * runs and tests; the results agree to the rounding of the fp32 sums): "chain" (default: ALL num_iter steps in ONE
 * launch - one CTA per step x image tile, tile-level step counters in the workspace instead of grid barriers;
 * "chain<G>" fixes the number of images per group), "tile" (one launch per step, steps 2..T as programmatic dependent
 * launches), "smem" (the generic per-step kernel that non-reference dilation sets use).  Returns COSA_E_ARG for an
 * unknown name. */
int cosa_par_set_step_mode(const char *name);

/* The affinity alone, [B, 8*n_dil, h, w] (PAR.py:69-85); used by tests and by cam2mask internally. */
int cosa_par_affinity(const float *imgs, float *aff, int B, int h, int w, const int *dilations, int n_dil,
                      void *stream);

/* ------------------------------------------------------------------------------------------------
 * CAM normalise.                              replaces utils/seg_helper.py:264-270
 *   out = sum_s scale_maps[s]; out -= min over (H,W); out /= max over (H,W) + 1e-5, per (b,c) plane.
 *   scale_maps: HOST array of n_scales DEVICE pointers, each [planes, HW]; minmax_ws: 2*planes floats.
 * ---------------------------------------------------------------------------------------------- */
int cosa_cam_normalize(const float *const *scale_maps, int n_scales, float *out, int planes, long long HW,
                       float *minmax_ws, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-scale merge of the teacher's raw outputs.   replaces utils/seg_helper.py:253-270 (CAMs) and :260-262,273 (seg)
 *   raw: HOST array of n_scales DEVICE pointers, raw[s] = [2B, C, hs[s], ws[s]] (the image batch followed by its
 *   horizontally flipped copy, as multi_scale_camseg feeds the model, :250-251); hs/ws HOST arrays.
 *   cam merge : out[B,C1,H,W] = normalise( sum_s relu(max(up(raw_s[:B]), flip(up(raw_s[B:])))) ), `up` = bilinear,
 *               align_corners=False; normalise = per-plane (x - min) / (max - min + 1e-5).  minmax_ws: 2*B*C1 floats.
 *               (The auxiliary CAM of the reference keeps only the LAST scale, :258 - call with n_scales = 1.)
 *   seg merge : out[B,C,H,W] = sum_s ( up(raw_s[:B]) + flip(up(raw_s[B:])) ).
 * ---------------------------------------------------------------------------------------------- */
int cosa_multi_scale_cam_merge(const float *const *raw, const int *hs, const int *ws, int n_scales, float *out, int B,
                               int C1, int H, int W, float *minmax_ws, void *stream);
/* The same followed by cam_validation (seg_helper.py:547-551, main.py:137: the very next call on the merged CAMs):
 * out = cls_label[b,c] * merged.  Planes whose label is 0 are zero-filled without being merged - 18 of 20 at VOC. */
int cosa_multi_scale_cam_merge_valid(const float *const *raw, const int *hs, const int *ws, int n_scales,
                                     const float *cls_label, float *out, int B, int C1, int H, int W,
                                     float *minmax_ws, void *stream);
/* The same for a consumer that reads only the planes of present classes (cosa_cam2mask_ex with
 * COSA_CAM2MASK_CAMS_UNVALIDATED): planes whose label is 0 are NOT written at all and the planes of present classes
 * hold the merged, normalised CAM without the label factor.  cosa_cam_validation(out, cls_label, out) turns the result
 * into what cosa_multi_scale_cam_merge_valid writes.  Requires W % 4 == 0, n_scales <= 5 and a 16-byte aligned `out`
 * (COSA_E_ARG otherwise: use the _valid form). */
int cosa_multi_scale_cam_merge_present(const float *const *raw, const int *hs, const int *ws, int n_scales,
                                       const float *cls_label, float *out, int B, int C1, int H, int W,
                                       float *minmax_ws, void *stream);
int cosa_multi_scale_seg_merge(const float *const *raw, const int *hs, const int *ws, int n_scales, float *out, int B,
                               int C, int H, int W, void *stream);

/* denormalize_img (utils/torch_helper.py:354-367; main.py:117 derives cam2mask's [0,1] image from the network
 * input on the device): out = (uint8)(imgs * std[c] + mean[c]) / 255 for imgs [B,3,H,W], HW = H*W.  Product and sum
 * are rounded separately (torch evaluates two ops), the cast truncates; inputs whose de-normalised value leaves
 * [0,256) are unspecified in the reference (float -> uint8 overflow) and clamp here. */
int cosa_denormalize_img(const float *imgs, float *out, int B, long long HW, const float mean[3], const float std[3],
                         void *stream);

/* cam_validation: out[b,c,:,:] = cls_label[b,c] * cam[b,c,:,:].   utils/seg_helper.py:547-551 */
int cosa_cam_validation(const float *cam, const float *cls_label, float *out, int B, int C1, long long HW,
                        void *stream);

/* ------------------------------------------------------------------------------------------------
 * cam_to_label.                               replaces utils/seg_helper.py:515-545
 *   cam [B,C1,H,W]; cls_label [B,C1] or NULL; label_out int64 [B,H,W].
 *   boxes: int32 [B,4] = (y0,y1,x0,x1) already resolved to 0 <= y0 <= y1 <= H (Python slice rules), or
 *   NULL for the `img_box is None` branch (label only, no ignore_mid, no box fill).
 *   valid_cam_out may be NULL.
 * ---------------------------------------------------------------------------------------------- */
int cosa_cam_to_label(const float *cam, const float *cls_label, const int *boxes, float *valid_cam_out,
                      long long *label_out, int B, int C1, int H, int W, float bkg_thre, float high_thre,
                      float low_thre, int ignore_mid, long long ignore_index, void *stream);

/* ------------------------------------------------------------------------------------------------
 * cam2mask — swapped-assignment pseudo-labels.     replaces utils/seg_helper.py:721-785 + :787-797
 *   images [B,3,H,W] in [0,1]; cams [B,C1,H,W]; cls_labels [B,C1]; boxes int32 [B,4] resolved as above;
 *   label_out float32 [B,H,W] with values {0..C1, ignore_index}.
 *   downscale: 0 = none, else the half-resolution is (H/downscale, W/downscale) (seg_helper.py:738-739).
 *   use_par = 0 reproduces `refine_model=None` (the shipped default); use_par = 1 runs PAR with
 *   dilations[n_dil] (HOST) and num_iter on both threshold stacks of every image in one batch.
 *   label_high_out / label_low_out (float32 [B,H,W], may be NULL) expose the two per-threshold maps
 *   before the merge rule (seg_helper.py:781-783).
 * ---------------------------------------------------------------------------------------------- */
size_t cosa_cam2mask_ws_bytes(int B, int C1, int H, int W, int downscale, int use_par, int n_dil);
int cosa_cam2mask(const float *images, const int *boxes, const float *cams, const float *cls_labels,
                  float threshold_high, float threshold_low, float ignore_index, int downscale, int use_par,
                  const int *dilations, int n_dil, int num_iter, float *label_out, float *label_high_out,
                  float *label_low_out, int B, int C1, int H, int W, void *ws, size_t ws_bytes, void *stream);

/* cosa_cam2mask with option flags.  COSA_CAM2MASK_REUSE_AFFINITY: the caller guarantees that the previous call on
 * this workspace (same stream, nothing else written to it since) had the same images, geometry and dilations - e.g.
 * main.py:158 and :191, which label the CAMs and the auxiliary CAMs of one batch - so the PAR affinity still in the
 * workspace is used instead of being computed again. */
#define COSA_CAM2MASK_REUSE_AFFINITY 1
int cosa_cam2mask_flags(const float *images, const int *boxes, const float *cams, const float *cls_labels,
                        float threshold_high, float threshold_low, float ignore_index, int downscale, int use_par,
                        const int *dilations, int n_dil, int num_iter, float *label_out, float *label_high_out,
                        float *label_low_out, int B, int C1, int H, int W, void *ws, size_t ws_bytes, int flags,
                        void *stream);

/* cosa_cam2mask_flags plus two producers folded into its first kernel, so that neither main.py:117's [0,1] image
 * nor main.py:137's validated CAM tensor has to exist in HBM:
 *   denorm_mean / denorm_std (HOST float[3], both or neither): `images` is the ImageNet-normalised network input and
 *     every source pixel is de-normalised exactly as cosa_denormalize_img does (utils/torch_helper.py:354-367);
 *   COSA_CAM2MASK_CAMS_UNVALIDATED: `cams` has not been through cam_validation; every source value is multiplied by
 *     cls_labels[b,c] (utils/seg_helper.py:547-551).  Only planes of present classes are read either way.
 * COSA_CAM2MASK_ALL_CHANNELS: with PAR, cosa_cam2mask refines softmax stacks whose channels sum to 1, and a PAR step
 *   multiplies that sum by the constant row sum of its weights (PAR.py:85-89), so by default the last live channel of
 *   each stack is not propagated but evaluated as row_sum^num_iter - (sum of the others) by the labelling kernel (one
 *   third of the propagation work at two foreground classes; the value differs from a propagated one by ~1e-6, so a
 *   label can differ only where the reference's own top-1/top-2 margin is a numerical tie).  This flag propagates
 *   every channel like the reference does. */
#define COSA_CAM2MASK_ALL_CHANNELS 2
#define COSA_CAM2MASK_CAMS_UNVALIDATED 4
/* Scratch sizing: the mask buffers hold both threshold stacks of (1 + present classes) channels per image, and by
 * default are sized for every class being present (1.7 GB at VOC B = 32).  COSA_CAM2MASK_MAX_CLASSES(n) (bits 8..15 of
 * `flags`; give the same flags to cosa_cam2mask_ws_bytes_ex) sizes them for at most n present classes per image (VOC
 * images have at most 6: 0.8 GB).  An image with more is detected on the device: every label of the call becomes NaN. */
#define COSA_CAM2MASK_MAX_CLASSES(n) (((n) & 0xff) << 8)
#define COSA_CAM2MASK_MAX_CLASSES_MASK 0xff00
size_t cosa_cam2mask_ws_bytes_ex(int B, int C1, int H, int W, int downscale, int use_par, int n_dil, int flags);
int cosa_cam2mask_ex(const float *images, const int *boxes, const float *cams, const float *cls_labels,
                     float threshold_high, float threshold_low, float ignore_index, int downscale, int use_par,
                     const int *dilations, int n_dil, int num_iter, float *label_out, float *label_high_out,
                     float *label_low_out, int B, int C1, int H, int W, void *ws, size_t ws_bytes, int flags,
                     const float *denorm_mean, const float *denorm_std, void *stream);

/* _refine_cams tail: bilinear (align_corners=False) resize of refined [B,nc,h,w] to (H,W), argmax over
 * channels (first max wins), label = valid_key[argmax].      utils/seg_helper.py:793-795 */
int cosa_upsample_argmax(const float *refined, const long long *valid_key, long long *label_out, int B, int nc,
                         int h, int w, int H, int W, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Bilateral filter on the permutohedral lattice.
 *   replaces utils/bilateralfilter/bilateralfilter.hpp:12 bilateralfilter_batch (SWIG: bilateralfilter.i)
 *   images [N,3,H,W] (planar RGB, 0..255), ins/outs [N,K,H,W]; outs is the UN-normalised response.
 *   `_host` takes HOST pointers, allocates its own device scratch and is synchronous: it is the
 *   9-argument call of utils/seg_helper.py:887.
 * ---------------------------------------------------------------------------------------------- */
size_t cosa_bilateral_ws_bytes(int N, int K, int H, int W);
int cosa_bilateralfilter_batch(const float *images, const float *ins, float *outs, int N, int K, int H, int W,
                               float sigmargb, float sigmaxy, void *ws, size_t ws_bytes, void *stream);
int cosa_bilateralfilter_batch_host(const float *images, const float *ins, float *outs, int N, int K, int H,
                                    int W, float sigmargb, float sigmaxy);
/* Lattice statistics of the last build in `ws` (same N,K,H,W as that call; for the dense-energy entry points
 * the lattice is the tail of their workspace), copied to host after synchronising `stream`:
 * stats[0] = M (vertices over the whole batch), stats[1] = error flags (1 = key range, 2 = capacity), stats[2] = table
 * capacity in use, stats[3] = max probe length.  Returns COSA_E_KEYRANGE / COSA_E_WORKSPACE when a flag is set. */
int cosa_bilateral_stats(const void *ws, int N, int K, int H, int W, long long stats[4], void *stream);

/* ------------------------------------------------------------------------------------------------
 * DenseEnergyLossFunction.                    replaces utils/seg_helper.py:864-903 (rrm_utils.py:352-391)
 *   forward : gate = clamp_min(ROI - max_k S, 0) with gate[unlabel] = 1; S <- S*ROI; AS = filter(S);
 *             AS <- AS*gate; loss = -<S, AS>/N.   Writes as_out [N,K,H,W] (gated, saved for backward)
 *             and loss_out[0].  unlabel: uint8 [N,H,W].
 *   backward: grad_seg = -2 * grad_out[0] * AS / N * ROI      (grad_out is a DEVICE scalar)
 * ---------------------------------------------------------------------------------------------- */
size_t cosa_dense_energy_ws_bytes(int N, int K, int H, int W);
int cosa_dense_energy_forward(const float *images, const float *segs, const float *rois,
                              const unsigned char *unlabel, float sigmargb, float sigmaxy, float *as_out,
                              float *loss_out, int N, int K, int H, int W, void *ws, size_t ws_bytes,
                              void *stream);
int cosa_dense_energy_backward(const float *as_saved, const float *rois, const float *grad_out, float *grad_seg,
                               int N, int K, int H, int W, void *stream);

/* ------------------------------------------------------------------------------------------------
 * get_energy_loss fused.      replaces utils/seg_helper.py:210-230 + DenseEnergyLoss.forward :199-208
 *   simg [B,3,H,W] ImageNet-normalised; logit [B,C,H,W]; label float32 [B,H,W] ({0..C-1,255});
 *   boxes int32 [B,4] resolved; mean/std: HOST float[3].  Half resolution is (H/2, W/2) (scale 0.5,
 *   H and W even): images/ROI/label nearest, softmax(logit) bilinear.  loss_out[0] = weight * energy.
 *   forward keeps what backward needs in `saved` (cosa_energy_loss_saved_bytes): gated AS and ROI.
 *   backward writes grad_logit [B,C,H,W] = d(loss_out)/d(logit) * grad_out[0].
 * ---------------------------------------------------------------------------------------------- */
size_t cosa_energy_loss_ws_bytes(int B, int C, int H, int W);
size_t cosa_energy_loss_saved_bytes(int B, int C, int H, int W);
int cosa_energy_loss_forward(const float *simg, const float *logit, const float *label, const int *boxes,
                             const float *mean, const float *std, float weight, float sigmargb, float sigmaxy_scaled,
                             float *loss_out, void *saved, int B, int C, int H, int W, void *ws, size_t ws_bytes,
                             void *stream);
/* The lattice of get_energy_loss depends only on the image (utils/bilateralfilter/bilateralfilter.cpp:4-19 builds it
 * from the image alone), not on the labels or logits.  cosa_energy_loss_prebuild runs that image-only half - the
 * de-normalised nearest 2:1 image and the lattice build - into `ws` on `stream`, which may be a second stream so that
 * the build overlaps cam2mask.  A later cosa_energy_loss_forward_flags(..., COSA_ENERGY_LATTICE_PREBUILT, ...) with
 * the same simg / mean / std / sigmas / B,C,H,W / ws (ordered after the prebuild by the caller: event or same stream)
 * skips the build; the kernels and their inputs are those of cosa_energy_loss_forward.  Batches that need more than one
 * lattice chunk (B > 64) return COSA_E_ARG from both. */
#define COSA_ENERGY_LATTICE_PREBUILT 1
/* Vertex budget (bits 8..15 of `flags`, in sixteenths of a vertex per half-resolution pixel; 0 = worst case).  The
 * lattice of an image can have up to 6 vertices per pixel, and the default workspace is sized for that (2.9 GB at
 * VOC B = 32, of which the two value buffers are 1.9 GB); natural images need 0.2 - 0.6 (the synthetic VOC batch: 0.50,
 * uniform noise: 2.4).  COSA_ENERGY_VERTEX_BUDGET(1.5) sizes the vertex arrays for 1.5 vertices per pixel (0.8 GB at VOC
 * B = 32).  The same flags must be given to ws_bytes_ex, prebuild_ex and forward_flags.  A batch that outgrows its
 * budget is detected, not overrun: the loss and its gradient become NaN and cosa_bilateral_stats on the lattice
 * (cosa_energy_loss_lattice_offset bytes into `ws`) returns COSA_E_WORKSPACE. */
#define COSA_ENERGY_VERTEX_BUDGET(vertices_per_pixel) (((int)((vertices_per_pixel) * 16.0f + 0.5f) & 0xff) << 8)
#define COSA_ENERGY_VERTEX_BUDGET_MASK 0xff00
size_t cosa_energy_loss_ws_bytes_ex(int B, int C, int H, int W, int flags);
size_t cosa_energy_loss_lattice_offset(int B, int C, int H, int W);
int cosa_energy_loss_prebuild_ex(const float *simg, const float *mean, const float *std, float sigmargb,
                                 float sigmaxy_scaled, int B, int C, int H, int W, void *ws, size_t ws_bytes, int flags,
                                 void *stream);
int cosa_energy_loss_prebuild(const float *simg, const float *mean, const float *std, float sigmargb,
                              float sigmaxy_scaled, int B, int C, int H, int W, void *ws, size_t ws_bytes,
                              void *stream);
int cosa_energy_loss_forward_flags(const float *simg, const float *logit, const float *label, const int *boxes,
                                   const float *mean, const float *std, float weight, float sigmargb,
                                   float sigmaxy_scaled, float *loss_out, void *saved, int B, int C, int H, int W,
                                   void *ws, size_t ws_bytes, int flags, void *stream);
/* The same, with the hand-over of a lattice prebuilt on ANOTHER stream made inside the call: `lattice_ready`
 * (a cudaEvent_t recorded behind cosa_energy_loss_prebuild on that stream; needs COSA_ENERGY_LATTICE_PREBUILT) is
 * waited for on `stream` AFTER the softmax / gate kernel has been enqueued, so that HBM-bound kernel runs beside the
 * latency-bound build instead of behind it.  NULL: plain cosa_energy_loss_forward_flags. */
int cosa_energy_loss_forward_ev(const float *simg, const float *logit, const float *label, const int *boxes,
                                const float *mean, const float *std, float weight, float sigmargb, float sigmaxy_scaled,
                                float *loss_out, void *saved, int B, int C, int H, int W, void *ws, size_t ws_bytes,
                                int flags, void *lattice_ready, void *stream);
int cosa_energy_loss_backward(const float *logit, const void *saved, const float *grad_out, float weight,
                              float *grad_logit, int B, int C, int H, int W, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Dense-CRF mean-field inference (evaluation-time post-processing).
 *   replaces utils/seg_helper.py:961-996 (class DenseCRF.__call__ / crf_inference_infv2, evaluation_engine.py:205-211)
 *   and :905-922 (crf_inference_inf), i.e. pydensecrf's DenseCRF2D with setUnaryEnergy(unary_from_softmax(probs)),
 *   addPairwiseGaussian(sxy=pos_xy_std, compat=pos_w), addPairwiseBilateral(sxy=bi_xy_std, srgb=bi_rgb_std,
 *   compat=bi_w) and inference(iter_max), kernel normalisation NORMALIZE_SYMMETRIC (pydensecrf's default).
 *   images [N,3,H,W] planar RGB 0..255 (the reference passes HWC uint8: the host mirror converts), probs [N,C,H,W]
 *   (softmax output), q_out [N,C,H,W] the mean-field marginals after iter_max iterations.  N <= 64.
 *   pydensecrf is not vendored by the reference (README.md:104 installs its git master): parity of this function is
 *   pinned on the reference's own permutohedral C++ for the two filters and on the published update rule - DESIGN.md.
 * ---------------------------------------------------------------------------------------------- */
size_t cosa_crf_inference_ws_bytes(int N, int C, int H, int W);
int cosa_crf_inference(const float *images, const float *probs, float *q_out, int N, int C, int H, int W, int iter_max,
                       float pos_w, float pos_xy_std, float bi_w, float bi_xy_std, float bi_rgb_std, void *ws,
                       size_t ws_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * The consumers either side of the path.
 *
 * seg_loss: fg/bg-balanced cross-entropy of seg_pred [B,C,H,W] against the pseudo-label map mask_label [B,H,W]
 *   (float, values 0..C-1 or ignore_index).                                  replaces utils/seg_helper.py:800-813
 *   loss = (1-fg_alpha) * CE_sum(label == 0)/(n_bg + 1e-6) + fg_alpha * CE_sum(label != 0, != ignore)/(n_fg + 1e-6).
 *   `stats` (cosa_seg_loss_stats_bytes() bytes, device) carries the sums and counts to the backward, which writes
 *   grad_pred = grad_out * d loss / d seg_pred.  fg_alpha outside [0,1] -> COSA_E_ARG (the reference asserts).
 * seg_refine_by_label: out = softmax over C of (seg with the channels of absent classes set to -1e5) / softmaxtemp
 *   (after_softmax = 0), or cls * softmax(seg / softmaxtemp) (after_softmax = 1); channel 0 is always present;
 *   cls_label [B,C-1] float 0/1.                                             replaces utils/seg_helper.py:553-568
 * cam_loss: F.multilabel_soft_margin_loss(relu(cam) [B,C,H,W], bilinear(seg_ps[:,1:] [B,C+1,Hs,Ws] -> H,W)).
 *   target_ws: B*C*H*W floats (kept for the backward), acc_ws: one double.    replaces utils/seg_helper.py:593-602
 * ---------------------------------------------------------------------------------------------- */
size_t cosa_seg_loss_stats_bytes(void);
int cosa_seg_loss_forward(const float *seg_pred, const float *mask_label, float fg_alpha, int ignore_index,
                          float *loss_out, void *stats, int B, int C, int H, int W, void *stream);
int cosa_seg_loss_backward(const float *seg_pred, const float *mask_label, const void *stats, const float *grad_out,
                           float fg_alpha, int ignore_index, float *grad_pred, int B, int C, int H, int W,
                           void *stream);
int cosa_seg_refine_by_label(const float *seg, const float *cls_label, float softmaxtemp, int after_softmax,
                             float *out, int B, int C, int H, int W, void *stream);
int cosa_cam_loss_forward(const float *cam, const float *seg_ps, int is_relu, float *loss_out, float *target_ws,
                          double *acc_ws, int B, int C, int H, int W, int Hs, int Ws, void *stream);
int cosa_cam_loss_backward(const float *cam, const float *target_ws, const float *grad_out, int is_relu,
                           float *grad_cam, int B, int C, int H, int W, void *stream);


/* Bilinear enlargement of the segmentation logits to the label size, main.py:167
 *   seg_pred = F.interpolate(seg_pred, size=(H, W), mode='bilinear', align_corners=False)
 * in [planes, h, w] -> out [planes, H, W] (planes = B*C), bit-exact against torch's CPU kernel for integer ratios
 * (28 -> 448), within an ulp otherwise; and its adjoint
 * (what autograd runs for the step): grad_out [planes, H, W] -> grad_in [planes, h, w], as a two-pass gather with a
 * caller-owned scratch of cosa_upsample_bilinear_backward_ws_bytes(planes, h, W) bytes. */
int cosa_upsample_bilinear(const float *in, float *out, long long planes, int h, int w, int H, int W, void *stream);
size_t cosa_upsample_bilinear_backward_ws_bytes(long long planes, int h, int W);
int cosa_upsample_bilinear_backward(const float *grad_out, float *grad_in, long long planes, int h, int w, int H,
                                    int W, void *ws, size_t ws_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* COSA_B200_H_ */
