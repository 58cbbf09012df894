/*
 * ocflow_b200 -- C ABI of the B200 (sm_100a) hot path of OCFlow.
 *
 * The reference (dongliangcao/OCFlow) is pure Python/PyTorch and has no FFI layer; its "plugin
 * boundary" for this path is a set of Python symbols.  This header declares the
 * C entry points that the Python mirror of those symbols (package ocflow_b200) binds with ctypes --
 * i.e. exactly what a maintainer of the reference would bind to replace each symbol.  Every entry
 * cites the reference interface it replaces (file:line relative to the reference root).
 *
 * Conventions (all entry points):
 *   - tensors are fp32, NCHW, contiguous unless a stride argument says otherwise; pointers are DEVICE
 *     pointers except in the ocf_host_* entry points, which take HOST pointers;
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is allocated, freed or
 *     cached inside the library and there is no global mutable state => re-entrant;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and the call returns without
 *     synchronising => CUDA-graph capturable; scatter targets are zeroed inside the call;
 *   - return value: 0 = OCF_OK, negative = OCF_E* argument error (nothing was launched),
 *     positive = cudaError_t reported by the launch;
 *   - there is NO CPU fallback and no dispatch on device type.
 */
#ifndef OCFLOW_B200_H
#define OCFLOW_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ocf_stream_t; /* cudaStream_t */

#define OCF_OK 0
#define OCF_ENULL (-1)        /* required pointer is NULL */
#define OCF_ESHAPE (-2)       /* non-positive or inconsistent dimension */
#define OCF_EUNSUPPORTED (-3) /* argument value outside what the kernels implement */
#define OCF_EALIGN (-4)       /* pointer not aligned to 4 bytes / stride not usable */

#define OCF_ABI_VERSION 3
#define OCF_MAX_DISPLACEMENT 16

int ocf_abi_version(void);
/* 100 for sm_100a builds */
int ocf_build_sm(void);
/* static string for an OCF_E* code or a cudaError_t */
const char* ocf_error_string(int code);

/* ---------------------------------------------------------------------------------------------
 * Cost volume.   Replaces compute_cost_volume(features1, features2, max_displacement=4)
 *                models/networks/correlation_layer.py:7-40 and the missing CostVolumeLayer
 *                (call sites cost_volume_flow_occ_net.py:53,188; flow_occ_net_c.py:26,99).
 *   out[b, (dy+d)*(2d+1)+(dx+d), y, x] = act( (1/C) * sum_c f1[b,c,y,x] * f2[b,c,y+dy,x+dx] )
 *   zero outside the image.  act = LeakyReLU(leaky_slope) when leaky_slope != 1 (the callers'
 *   next op, cost_volume_flow_net.py:173); pass 1.0f for the bare cost volume.
 *   out_bstride: elements between consecutive batch items of `out` (>= (2d+1)^2*H*W) so the result
 *   can be written straight into a wider channel-concatenated buffer; 0 means dense.
 *   norm: optional DEVICE pointer to {mean, inv_std}; when non-NULL both inputs are normalised on the
 *   fly as (x-mean)*inv_std inside the image (normalize_features with default flags,
 *   correlation_layer.py:42-82, fused with the correlation).
 *   mask_out: optional (d == 4, norm == NULL only) sign bitmask of the pre-activation cost volume for the backward of
 *   the fused LeakyReLU: [B][(2d+1)^2][H][ceil(W/8)] bytes, bit p of a byte = pixel x = 8*byte + p, set when the value
 *   is > 0.  Saving it instead of the activated output costs 1/32 of the bytes (B*81*H*ceil(W/8)).
 * ------------------------------------------------------------------------------------------- */
int ocf_corr_fwd(const float* f1, const float* f2, float* out, int B, int C, int H, int W, int d,
                 long long out_bstride, float leaky_slope, const float* norm, unsigned char* mask_out,
                 ocf_stream_t stream);

/* Backward of the above (autograd of correlation_layer.py:33-39).
 *   df1[b,c,y,x] = (1/C) sum_k g'[b,k,y,x] * f2[b,c,y+dy,x+dx]
 *   df2[b,c,y,x] = (1/C) sum_k g'[b,k,y-dy,x-dx] * f1[b,c,y-dy,x-dx]
 *   g' = grad_out, or grad_out * LeakyReLU'(out_act) when out_act != NULL (out_act = the activated
 *   forward output).  df1 or df2 may be NULL (not needed).
 *   g_bstride / act_bstride: batch strides of grad_out / out_act in elements, 0 = dense.  The gradient of a
 *   cost volume that was concatenated into a wider tensor (cost_volume_flow_net.py:176-180) arrives as a channel
 *   slice of the concat gradient: dense per batch item, wider batch stride -- it is consumed in place.
 *   mask: optional sign bitmask written by ocf_corr_fwd (d == 4); when non-NULL it replaces out_act (which may then be
 *   NULL) as the source of LeakyReLU'. */
int ocf_corr_bwd(const float* grad_out, const float* out_act, const float* f1, const float* f2,
                 float* df1, float* df2, int B, int C, int H, int W, int d, long long g_bstride,
                 long long act_bstride, float leaky_slope, const unsigned char* mask, ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused pyramid level (FlowNetCV, cost_volume_flow_net.py:186-190 and the 4 other levels):
 *   corr = LeakyReLU(compute_cost_volume(c1n, c2n, 4)), c1n = (f1 - mean) * inv_std, c2n likewise, with
 *   norm = {mean, inv_std} (device, from ocf_normalize_stats) applied ON LOAD; zero padding stays zero after the
 *   normalisation exactly as in normalize_features -> F.pad.  d = 4 only; runs on the tensor cores (tcgen05, 3xTF32).
 *   out / f1n_out may point into a wider concat buffer (batch strides in elements, 0 = dense); f2n_out (dense, may be
 *   NULL) receives the normalised second feature map for the backward; mask_out as in ocf_corr_fwd.
 *   norm == NULL: plain correlation (f1n_out / f2n_out must then be NULL).
 * ocf_level_corr_bwd: ocf_corr_bwd(d = 4, sign bitmask) with f1n read in place (batch stride f1n_bstride).
 * ------------------------------------------------------------------------------------------- */
int ocf_level_corr_fwd(const float* f1, const float* f2, const float* norm, float* out, long long out_bstride,
                       float* f1n_out, long long f1n_bstride, float* f2n_out, unsigned char* mask_out, int B, int C,
                       int H, int W, float leaky_slope, ocf_stream_t stream);
/* ocf_corr_fwd (d = 4, LeakyReLU + sign bitmask) with f1 read in place from a wider buffer (batch stride in elements) and the
 * cost volume written into one (out_bstride): the fp32 FMA form of the fused level, fed by ocf_normalize_apply, which
 * leaves the normalised first feature map where the decoder's concat (cost_volume_flow_net.py:190) wants it.  W % 4 == 0. */
int ocf_corr_fwd_strided(const float* f1, long long f1_bstride, const float* f2, float* out, long long out_bstride,
                         unsigned char* mask_out, int B, int C, int H, int W, float leaky_slope, ocf_stream_t stream);
int ocf_level_corr_bwd(const float* grad_out, long long g_bstride, const unsigned char* mask, const float* f1n,
                       long long f1n_bstride, const float* f2n, float* df1, float* df2, int B, int C, int H, int W,
                       float leaky_slope, ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Feature normalisation.  Replaces normalize_features(feature_list, normalize, center,
 *   moments_across_channels, moments_across_images)  models/networks/correlation_layer.py:42-82.
 *   All T tensors of the list must share one shape [B,C,H,W] (they do at every call site).
 *   flags: bit0 normalize, bit1 center, bit2 moments_across_channels, bit3 moments_across_images.
 *   stats workspace (device, caller-owned, 8-byte aligned): 8*T*B*G + 8 floats, G = 1 (across channels) or C
 *   -- fp64 partial sums, then per-group {mean, var}, then the {mean, inv_std} actually applied.
 * ------------------------------------------------------------------------------------------- */
#define OCF_NORM_NORMALIZE 1
#define OCF_NORM_CENTER 2
#define OCF_NORM_ACROSS_CHANNELS 4
#define OCF_NORM_ACROSS_IMAGES 8
int ocf_normalize_fwd(const float* const* xs, float* const* ys, int T, int B, int C, int H, int W,
                      int flags, float* stats, ocf_stream_t stream);
/* statistics pass of ocf_normalize_fwd alone (fills `stats`; the {mean, inv_std} applied to group g are
 * stats[6*NG + 2*g], stats[6*NG + 2*g + 1], NG = T*B*G).  With moments_across_images (the FlowNetCV call sites,
 * cost_volume_flow_net.py:171,187,...) every group carries the same scalar pair. */
int ocf_normalize_stats(const float* const* xs, int T, int B, int C, int H, int W, int flags, float* stats, ocf_stream_t stream);
/* apply pass alone: ys[t] = (xs[t] - mean) * inv_std with the statistics of a previous ocf_normalize_stats; ys[t] may point
 * into a wider buffer (y_bstrides[t] = elements between batch items, 0 or NULL array = dense). */
int ocf_normalize_apply(const float* const* xs, float* const* ys, const long long* y_bstrides, int T, int B, int C, int H, int W,
                        int flags, const float* stats, ocf_stream_t stream);
/* grads wrt every input, differentiating through the statistics (no detach in the reference).
 * red workspace (8-byte aligned): 8*T*B*G floats. */
int ocf_normalize_bwd(const float* const* grad_ys, const float* const* xs, float* const* grad_xs,
                      int T, int B, int C, int H, int W, int flags, const float* stats, float* red,
                      ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Bias + LeakyReLU epilogue of the FlowNetCV convolution blocks (cost_volume_flow_net.py:11-15: Conv2d(bias=True) followed by
 * LeakyReLU(0.1)); the convolution itself stays on cuDNN (out of scope), called without bias.
 *   fwd: y[b,c,:] = lrelu(x[b,c,:] + bias[c]) in one pass (y may alias x).
 *   bwd: grad_x = grad_y * (y > 0 ? 1 : slope), grad_bias[c] = sum over b and pixels of grad_x (zeroed inside), one pass.
 * ------------------------------------------------------------------------------------------- */
int ocf_bias_lrelu_fwd(const float* x, const float* bias, float* y, int B, int C, long long HW, float slope, ocf_stream_t stream);
int ocf_bias_lrelu_bwd(const float* grad_y, const float* y, float* grad_x, float* grad_bias, int B, int C, long long HW, float slope,
                       ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Bilinear resampling, align_corners=True.  Replaces the two F.interpolate calls either side of the hot path:
 *   F.interpolate(flow2, scale_factor=4, mode='bilinear', align_corners=True) * 20   (cost_volume_flow_net.py:245; mul = 20)
 *   F.interpolate(img1, scale_factor=0.25, mode='bilinear', align_corners=True)      (models/model.py:396)
 * in [planes, Hi, Wi] -> out [planes, Ho, Wo] (planes = B*C), out = mul * bilinear(in); ATen's upsample_bilinear2d
 * arithmetic.  The backward is a gather (no atomics): grad_in [planes, Hi, Wi] is fully written.
 * ------------------------------------------------------------------------------------------- */
int ocf_resize_bilinear_fwd(const float* in, float* out, int planes, int Hi, int Wi, int Ho, int Wo, float mul,
                            ocf_stream_t stream);
int ocf_resize_bilinear_bwd(const float* grad_out, float* grad_in, int planes, int Hi, int Wi, int Ho, int Wo, float mul,
                            ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Bilinear backward warp.  Replaces the 11 `warp`/`backwarp` bodies:
 *   align_corners=1: utils.py:20-58 (is_mask), models/model.py:191-221, 962-992, 1159-1189,
 *                    models/flow_model.py:49-79, models/networks/pwc_net.py:6-29
 *   align_corners=0: cost_volume_flow_net.py:121-151, cost_volume_flow_occ_net.py:137-167,419-449,
 *                    flow_net.py:57-87, flow_occ_net.py:96-126, inpainting_model.py:22-52
 *   out[b,c,y,x] = bilinear(img[b,c], ix, iy), zero padding, 4 taps bounds-tested independently;
 *   (ix,iy) reproduce the reference's normalise / ATen's un-normalise op order in fp32.
 *   flags: bit0 align_corners, bit1 is_mask (utils.py:49-57).
 *   scale: multiplies flow before use (the callers' `up_flow*0.625` etc., cost_volume_flow_net.py:186).
 *   occ: optional [B,1,H,W] multiplier applied to the output (the "woc" chain,
 *        cost_volume_flow_occ_net.py:204-205); NULL = none.
 * ------------------------------------------------------------------------------------------- */
#define OCF_WARP_ALIGN_CORNERS 1
#define OCF_WARP_IS_MASK 2
int ocf_warp_fwd(const float* img, const float* flow, const float* occ, float* out, int B, int C, int H,
                 int W, int flags, float scale, ocf_stream_t stream);
/* d_img (scatter, zeroed inside), d_flow (already multiplied by `scale`), d_occ: each may be NULL. */
int ocf_warp_bwd(const float* grad_out, const float* img, const float* flow, const float* occ,
                 float* d_img, float* d_flow, float* d_occ, int B, int C, int H, int W, int flags,
                 float scale, ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Range map / occlusion.  Replaces compute_range_map (models/model.py:243-305 =
 *   models/flow_model.py:101-163) with flow_to_warp (:223-241) folded in, and optionally
 *   occ = 1 - clamp(range, 0, 1) (models/model.py:391).
 *   range_out [B,1,H,W] is zeroed inside the call; occ_out may be NULL.
 * ------------------------------------------------------------------------------------------- */
int ocf_range_map(const float* flow, float* range_out, float* occ_out, int B, int H, int W,
                  ocf_stream_t stream);
/* flow_to_warp: [B,H,W,2] -> [B,H,W,2] endpoints (models/model.py:223-241) */
int ocf_flow_to_warp(const float* flow_bhw2, float* out_bhw2, int B, int H, int W, ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Charbonnier / photometric.  Replaces robust_l1 (models/model.py:27-35), charbonnier_loss
 *   (utils.py:8-18) and photometric_error(img_pred, img, occ) (models/model.py:37-46).
 * ------------------------------------------------------------------------------------------- */
int ocf_robust_l1_fwd(const float* x, float* y, long long n, float alpha, ocf_stream_t stream);
int ocf_robust_l1_bwd(const float* grad_y, const float* x, float* grad_x, long long n, float alpha,
                      ocf_stream_t stream);
/* sums[0] = sum rho(pred-img)*(1-occ) (or sum rho when occ==NULL), sums[1] = sum (1-occ) over B*H*W.
 * `sums` (2 doubles, device) is zeroed inside.  The scalar loss is assembled by the host mirror:
 * sums[0]/(3*sums[1]+1e-16) resp. sums[0]/n. */
int ocf_photometric_fwd(const float* pred, const float* img, const float* occ, double* sums, int B, int C,
                        int H, int W, float alpha, ocf_stream_t stream);
/* coef (device, 2 floats): d loss/d sums[0], d loss/d sums[1].  Any of d_pred/d_img/d_occ may be NULL. */
int ocf_photometric_bwd(const float* pred, const float* img, const float* occ, const float* coef,
                        float* d_pred, float* d_img, float* d_occ, int B, int C, int H, int W, float alpha,
                        ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Smoothness.  Replaces gradient (models/model.py:53-66), first_order_smoothness_loss (:93-101) and
 *   second_order_smoothness_loss (:103-114).  order = 1 or 2.  img [B,Ci,H,W], flow [B,Cf,H,W].
 *   sums (2 doubles, device, zeroed inside): sum_x, sum_y of w*rho(.) ; the host divides by the
 *   element counts and halves.
 * ------------------------------------------------------------------------------------------- */
int ocf_smooth_fwd(const float* img, const float* flow, double* sums, int B, int Ci, int Cf, int H, int W,
                   int order, float alpha_edge, float alpha_rho, ocf_stream_t stream);
/* coef (device, 2 floats): d loss / d sums[0..1].  d_img / d_flow are zeroed inside; may be NULL. */
int ocf_smooth_bwd(const float* img, const float* flow, const float* coef, float* d_img, float* d_flow,
                   int B, int Ci, int Cf, int H, int W, int order, float alpha_edge, float alpha_rho,
                   ocf_stream_t stream);
/* forward differences with a stride (models/model.py:53-66): dx [B,C,H,W-s], dy [B,C,H-s,W] */
int ocf_gradient(const float* img, float* dx, float* dy, int B, int C, int H, int W, int stride,
                 ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused occlusion-aware loss pass (one kernel per pyramid level of the loss):
 *   replaces the chain models/model.py:379,391,394,403,405,407 --
 *   warp(img2, flow) [align_corners=True] -> occ = 1-clamp(range,0,1) -> photometric(occ) ->
 *   photometric(1-occ) -> mse(flow, flow_gt) -> bce(occ_gt, occ), and emits the un-normalised
 *   d photo / d flow in the same pass.
 *   sums (8 doubles, device, zeroed inside):
 *     [0] sum rho*(1-occ)  [1] sum (1-occ)  [2] sum rho*occ  [3] sum occ
 *     [4] sum (flow-flow_gt)^2  [5] sum bce(occ_gt, occ)  [6..7] reserved
 *   dflow_unit [B,2,H,W] (may be NULL): sum_c rho'(diff_c)*(1-occ)*d warp_c/d(u,v).
 *   range_map / flow_gt / occ_gt may be NULL (terms skipped; occ = 0 when range_map is NULL).
 * ------------------------------------------------------------------------------------------- */
int ocf_occ_photo_fused(const float* img1, const float* img2, const float* flow, const float* range_map,
                        const float* flow_gt, const float* occ_gt, double* sums, float* dflow_unit,
                        float* warped_out, int B, int C, int H, int W, float alpha, ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Supervised losses (a-13): sums of |a-b| / (a-b)^2 / BCE / focal-BCE, one pass.
 *   kind: 0 = L1, 1 = MSE, 2 = BCE(p, target), 3 = focal(gamma=2) BCE
 *   (flow_model.py:173-186, flow_occ_model.py:48-55, occlusion_model.py:45-62).
 *   sum_out: 1 double (device, zeroed inside).  grad (may be NULL): d sum / d a, elementwise.
 * ------------------------------------------------------------------------------------------- */
int ocf_pair_loss(const float* a, const float* b, double* sum_out, float* grad, long long n, int kind,
                  ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * SSIM.  Replaces _ssim / ssim / SSIM.forward of inpainting_metrics/ssim/ssim.py:7-75 (the only "SSIM-style" term the
 *   reference defines): window x window Gaussian (sigma 1.5), zero padding window//2, C1 = 0.01^2, C2 = 0.03^2,
 *   depthwise over the C channels.  The map has Ho x Wo = (H + 2*(window/2) - window + 1) x (...) pixels per channel
 *   ((H+1) x (W+1) for even windows, as F.conv2d gives).  window <= 15.
 *   sums: B doubles, the sum of the map per batch item (mean = sums[b] / (C*Ho*Wo); the reference's size_average
 *   variants are both derived from it).
 *   coef: optional workspace of 4*B*C*Ho*Wo floats receiving the per-pixel partial derivatives for ocf_ssim_bwd.
 * ------------------------------------------------------------------------------------------- */
int ocf_ssim_fwd(const float* img1, const float* img2, double* sums, float* coef, int B, int C, int H, int W,
                 int window, ocf_stream_t stream);
/* d_img1 / d_img2 (either may be NULL) = scale[b] * d(sums[b]) / d img;  scale: B floats on the device. */
int ocf_ssim_bwd(const float* img1, const float* img2, const float* coef, const float* scale, float* d_img1,
                 float* d_img2, int B, int C, int H, int W, int window, ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Soft census (ternary) photometric term.  north_star lists a census term; the reference defines NONE (SURVEY.md section
 *   8a-14), so there is no reference interface to cite: PARITY UNPINNED, the semantics are those of
 *   oracle/ocflow_oracle.py::census_loss (UnFlow soft census: grey*255, (2m+1)^2 patch, t = u/sqrt(0.81+u^2), soft Hamming
 *   dt^2/(0.1+dt^2) averaged over the patch, m-pixel border masked) with the occlusion weighting of photometric_error
 *   (models/model.py:37-46).  max_distance m in 1..3.
 *   sums (2 doubles, device, zeroed inside): [0] sum dist*w, [1] sum w, w = interior * (1 - occ); occ may be NULL.
 *   The host forms sums[0] / (sums[1] + 1e-16).
 * ------------------------------------------------------------------------------------------- */
int ocf_census_fwd(const float* pred, const float* img, const float* occ, double* sums, int B, int C, int H, int W,
                   int max_distance, ocf_stream_t stream);
/* d_pred = coef[0] * d sums[0] / d pred  (coef: 1 float on the device; img is data, occ is computed under no_grad). */
int ocf_census_bwd(const float* pred, const float* img, const float* occ, const float* coef, float* d_pred, int B, int C,
                   int H, int W, int max_distance, ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * End-point-error metrics.  Replaces flow_error (models/data/utils/flow_utils.py:179-232) and
 *   flow_kitti_error (:234-271), which the reference evaluates in numpy on the host.  gt, pred: [B,2,H,W].
 *   kitti == 0: pixels whose ground truth exceeds 1e7 in magnitude are zeroed in all maps (:201-206) and still counted;
 *               sums[0] = sum of end-point errors, sums[1] = number of pixels.
 *   kitti == 1: only pixels with mask != 0 count (mask [B,1,H,W], NULL = all); sums[2] = number of outliers
 *               (epe > 3 and epe / (|gt| + 1e-5) > 0.05, :258-264).
 *   sums: 3 doubles on the device, zeroed inside.
 * ------------------------------------------------------------------------------------------- */
int ocf_flow_metrics(const float* gt, const float* pred, const float* mask, double* sums, int B, int H, int W, int kitti,
                     ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * On-device input pipeline.  Replaces, per sample, the host chain StaticCenterCrop
 *   (models/data/datasets.py:50-55, :164-166, :182) -> ToTensor -> Normalize(0.5, 0.5) (models/lightning_datamodule.py:20-23)
 *   -> cat(img1, img2) (datasets.py:179) and flow.transpose(2,0,1) (:185).
 *   img1, img2: [B,H0,W0,3] uint8 (HWC, as decoded); flow_hw2: optional [B,H0,W0,2] fp32; crop origin (y0, x0), size HxW.
 *   imgs: [B,6,H,W] fp32 = (v/255 - 0.5)/0.5 in torchvision's op order; flow: [B,2,H,W].
 * ------------------------------------------------------------------------------------------- */
int ocf_pack_pairs(const unsigned char* img1, const unsigned char* img2, const float* flow_hw2, float* imgs, float* flow,
                   int B, int H0, int W0, int H, int W, int y0, int x0, ocf_stream_t stream);
/* FlyingChairs2 ground-truth occlusion mask (models/data/datasets.py:660-669): occ_u8 [B,H0,W0] uint8 as decoded -> crop ->
 * occ [B,1,H,W] fp32 in {0, 1} (occ[occ > 0.5] = 1 ; occ[occ != 1] = 0 on the float copy of the decoded values). */
int ocf_pack_occ(const unsigned char* occ_u8, float* occ, int B, int H0, int W0, int H, int W, int y0, int x0,
                 ocf_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer convenience entry points (HOST pointers; allocate, copy in, run, copy out, free,
 * synchronise).  They exist so the C ABI can be exercised end-to-end without any Python/torch.
 * ------------------------------------------------------------------------------------------- */
int ocf_host_corr_fwd(const float* f1, const float* f2, float* out, int B, int C, int H, int W, int d);
int ocf_host_warp_fwd(const float* img, const float* flow, float* out, int B, int C, int H, int W, int flags);
int ocf_host_range_map(const float* flow, float* range_out, int B, int H, int W);

#ifdef __cplusplus
}
#endif
#endif /* OCFLOW_B200_H */
