/*
 * gsr_b200.h — C ABI of the B200-native differentiable Gaussian rasterizer.
 *
 * Drop-in boundary for W-Ted/GScream's `submodules/diff-gaussian-rasterization`
 * (paths below are relative to that directory; CR/ = cuda_rasterizer/).  The
 * reference binds its CUDA through five pybind11 functions (ext.cpp:16-20) over the
 * C++ statics CudaRasterizer::Rasterizer::{forward,backward,visible_filter,
 * position2D_filter,markVisible} (CR/rasterizer.h:20-133).  This header is what a
 * replacement of those statics binds instead: plain device pointers, ints, floats and a
 * cudaStream_t; no torch types, no exceptions, no allocation, no state between calls.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *  - all floating point is fp32, indices int32/uint32;
 *  - viewmatrix / projmatrix are the 16 floats of the reference's (already transposed,
 *    i.e. column-major) world_view_transform / full_proj_transform (scene/cameras.py:64-67);
 *  - every function returns 0 on success, a cudaError_t value (> 0) if a launch failed, or a
 *    negative GSR_E_* code for argument errors.  Errors are sticky-free: nothing is cached.
 *  - `stream` is an opaque cudaStream_t (pass NULL for the legacy default stream);
 *  - the three scratch buffers (geom / binning / image) are caller-allocated device byte
 *    buffers sized by the gsr_*_bytes() queries and handed back unchanged to gsr_backward,
 *    exactly like the reference's geomBuffer / binningBuffer / imgBuffer
 *    (rasterize_points.cu:77-82, diff_gaussian_rasterization/__init__.py:105,122).
 *    Their layout is private to this library.
 */
#ifndef GSR_B200_H_INCLUDED
#define GSR_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_ABI_VERSION 2

#if defined(__GNUC__)
#define GSR_API __attribute__((visibility("default")))
#else
#define GSR_API
#endif

#define GSR_E_BADARG      (-1) /* null pointer / negative size / misaligned buffer            */
#define GSR_E_CHANNELS    (-2) /* unsupported channel count (see gsr_supported_channels)       */
#define GSR_E_WORKSPACE   (-3) /* a scratch buffer is smaller than gsr_*_bytes() requires       */
#define GSR_E_SH_CHANNELS (-4) /* SH input with C != 3 (CR/rasterizer_impl.cu:246-249)         */

typedef void *gsr_stream_t;

/* ABI / capability queries ------------------------------------------------------------- */
GSR_API int gsr_abi_version(void);
/* returns 1 if `channels` colour/feature channels can be blended (compiled instantiations) */
GSR_API int gsr_supported_channels(int channels);
/* human-readable text for a return code of this library (CUDA codes are forwarded to
 * cudaGetErrorString) */
GSR_API const char *gsr_error_string(int code);

/* scratch sizing — replaces required<GeometryState/ImageState/BinningState>()
 * (CR/rasterizer_impl.h:64-73, CR/rasterizer_impl.cu:155-195) */
GSR_API size_t gsr_geom_bytes(int P);
GSR_API size_t gsr_image_bytes(int width, int height);
GSR_API size_t gsr_binning_bytes(int P, int64_t num_rendered, int width, int height);
/* The binning buffer describes itself: its layout follows from its size (the largest instance count whose layout fits), so a
 * buffer may be sized from an ESTIMATE of num_rendered before the exact value has reached the host.  gsr_binning_capacity
 * returns that count for a buffer of `binning_bytes` bytes (0: too small for anything). */
GSR_API int64_t gsr_binning_capacity(int P, int width, int height, size_t binning_bytes);

/*
 * Forward, first half — replaces the part of Rasterizer::forward before the blocking
 * cudaMemcpy of num_rendered (CR/rasterizer_impl.cu:199-287): per-Gaussian cull + EWA
 * projection (preprocessCUDA, CR/forward.cu:157-267), depth ordering and the prefix sum of
 * tiles_touched.
 *   means3D[P,3] opacities[P] uncertainties[P] scales[P,3] rotations[P,4] (or cov3D_precomp[P,6])
 *   shs[P,M,3] with sh_degree (or NULL/0 when colours are precomputed — the GScream case)
 *   colors_precomp[P,C] (or NULL when shs is given; then C must be 3)
 *   radii[P] (out, int32) — the reference's `radii` return value
 *   num_rendered_host: PINNED host int64; written asynchronously on `stream`.  The caller
 *   synchronises the stream (or an event) before reading it — this is the one host sync the
 *   reference API forces (it returns num_rendered as a Python int).
 */
GSR_API int gsr_forward_stage1(
    int P, int C, int sh_degree, int M,
    const float *means3D, const float *shs, const float *colors_precomp,
    const float *opacities, const float *uncertainties,
    const float *scales, float scale_modifier, const float *rotations, const float *cov3D_precomp,
    const float *viewmatrix, const float *projmatrix, const float *campos,
    int width, int height, float tan_fovx, float tan_fovy, int prefiltered,
    int *radii, void *geom_buffer, size_t geom_bytes,
    int64_t *num_rendered_host, gsr_stream_t stream);

/*
 * Forward, second half — replaces CR/rasterizer_impl.cu:289-346: (tile, depth) instance
 * emission (duplicateWithKeys :70-111), stable sort (:309-314), identifyTileRanges (:116-138)
 * and the blend kernel renderCUDA (CR/forward.cu:441-568).
 *   background[C]; out_color[C,H,W]; out_depth[1,H,W]; out_uncertainty[1,H,W]
 * No host round trip is needed between the two halves: the kernels read num_rendered from device memory (stage 1 left it in
 * geom_buffer) and take their grids from the capacity of `binning_buffer` (gsr_binning_capacity).
 *   num_rendered >= 0: the exact value, if the caller already has it (checked against the capacity: GSR_E_WORKSPACE);
 *   num_rendered == -1: not known yet.  The caller launches this call right behind stage 1 with a buffer sized from an
 *                  estimate, then waits for stage 1's pinned counter only; if the counter exceeds gsr_binning_capacity() the
 *                  call has rendered nothing but the background (every tile range empty, no out-of-bounds access) and is
 *                  simply repeated with a large enough buffer.
 */
GSR_API int gsr_forward_stage2(
    int P, int C, int64_t num_rendered,
    const float *colors_precomp, const float *background,
    int width, int height,
    void *geom_buffer, size_t geom_bytes,
    void *binning_buffer, size_t binning_bytes,
    void *image_buffer, size_t image_bytes,
    float *out_color, float *out_depth, float *out_uncertainty, gsr_stream_t stream);

/*
 * Backward — replaces Rasterizer::backward (CR/rasterizer_impl.cu:536-643): blend backward
 * (renderCUDA, CR/backward.cu:409-604), computeCov2DCUDA (:144-274) and preprocessCUDA
 * backward (:346-406).
 * Gradient outputs (all fp32, shapes as rasterize_points.cu:160-170):
 *   dL_dmeans2D[P,3] dL_dcolors[P,C] dL_dopacity[P] dL_duncertainty[P]  (accumulated with atomics)
 *   dL_dmeans3D[P,3] dL_dcov3D[P,6] dL_dscales[P,3] dL_drotations[P,4] dL_dsh[P,M,3] (written)
 * accumulate == 0: every output is overwritten (zero for culled Gaussians) — the reference's
 *                  semantics with its torch::zeros allocation folded in;
 * accumulate != 0: outputs are added to (gradient accumulation over views into one bucket;
 *                  the caller zeroes the bucket once per step).
 * dL_dcov3D / dL_dsh / dL_dscales / dL_drotations may be NULL when not wanted.
 */
GSR_API int gsr_backward(
    int P, int C, int sh_degree, int M, int64_t num_rendered,
    const float *background, int width, int height,
    const float *means3D, const float *shs, const float *colors_precomp,
    const float *scales, float scale_modifier, const float *rotations, const float *cov3D_precomp,
    const float *viewmatrix, const float *projmatrix, const float *campos,
    float tan_fovx, float tan_fovy, const int *radii,
    void *geom_buffer, size_t geom_bytes,
    void *binning_buffer, size_t binning_bytes,
    void *image_buffer, size_t image_bytes,
    const float *dL_dout_color, const float *dL_dout_depth, const float *dL_dout_uncertainty,
    float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity, float *dL_duncertainty,
    float *dL_dmeans3D, float *dL_dcov3D, float *dL_dsh, float *dL_dscales, float *dL_drotations,
    int accumulate, gsr_stream_t stream);

/* Anchor pre-filters — replace Rasterizer::visible_filter / position2D_filter /
 * markVisible (CR/rasterizer_impl.cu:350-406, 470-530, 141-153).  No scratch needed.
 * scales_stride: floats between consecutive rows of `scales` (3 = contiguous [P,3]).  GScream calls the filters with
 * `get_scaling[:, :3]`, a row-strided view of a [A,6] tensor (gaussian_renderer/__init__.py:298): it is read in place. */
GSR_API int gsr_visible_filter(
    int P, const float *means3D, const float *scales, int scales_stride, float scale_modifier, const float *rotations,
    const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix,
    int width, int height, float tan_fovx, float tan_fovy, int prefiltered,
    int *radii, gsr_stream_t stream);

GSR_API int gsr_position2d_filter(
    int P, const float *means3D, const float *scales, int scales_stride, float scale_modifier, const float *rotations,
    const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix,
    int width, int height, float tan_fovx, float tan_fovy, int prefiltered,
    int *radii, float *position2D_x, float *position2D_y, gsr_stream_t stream);

GSR_API int gsr_mark_visible(
    int P, const float *means3D, const float *viewmatrix, const float *projmatrix,
    uint8_t *present, gsr_stream_t stream);

/*
 * Anchor -> neural-Gaussian decode — replaces generate_neural_gaussians
 * (gaussian_renderer/__init__.py:18-102 of W-Ted/GScream, use_feat_bank = False): visible-anchor gather, the view
 * features ob_view / ob_dist, the four 36->32->{k, k, 7k, 3k} MLPs (scene/gaussian_model.py:118-144), the
 * neural_opacity > 0 selection and the post-processing into the rasterizer's inputs.  Like the rasterizer it is split
 * at the one host sync the reference's interface forces (the output sizes are data dependent).
 *   anchor[A,3] anchor_feat[A,32] offset[A,k,3] scaling[A,6] (= exp(_scaling), scene/gaussian_model.py:241-242)
 *   visible_mask[A] bytes (0 / non-0) or NULL for "all anchors"; campos[3]
 *   mlp_params: HOST array of 16 device pointers, torch.nn.Linear layout, in the order
 *               {opacity, uncertainty, cov, colour} x {w1[32,36], b1[32], w2[n_out,32], b2[n_out]}, n_out = k, k, 7k, 3k
 *   scratch: gsr_decode_scratch_bytes(A) bytes; written by stage 1, read by stage 2 and by the backward.
 * feat_dim must be 32 and 1 <= n_offsets <= 16 (gsr_decode_supported), else GSR_E_BADARG; anchor_feat and g_feat must be
 * 16-byte aligned (rows are moved as float4), else GSR_E_BADARG.
 *
 * Stage 1: neural_opacity[A*k] (tanh of the opacity MLP, rows of visible anchors in order, first n_vis*k entries valid),
 *          mask[A*k] (neural_opacity > 0), counts_host (PINNED int64[2], pre-zeroed by the callee): [0] n_vis, [1] P.
 */
GSR_API int gsr_decode_supported(int feat_dim, int n_offsets);
GSR_API size_t gsr_decode_scratch_bytes(int A);
GSR_API int gsr_decode_stage1(
    int A, int feat_dim, int n_offsets,
    const float *anchor, const float *anchor_feat, const uint8_t *visible_mask, const float *campos,
    const float *const *mlp_params, void *scratch, size_t scratch_bytes,
    float *neural_opacity, uint8_t *mask, int64_t *counts_host, gsr_stream_t stream);
/* Stage 2: xyz[P,3] color[P,3] opacity[P] uncertainty[P] out_scaling[P,3] rot[P,4] of the kept offsets, in
 * (visible anchor, offset) order — gaussian_renderer/__init__.py:66-101.
 * n_vis == -1: stage 1 ran with a visibility mask and its counts have not reached the host yet; the kernel reads n_vis from the
 * scratch buffer and P is the CAPACITY (rows) of the six outputs, at least A * n_offsets.  The caller launches this right behind
 * stage 1, waits for the counts only, and narrows the outputs to their first P rows. */
GSR_API int gsr_decode_stage2(
    int A, int feat_dim, int n_offsets, int64_t n_vis, int64_t P,
    const float *anchor, const float *anchor_feat, const float *offset, const float *scaling, const float *campos,
    const float *const *mlp_params, void *scratch, size_t scratch_bytes, const float *neural_opacity,
    float *xyz, float *color, float *opacity, float *uncertainty, float *out_scaling, float *rot, gsr_stream_t stream);
/* Backward of both stages.  Upstream gradients (any may be NULL = zero): d_xyz[P,3] d_color[P,3] d_opacity[P]
 * d_uncertainty[P] d_scaling[P,3] d_rot[P,4] and d_neural_opacity[n_vis*k].  Outputs g_anchor[A,3] g_feat[A,32]
 * g_offset[A,k,3] g_scaling[A,6] must arrive zero-filled (rows of invisible anchors are not touched);
 * g_mlp_params (HOST array of 16 device pointers, same order as mlp_params) is accumulated into. */
GSR_API int gsr_decode_backward(
    int A, int feat_dim, int n_offsets, int64_t n_vis, int64_t P,
    const float *anchor, const float *anchor_feat, const float *offset, const float *scaling, const float *campos,
    const float *const *mlp_params, void *scratch, size_t scratch_bytes,
    const float *d_xyz, const float *d_color, const float *d_opacity, const float *d_uncertainty,
    const float *d_scaling, const float *d_rot, const float *d_neural_opacity,
    float *g_anchor, float *g_feat, float *g_offset, float *g_scaling, float *const *g_mlp_params, gsr_stream_t stream);

/*
 * Densification statistics (first half) — replaces GaussianModel.training_statis
 * (scene/gaussian_model.py:729-757): ~15 boolean-mask index kernels (each a host sync) become two scans and one kernel.
 *   anchor_visible_mask[A], offset_selection_mask[n_vis*k] (= neural_opacity > 0), update_filter[P] (= radii > 0): bytes
 *   neural_opacity[n_vis*k], viewspace_grad[P,3] (the .grad of render()'s screenspace_points = dL_dmeans2D)
 *   in/out (fp32, updated in place): opacity_accum[A], anchor_demon[A], offset_gradient_accum[A*k], offset_denom[A*k]
 *   scratch: gsr_training_statis_scratch_bytes(A, k).
 */
GSR_API size_t gsr_training_statis_scratch_bytes(int A, int n_offsets);
GSR_API int gsr_training_statis(
    int A, int n_offsets, int64_t n_vis, int64_t P,
    const uint8_t *anchor_visible_mask, const uint8_t *offset_selection_mask, const uint8_t *update_filter,
    const float *neural_opacity, const float *viewspace_grad,
    float *opacity_accum, float *anchor_demon, float *offset_gradient_accum, float *offset_denom,
    void *scratch, size_t scratch_bytes, gsr_stream_t stream);

/*
 * Fused photometric loss, L1 + SSIM — replaces l1_loss / l1_loss_masked
 * (utils/loss_utils.py:27-31) and ssim / ssim_masked (utils/loss_utils.py:131-207, 11x11 Gaussian window, sigma 1.5,
 * zero padding, C1 = 0.01^2, C2 = 0.03^2, mean over all elements) as used by train.py:535-545.
 *   image, target: [planes, H, W] (planes = batch x channels; the window is depthwise); mask: [mask_planes, H, W] with
 *   mask_planes == 1 (broadcast) or == planes, or NULL; taps11_host: HOST pointer to the 11 fp32 taps of the normalised 1-D
 *   window (the reference's 2-D window is their outer product, loss_utils.py:112-121).
 * forward : sums (DEVICE fp64[2]) = { sum(ssim_map * mask), sum(|image - target| * mask) } — divide by planes*H*W for the
 *           reference's means; partials (DEVICE fp32 [3, planes, H, W], or NULL if no backward follows) is scratch for
 *           the backward.
 * backward: upstream (DEVICE fp32[2]) = { dL/d(mean ssim), dL/d(mean l1) }; grad_image [planes, H, W] is overwritten.
 */
GSR_API int gsr_l1_ssim_forward(
    int planes, int height, int width, const float *taps11_host,
    const float *image, const float *target, const float *mask, int mask_planes,
    double *sums, float *partials, gsr_stream_t stream);
GSR_API int gsr_l1_ssim_backward(
    int planes, int height, int width, const float *taps11_host,
    const float *image, const float *target, const float *mask, int mask_planes,
    const float *partials, const float *upstream, float *grad_image, gsr_stream_t stream);

/*
 * Scale / shift aligned depth L1 — replaces compute_scale_and_shift
 * (utils/loss_utils.py:80-102) + torch.abs(scale) + l1_loss / l1_loss_masked on the aligned depth as train.py:548-569 combines
 * them, including the gradient through the closed-form fit.
 *   depth, target, fit_mask (NULL = ones), loss_mask (NULL = l1_loss, else l1_loss_masked): [batch, H, W]
 *   state: DEVICE fp64[1 + 7 * batch]: [0] = sum(|abs(s) depth + t - target| * loss_mask) over the whole batch (divide by
 *          batch*H*W for the reference's mean), then the five fit sums and two backward sums per image.
 * backward: upstream = DEVICE fp32 scalar dL/d(mean); grad_depth [batch, H, W] is overwritten.
 */
GSR_API int gsr_depth_align_l1_forward(
    int batch, int height, int width, const float *depth, const float *target, const float *fit_mask, const float *loss_mask,
    double *state, gsr_stream_t stream);
GSR_API int gsr_depth_align_l1_backward(
    int batch, int height, int width, const float *depth, const float *target, const float *fit_mask, const float *loss_mask,
    const double *state, const float *upstream, float *grad_depth, gsr_stream_t stream);

/*
 * Multi-scale gradient-matching depth loss — replaces the loop of train.py:556-560 / :571-574:
 *     for scale in range(4): loss += w * gradient_loss(aligned[:, ::2^scale, ::2^scale], target[...], mask[...])
 * with gradient_loss / reduction_image_based of train.py:221-251 (sum of |horizontal| + |vertical| first differences of
 * mask * (prediction - target), pair-masked, divided by the mask sum of the sub-sampled grid, mean over the batch).
 *   prediction, target, mask (NULL = ones): [batch, H, W]; n_scales in 1..4 (strides 1, 2, 4, 8).
 *   fit_state: NULL -> `prediction` is used as it is (plain gradient_loss, gradient w.r.t. prediction);
 *              else the `state` written by gsr_depth_align_l1_forward for the same depth / target / fit_mask: `prediction` is
 *              the raw rendered depth, aligned inside as abs(s) * depth + t, and the backward carries the gradient through the fit.
 *   gstate: DEVICE fp64[1 + 16 * batch]: [0] = the summed loss, then {sum, M, d sum / d abs(s), d sum / d t} per (image, scale).
 * backward: upstream = DEVICE fp32 scalar dL/d(summed loss); grad [batch, H, W] is overwritten (accumulate = 0) or added to.
 */
GSR_API int gsr_depth_grad_forward(
    int batch, int height, int width, int n_scales, const float *prediction, const float *target, const float *mask,
    const double *fit_state, double *gstate, gsr_stream_t stream);
GSR_API int gsr_depth_grad_backward(
    int batch, int height, int width, int n_scales, const float *prediction, const float *target, const float *mask,
    const float *fit_mask, const double *fit_state, const double *gstate, const float *upstream, float *grad, int accumulate,
    gsr_stream_t stream);

/*
 * Multi-tensor Adam step (second half) — replaces `gaussians.optimizer.step()` (train.py:611) of the
 * optimizer built by GaussianModel.training_setup as torch.optim.Adam(l, lr=0.0, eps=1e-15) (scene/gaussian_model.py:374-407):
 * every tensor of every parameter group in ONE launch (24 tensors per launch), each element read and written once.
 * Arithmetic and order of operations are torch.optim.Adam's (amsgrad = False, maximize = False):
 *     g += weight_decay * p;  m += (g - m) (1 - beta1);  v = v beta2 + (1 - beta2) g g;
 *     p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 * with the two bias corrections evaluated in double precision on the host from each tensor's own `step`.
 *   tensors_host: HOST array of n_tensors descriptors of DEVICE fp32 arrays (contiguous, numel elements each; numel == 0 is
 *   skipped); `step` is the step count INCLUDING this update (>= 1).  param / exp_avg / exp_avg_sq are updated in place.
 */
typedef struct gsr_adam_tensor {
    float *param;
    const float *grad;
    float *exp_avg;
    float *exp_avg_sq;
    int64_t numel;
    int64_t step;
    double lr, beta1, beta2, eps, weight_decay;   /* doubles: torch evaluates 1 - beta and lr / (1 - beta^step) in Python floats */
} gsr_adam_tensor;
GSR_API int gsr_adam_step(int n_tensors, const gsr_adam_tensor *tensors_host, gsr_stream_t stream);

/*
 * Introspection for parity tests (device -> device copies out of the private scratch layout).
 * Any output pointer may be NULL.  Shapes: xy[P,2] depths[P] conic_opacity[P,4]
 * tiles_touched[P] point_list[R] ranges[tiles,2] final_T[H*W] n_contrib[H*W].
 * Entries of culled Gaussians (radii == 0) in xy/depths/conic_opacity are unspecified.
 */
GSR_API int gsr_debug_export(
    int P, int64_t num_rendered, int width, int height,
    const void *geom_buffer, const void *binning_buffer, size_t binning_bytes, const void *image_buffer,
    float *xy, float *depths, float *conic_opacity, uint32_t *tiles_touched,
    uint32_t *point_list, uint32_t *ranges, float *final_T, uint32_t *n_contrib,
    gsr_stream_t stream);

/* Test hook: the sorted instance list normally carries an 8-bit per-warp overlap mask above a 24-bit Gaussian id; for more than
 * 2^24 Gaussians it falls back to plain ids (every warp then treats every instance as a candidate: same results, slower).
 * gsr_debug_plain_point_list(1) forces that fallback at any size so that it can be tested; returns the previous setting.
 * Must not be toggled between a forward and its backward. */
GSR_API int gsr_debug_plain_point_list(int on);

/*
 * Optional per-stage device timing for benchmarks.  gsr_profile_enable(1) (re)starts recording: every stage
 * launch is bracketed by a cudaEvent pair on the caller's stream (up to 256 per stage, nothing is
 * synchronised); gsr_profile_enable(0) stops.  After the caller has synchronised, gsr_profile_read() copies the
 * elapsed milliseconds of the recorded launches of one stage to HOST memory and returns how many there were.
 * Stages: 0 preprocess, 1 depth order + scan, 2 instance binning, 3 blend forward, 4 blend backward,
 * 5 per-Gaussian backward, 6 decode forward (stages 1 + 2), 7 decode backward,
 * 8 image / depth losses forward, 9 image / depth losses backward, 10 Adam step.
 */
GSR_API int gsr_profile_enable(int on);
GSR_API int gsr_profile_read(int stage, float *ms_host, int capacity);

/* Number of this library's kernel launches since the last call with reset != 0
 * (bench.py reports it as `gpu_launches`). */
GSR_API int64_t gsr_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* GSR_B200_H_INCLUDED */
