// gsr_loss.cuh — declarations of the fused L1 + SSIM image loss (gsr_loss.cu), shared with gsr_api.cu.
#pragma once
#include "gsr_common.cuh"

namespace gsr {

// x = rendered planes, y = target planes, both [planes][H][W]; mask [mask_planes][H][W] with mask_planes in {1, planes}, or null.
// sums (device, fp64[2]) receives sum(ssim_map * mask) and sum(|x - y| * mask); p1..p3 (each [planes][H][W], or all null when no
// backward will follow) the mask-weighted partial derivatives the backward convolves.  taps11: HOST pointer to the 11 window taps.
cudaError_t launch_l1_ssim_forward(int planes, int H, int W, const float *taps11, const float *x, const float *y, const float *mask,
                                   int mask_planes, double *sums, float *p1, float *p2, float *p3, cudaStream_t stream);
// upstream (device, fp32[2]): dL/d(mean ssim), dL/d(mean l1).  dx [planes][H][W] is overwritten.
cudaError_t launch_l1_ssim_backward(int planes, int H, int W, const float *taps11, const float *x, const float *y, const float *mask,
                                    int mask_planes, const float *p1, const float *p2, const float *p3, const float *upstream, float *dx,
                                    cudaStream_t stream);

// Scale / shift aligned depth L1.  d, y (and the optional masks) are [B][n]; sums fp64[B][5], aux fp64[B][2], loss_sum fp64[1]
// (all device, written by the forward, read by the backward); upstream: device fp32 scalar dL/dloss; dd [B][n] is overwritten.
cudaError_t launch_depth_align_l1_forward(int B, int n, const float *d, const float *y, const float *fit_mask, const float *loss_mask,
                                          double *sums, double *aux, double *loss_sum, cudaStream_t stream);
cudaError_t launch_depth_align_l1_backward(int B, int n, const float *d, const float *y, const float *fit_mask, const float *loss_mask,
                                           const double *sums, const double *aux, const float *upstream, float *dd, cudaStream_t stream);

// Multi-scale gradient-matching loss (train.py:232-251, :556-560, :571-574).  pred / target / mask (NULL = ones) / fit_mask: [B][H][W];
// fit_sums: fp64[B][5] as written by launch_depth_align_l1_forward (pred is then the raw depth, aligned inside), or NULL (pred is
// used as is).  gstate: fp64[1 + 16 B]: [0] = sum over scales of the reference's gradient_loss, then {sum, M, S1, S0} per (image, scale).
cudaError_t launch_depth_grad_forward(int B, int H, int W, int n_scales, const float *d, const float *y, const float *mask,
                                      const double *fit_sums, double *gstate, cudaStream_t stream);
cudaError_t launch_depth_grad_backward(int B, int H, int W, int n_scales, const float *d, const float *y, const float *mask,
                                       const float *fit_mask, const double *fit_sums, const double *gstate, const float *upstream,
                                       float *dd, int accumulate, cudaStream_t stream);

} // namespace gsr
