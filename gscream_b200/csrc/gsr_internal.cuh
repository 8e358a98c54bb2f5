// gsr_internal.cuh — declarations shared between the C-ABI translation unit (gsr_api.cu) and the kernel files.
#pragma once
#include "gsr_common.cuh"

namespace gsr {

// arguments of the per-Gaussian forward kernels (gsr_preprocess.cu)
struct PreArgs {
	int P, C, D, M;
	const float *means3D, *scales, *rotations, *opacities, *uncertainties, *cov3D_precomp, *shs, *colors_precomp;
	const float *view, *proj, *campos;
	float scale_modifier;
	int scales_stride;   // floats between consecutive rows of `scales` (3 = contiguous)
	int W, H;
	float tan_fovx, tan_fovy, focal_x, focal_y;
	int gx, gy;
	int prefiltered;
	int *radii;
	float *rec;
	uint32_t *tiles_touched, *depth_key, *depth_val;
	float *pos_x, *pos_y;
	uint8_t *clamped;
	float *rgb;
};

// arguments of the fused per-Gaussian backward kernel (gsr_preprocess.cu)
struct PreBwdArgs {
	int P, C, D, M;
	const float *means3D, *scales, *rotations, *cov3D_precomp, *shs;
	const float *view, *proj, *campos;
	const uint8_t *clamped;
	float scale_modifier;
	int W, H;
	float tan_fovx, tan_fovy, h_x, h_y;
	const int *radii;
	const float *gacc;
	float *dL_dmeans2D, *dL_dopacity, *dL_duncertainty, *dL_dcolors;
	float *dL_dmeans3D, *dL_dcov3D, *dL_dsh, *dL_dscales, *dL_drotations;
	int accumulate;
};

// launchers (each returns the launch status; all asynchronous on `stream`)
cudaError_t launch_preprocess(int mode, const PreArgs &a, cudaStream_t stream);
cudaError_t launch_mark_visible(int P, const float *means3D, const float *view, uint8_t *present, cudaStream_t stream);
cudaError_t launch_preprocess_backward(const PreBwdArgs &a, cudaStream_t stream);
cudaError_t depth_order_and_scan(int P, char *geom, const GeomLayout &L, cudaStream_t stream);
cudaError_t bin_instances(int P, int64_t capacity, int W, int H, char *geom, const GeomLayout &GL, char *binning, const BinningLayout &BL,
                          char *image, const ImageLayout &IL, cudaStream_t stream);
int depth_order_index();
int point_list_index(int W, int H);
// (`header`: the binning buffer's self-description — the blend kernels read the point_list format from it)
cudaError_t launch_blend_forward(int C, int W, int H, const uint2 *ranges, const uint32_t *header, uint32_t *point_list, const float *rec,
                                 const float *features, const float *bg, float *final_T, uint32_t *n_contrib, float *out_color,
                                 float *out_depth, float *out_unc, cudaStream_t stream);
cudaError_t launch_blend_backward(int C, int W, int H, const uint2 *ranges, const uint32_t *header, const uint32_t *point_list, const float *rec,
                                  const float *features, const float *bg, const float *final_Ts, const uint32_t *n_contrib,
                                  const float *dL_dpixels, const float *dL_dpixel_depths, const float *dL_dpixel_uncs, float *gacc,
                                  float *dL_dcolors, cudaStream_t stream);

} // namespace gsr
