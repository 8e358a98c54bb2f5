// gsr_api.cu — the extern "C" boundary of libgsr_b200.so (declared in include/gsr_b200.h).
//
// Orchestrates the stages that CudaRasterizer::Rasterizer::{forward,backward,visible_filter,
// position2D_filter,markVisible} orchestrate in the reference (CR/rasterizer_impl.cu:141-153,
// 199-347, 350-406, 470-530, 536-643), on the caller's stream, with caller-owned scratch.
#include "../../include/gsr_b200.h"
#include "gsr_internal.cuh"
#include "gsr_decode.cuh"
#include "gsr_loss.cuh"
#include "gsr_optim.cuh"
#include <atomic>
#include <cmath>

namespace gsr {

static std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_plain_point_list{0};
bool point_list_packed(int P) { return P <= (1 << 24) && !g_plain_point_list.load(std::memory_order_relaxed); }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional per-stage device timing (bench.py's roofline line) ---------------------------------
// When enabled, every stage launch is bracketed by a cudaEvent pair on the caller's stream (a ring of
// kProfCap pairs per stage).  Nothing is synchronised here; gsr_profile_read() is called after the
// caller's own synchronise.  Disabled (the default) it costs one branch per stage.
enum Stage { kPre = 0, kDepthScan, kBin, kBlendFwd, kBlendBwd, kPreBwd, kDecodeFwd, kDecodeBwd, kLossFwd, kLossBwd, kOptim, kNumStages };
static constexpr int kProfCap = 256;
static bool g_prof_on = false;
static cudaEvent_t g_ev[kNumStages][kProfCap][2];
static int g_ev_n[kNumStages];
static bool g_ev_made = false;
struct StageTimer {
	cudaStream_t s; int st; int slot;
	StageTimer(int stage, cudaStream_t stream) : s(stream), st(stage), slot(-1)
	{
		if (!g_prof_on || g_ev_n[st] >= kProfCap) return;
		slot = g_ev_n[st]++;
		cudaEventRecord(g_ev[st][slot][0], s);
	}
	~StageTimer() { if (slot >= 0) cudaEventRecord(g_ev[st][slot][1], s); }
};

__global__ void export_records_kernel(int P, const float *__restrict__ rec, float *xy, float *depths, float *conic_opacity)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P) return;
	const float *r = rec + (size_t)i * GSR_REC_FLOATS;
	if (xy) { xy[2 * i] = r[0]; xy[2 * i + 1] = r[1]; }
	if (depths) depths[i] = r[6];
	if (conic_opacity) {
		conic_opacity[4 * i + 0] = r[2];
		conic_opacity[4 * i + 1] = r[3];
		conic_opacity[4 * i + 2] = r[4];
		conic_opacity[4 * i + 3] = r[5];
	}
}

// point_list entries carry the per-warp overlap mask in their top byte (gsr_common.cuh: point_list_packed); the export
// returns plain Gaussian ids like the reference's point_list (CR/rasterizer_impl.h:55-62)
__global__ void export_point_list_kernel(int64_t R, const uint32_t *__restrict__ src, const uint32_t *__restrict__ header, uint32_t *__restrict__ dst)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t mask = header[kHdrPacked] ? 0x00FFFFFFu : 0xFFFFFFFFu;
	if (i < R) dst[i] = src[i] & mask;
}

static bool channels_ok(int C) { return C == 3 || C == 32; }
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

} // namespace gsr

using namespace gsr;

#define GSR_CUDA(x)                                  \
	do {                                             \
		cudaError_t _e = (x);                        \
		if (_e != cudaSuccess) return (int)_e;       \
	} while (0)

extern "C" {

int gsr_abi_version(void) { return GSR_ABI_VERSION; }
int gsr_supported_channels(int channels) { return channels_ok(channels) ? 1 : 0; }

const char *gsr_error_string(int code)
{
	switch (code) {
	case 0: return "success";
	case GSR_E_BADARG: return "gsr: bad argument (null pointer, negative size or misaligned buffer)";
	case GSR_E_CHANNELS: return "gsr: unsupported channel count (supported: 3, 32)";
	case GSR_E_WORKSPACE: return "gsr: scratch buffer too small";
	case GSR_E_SH_CHANNELS: return "For non-RGB, provide precomputed Gaussian colors!";
	default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "gsr: unknown error";
	}
}

size_t gsr_geom_bytes(int P) { return geom_layout(P).total; }
size_t gsr_image_bytes(int width, int height) { return image_layout(width, height).total; }
size_t gsr_binning_bytes(int P, int64_t num_rendered, int width, int height) { return binning_layout(P, num_rendered, width, height).total; }
int64_t gsr_binning_capacity(int P, int width, int height, size_t binning_bytes) { return binning_capacity(P, width, height, binning_bytes); }

int64_t gsr_launch_count(int reset)
{
	return reset ? g_launches.exchange(0) : g_launches.load();
}

int gsr_forward_stage1(int P, int C, int sh_degree, int M, const float *means3D, const float *shs, const float *colors_precomp,
                       const float *opacities, const float *uncertainties, const float *scales, float scale_modifier,
                       const float *rotations, const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix,
                       const float *campos, int width, int height, float tan_fovx, float tan_fovy, int prefiltered, int *radii,
                       void *geom_buffer, size_t geom_bytes, int64_t *num_rendered_host, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P < 0 || width <= 0 || height <= 0 || !num_rendered_host) return GSR_E_BADARG;
	*num_rendered_host = 0;
	if (P == 0) return 0; // rasterize_points.cu:85
	if (!means3D || !opacities || !uncertainties || !viewmatrix || !projmatrix || !radii || !geom_buffer) return GSR_E_BADARG;
	if (!cov3D_precomp && (!scales || !rotations)) return GSR_E_BADARG;
	if (!colors_precomp && !shs) return GSR_E_BADARG;
	if (!colors_precomp && C != 3) return GSR_E_SH_CHANNELS; // CR/rasterizer_impl.cu:246-249
	if (!colors_precomp && !campos) return GSR_E_BADARG;
	if (!channels_ok(C)) return GSR_E_CHANNELS;
	if (rotations && !aligned16(rotations)) return GSR_E_BADARG;
	const GeomLayout L = geom_layout(P);
	if (geom_bytes < L.total || !aligned16(geom_buffer)) return GSR_E_WORKSPACE;
	char *geom = (char *)geom_buffer;

	PreArgs a{};
	a.P = P; a.C = C; a.D = sh_degree; a.M = M;
	a.means3D = means3D; a.scales = scales; a.rotations = rotations; a.opacities = opacities; a.uncertainties = uncertainties;
	a.cov3D_precomp = cov3D_precomp; a.shs = shs; a.colors_precomp = colors_precomp;
	a.view = viewmatrix; a.proj = projmatrix; a.campos = campos;
	a.scale_modifier = scale_modifier;
	a.scales_stride = 3;
	a.W = width; a.H = height;
	a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
	a.focal_y = height / (2.0f * tan_fovy); // CR/rasterizer_impl.cu:226-227
	a.focal_x = width / (2.0f * tan_fovx);
	a.gx = (width + GSR_BLOCK_X - 1) / GSR_BLOCK_X;
	a.gy = (height + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y;
	a.prefiltered = prefiltered;
	a.radii = radii;
	a.rec = (float *)(geom + L.rec);
	a.tiles_touched = (uint32_t *)(geom + L.tiles_touched);
	a.depth_key = (uint32_t *)(geom + L.depth_key[0]);
	a.depth_val = (uint32_t *)(geom + L.depth_val[0]);
	a.clamped = (uint8_t *)(geom + L.clamped);
	a.rgb = (float *)(geom + L.rgb);
	{ StageTimer t(kPre, stream); GSR_CUDA(launch_preprocess(0, a, stream)); }
	{ StageTimer t(kDepthScan, stream); GSR_CUDA(depth_order_and_scan(P, geom, L, stream)); }
	// R = offsets[P-1]: 4 bytes into the low half of the (pre-zeroed, little-endian) int64
	GSR_CUDA(cudaMemcpyAsync(num_rendered_host, geom + L.offsets + (size_t)(P - 1) * 4, 4, cudaMemcpyDeviceToHost, stream));
	return 0;
}

int gsr_forward_stage2(int P, int C, int64_t num_rendered, const float *colors_precomp, const float *background, int width, int height,
                       void *geom_buffer, size_t geom_bytes, void *binning_buffer, size_t binning_bytes, void *image_buffer,
                       size_t image_bytes, float *out_color, float *out_depth, float *out_uncertainty, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P < 0 || width <= 0 || height <= 0 || num_rendered < -1) return GSR_E_BADARG;
	if (P == 0) return 0;
	if (!channels_ok(C)) return GSR_E_CHANNELS;
	if (!background || !geom_buffer || !binning_buffer || !image_buffer || !out_color || !out_depth || !out_uncertainty) return GSR_E_BADARG;
	if (C > 3 && (!colors_precomp || !aligned16(colors_precomp))) return GSR_E_BADARG;
	const GeomLayout GL = geom_layout(P);
	const ImageLayout IL = image_layout(width, height);
	// the binning buffer's layout follows from its size; num_rendered itself is read by the kernels from device memory
	const int64_t capacity = binning_capacity(P, width, height, binning_bytes);
	if (capacity < 1 || num_rendered > capacity) return GSR_E_WORKSPACE;
	const BinningLayout BL = binning_layout(P, capacity, width, height);
	if (geom_bytes < GL.total || image_bytes < IL.total) return GSR_E_WORKSPACE;
	if (!aligned16(geom_buffer) || !aligned16(binning_buffer) || !aligned16(image_buffer)) return GSR_E_WORKSPACE;
	char *geom = (char *)geom_buffer, *binning = (char *)binning_buffer, *image = (char *)image_buffer;

	{ StageTimer t(kBin, stream); GSR_CUDA(bin_instances(P, capacity, width, height, geom, GL, binning, BL, image, IL, stream)); }
	{
		StageTimer t(kBlendFwd, stream);
		GSR_CUDA(launch_blend_forward(C, width, height, (const uint2 *)(image + IL.ranges), (const uint32_t *)(binning + BL.header),
		                              (uint32_t *)(binning + BL.val[point_list_index(width, height)]),
		                              (const float *)(geom + GL.rec), colors_precomp, background, (float *)(image + IL.final_T),
		                              (uint32_t *)(image + IL.n_contrib), out_color, out_depth, out_uncertainty, stream));
	}
	return 0;
}

int gsr_backward(int P, int C, int sh_degree, int M, int64_t num_rendered, const float *background, int width, int height,
                 const float *means3D, const float *shs, const float *colors_precomp, const float *scales, float scale_modifier,
                 const float *rotations, const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix,
                 const float *campos, float tan_fovx, float tan_fovy, const int *radii, void *geom_buffer, size_t geom_bytes,
                 void *binning_buffer, size_t binning_bytes, void *image_buffer, size_t image_bytes, const float *dL_dout_color,
                 const float *dL_dout_depth, const float *dL_dout_uncertainty, float *dL_dmeans2D, float *dL_dcolors,
                 float *dL_dopacity, float *dL_duncertainty, float *dL_dmeans3D, float *dL_dcov3D, float *dL_dsh,
                 float *dL_dscales, float *dL_drotations, int accumulate, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return GSR_E_BADARG;
	if (P == 0) return 0; // rasterize_points.cu:172
	if (!channels_ok(C)) return GSR_E_CHANNELS;
	if (!background || !means3D || !viewmatrix || !projmatrix || !radii || !geom_buffer || !binning_buffer || !image_buffer) return GSR_E_BADARG;
	if (!dL_dout_color || !dL_dout_depth || !dL_dout_uncertainty) return GSR_E_BADARG;
	if (!dL_dmeans2D || !dL_dcolors || !dL_dopacity || !dL_duncertainty || !dL_dmeans3D) return GSR_E_BADARG;
	if (!cov3D_precomp && (!scales || !rotations)) return GSR_E_BADARG;
	if (C > 3 && (!colors_precomp || !aligned16(colors_precomp))) return GSR_E_BADARG;
	if (rotations && !aligned16(rotations)) return GSR_E_BADARG;
	if (shs && (!campos || !dL_dsh)) return GSR_E_BADARG;
	const GeomLayout GL = geom_layout(P);
	const ImageLayout IL = image_layout(width, height);
	const int64_t capacity = binning_capacity(P, width, height, binning_bytes); // the layout the forward used for this buffer
	if (capacity < 1 || num_rendered > capacity) return GSR_E_WORKSPACE;
	const BinningLayout BL = binning_layout(P, capacity, width, height);
	if (geom_bytes < GL.total || image_bytes < IL.total) return GSR_E_WORKSPACE;
	char *geom = (char *)geom_buffer, *binning = (char *)binning_buffer, *image = (char *)image_buffer;

	float *gacc = (float *)(geom + GL.gacc);
	GSR_CUDA(cudaMemsetAsync(gacc, 0, (size_t)P * 32, stream));
	// colour gradients: accumulated straight into the caller's tensor by the blend backward.  With SH input
	// they are an intermediate that the per-Gaussian kernel consumes, so they always start from zero.
	if (!accumulate || shs) GSR_CUDA(cudaMemsetAsync(dL_dcolors, 0, (size_t)P * C * sizeof(float), stream));
	count_launch(2);
	if (num_rendered > 0) {
		StageTimer t(kBlendBwd, stream);
		GSR_CUDA(launch_blend_backward(C, width, height, (const uint2 *)(image + IL.ranges), (const uint32_t *)(binning + BL.header),
		                               (const uint32_t *)(binning + BL.val[point_list_index(width, height)]),
		                               (const float *)(geom + GL.rec), colors_precomp, background, (const float *)(image + IL.final_T),
		                               (const uint32_t *)(image + IL.n_contrib), dL_dout_color, dL_dout_depth, dL_dout_uncertainty, gacc,
		                               dL_dcolors, stream));
	}
	PreBwdArgs a{};
	a.P = P; a.C = C; a.D = sh_degree; a.M = M;
	a.means3D = means3D; a.scales = cov3D_precomp ? nullptr : scales; a.rotations = rotations; a.cov3D_precomp = cov3D_precomp; a.shs = shs;
	a.view = viewmatrix; a.proj = projmatrix; a.campos = campos;
	a.clamped = (const uint8_t *)(geom + GL.clamped);
	a.scale_modifier = scale_modifier;
	a.W = width; a.H = height;
	a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
	a.h_y = height / (2.0f * tan_fovy); // CR/rasterizer_impl.cu:580-581
	a.h_x = width / (2.0f * tan_fovx);
	a.radii = radii;
	a.gacc = gacc;
	a.dL_dmeans2D = dL_dmeans2D; a.dL_dopacity = dL_dopacity; a.dL_duncertainty = dL_duncertainty; a.dL_dcolors = dL_dcolors;
	a.dL_dmeans3D = dL_dmeans3D; a.dL_dcov3D = dL_dcov3D; a.dL_dsh = dL_dsh; a.dL_dscales = dL_dscales; a.dL_drotations = dL_drotations;
	a.accumulate = accumulate;
	{ StageTimer t(kPreBwd, stream); GSR_CUDA(launch_preprocess_backward(a, stream)); }
	return 0;
}

int gsr_debug_plain_point_list(int on)
{
	return g_plain_point_list.exchange(on ? 1 : 0);
}

int gsr_profile_enable(int on)
{
	if (on && !g_ev_made) {
		for (int s = 0; s < kNumStages; s++)
			for (int i = 0; i < kProfCap; i++)
				for (int k = 0; k < 2; k++) GSR_CUDA(cudaEventCreate(&g_ev[s][i][k]));
		g_ev_made = true;
	}
	for (int s = 0; s < kNumStages; s++) g_ev_n[s] = 0;
	g_prof_on = on != 0;
	return 0;
}

int gsr_profile_read(int stage, float *ms_host, int capacity)
{
	if (stage < 0 || stage >= kNumStages || !ms_host || !g_ev_made) return GSR_E_BADARG;
	int n = g_ev_n[stage] < capacity ? g_ev_n[stage] : capacity;
	for (int i = 0; i < n; i++)
		if (cudaEventElapsedTime(&ms_host[i], g_ev[stage][i][0], g_ev[stage][i][1]) != cudaSuccess) return GSR_E_BADARG;
	return n;
}

static int run_filter(int mode, int P, const float *means3D, const float *scales, int scales_stride, float scale_modifier, const float *rotations,
                      const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix, int width, int height,
                      float tan_fovx, float tan_fovy, int prefiltered, int *radii, float *px, float *py, cudaStream_t stream)
{
	if (P < 0 || width <= 0 || height <= 0) return GSR_E_BADARG;
	if (P == 0) return 0;
	if (!means3D || !viewmatrix || !projmatrix || !radii) return GSR_E_BADARG;
	if (!cov3D_precomp && (!scales || !rotations)) return GSR_E_BADARG;
	if (rotations && !aligned16(rotations)) return GSR_E_BADARG;
	if (mode == 2 && (!px || !py)) return GSR_E_BADARG;
	if (scales && scales_stride < 3) return GSR_E_BADARG;
	PreArgs a{};
	a.P = P; a.C = 3;
	a.scales_stride = scales_stride;
	a.means3D = means3D; a.scales = scales; a.rotations = rotations; a.cov3D_precomp = cov3D_precomp;
	a.view = viewmatrix; a.proj = projmatrix;
	a.scale_modifier = scale_modifier;
	a.W = width; a.H = height;
	a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
	a.focal_y = height / (2.0f * tan_fovy);
	a.focal_x = width / (2.0f * tan_fovx);
	a.gx = (width + GSR_BLOCK_X - 1) / GSR_BLOCK_X;
	a.gy = (height + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y;
	a.prefiltered = prefiltered;
	a.radii = radii; a.pos_x = px; a.pos_y = py;
	GSR_CUDA(launch_preprocess(mode, a, stream));
	return 0;
}

int gsr_visible_filter(int P, const float *means3D, const float *scales, int scales_stride, float scale_modifier, const float *rotations,
                       const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix, int width, int height,
                       float tan_fovx, float tan_fovy, int prefiltered, int *radii, gsr_stream_t stream)
{
	return run_filter(1, P, means3D, scales, scales_stride, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, width, height, tan_fovx,
	                  tan_fovy, prefiltered, radii, nullptr, nullptr, (cudaStream_t)stream);
}

int gsr_position2d_filter(int P, const float *means3D, const float *scales, int scales_stride, float scale_modifier, const float *rotations,
                          const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix, int width, int height,
                          float tan_fovx, float tan_fovy, int prefiltered, int *radii, float *position2D_x, float *position2D_y,
                          gsr_stream_t stream)
{
	return run_filter(2, P, means3D, scales, scales_stride, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, width, height, tan_fovx,
	                  tan_fovy, prefiltered, radii, position2D_x, position2D_y, (cudaStream_t)stream);
}

int gsr_mark_visible(int P, const float *means3D, const float *viewmatrix, const float *projmatrix, uint8_t *present, gsr_stream_t stream)
{
	(void)projmatrix; // the reference's in_frustum computes p_proj but only tests view-space z (CR/auxiliary.h:154)
	if (P < 0) return GSR_E_BADARG;
	if (P == 0) return 0;
	if (!means3D || !viewmatrix || !present) return GSR_E_BADARG;
	GSR_CUDA(launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream));
	return 0;
}

// ---- anchor -> neural-Gaussian decode ---------------------------------------------------------------------------------
int gsr_decode_supported(int feat_dim, int n_offsets) { return (feat_dim == 32 && n_offsets >= 1 && n_offsets <= kDecMaxK) ? 1 : 0; }
size_t gsr_decode_scratch_bytes(int A) { return decode_layout(A).total; }

static bool unpack_weights(const float *const *p, DecodeWeights &w)
{
	if (!p) return false;
	for (int m = 0; m < 4; m++) {
		w.w1[m] = p[4 * m + 0]; w.b1[m] = p[4 * m + 1]; w.w2[m] = p[4 * m + 2]; w.b2[m] = p[4 * m + 3];
		if (!w.w1[m] || !w.b1[m] || !w.w2[m] || !w.b2[m]) return false;
	}
	return true;
}

int gsr_decode_stage1(int A, int feat_dim, int n_offsets, const float *anchor, const float *anchor_feat, const uint8_t *visible_mask,
                      const float *campos, const float *const *mlp_params, void *scratch, size_t scratch_bytes, float *neural_opacity,
                      uint8_t *mask, int64_t *counts_host, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (A < 0 || !counts_host || !gsr_decode_supported(feat_dim, n_offsets)) return GSR_E_BADARG;
	counts_host[0] = visible_mask ? 0 : A;
	counts_host[1] = 0;
	if (A == 0) return 0;
	DecodeWeights wt;
	if (!anchor || !anchor_feat || !campos || !scratch || !neural_opacity || !mask || !unpack_weights(mlp_params, wt)) return GSR_E_BADARG;
	if (!aligned16(anchor_feat)) return GSR_E_BADARG;   // feature rows are read as float4
	const DecodeLayout L = decode_layout(A);
	if (scratch_bytes < L.total || !aligned16(scratch)) return GSR_E_WORKSPACE;
	StageTimer t(kDecodeFwd, stream);
	GSR_CUDA(decode_stage1(A, n_offsets, anchor, anchor_feat, visible_mask, campos, wt, (char *)scratch, L, neural_opacity, mask, counts_host, stream));
	return 0;
}

static int fill_decode_args(DecodeArgs &a, int A, int feat_dim, int n_offsets, int64_t n_vis, int64_t P, const float *anchor,
                            const float *anchor_feat, const float *offset, const float *scaling, const float *campos,
                            const float *const *mlp_params, void *scratch, size_t scratch_bytes)
{
	if (A <= 0 || n_vis < 0 || n_vis > A || P < 0 || P > n_vis * n_offsets || !gsr_decode_supported(feat_dim, n_offsets)) return GSR_E_BADARG;
	if (!anchor || !anchor_feat || !offset || !scaling || !campos || !scratch || !unpack_weights(mlp_params, a.wt)) return GSR_E_BADARG;
	if (!aligned16(anchor_feat)) return GSR_E_BADARG;   // feature rows are read as float4
	const DecodeLayout L = decode_layout(A);
	if (scratch_bytes < L.total || !aligned16(scratch)) return GSR_E_WORKSPACE;
	char *s = (char *)scratch;
	a.k = n_offsets; a.n_vis = (int)n_vis; a.n_vis_dev = nullptr;
	a.vis_ids = n_vis == A ? nullptr : (const uint32_t *)(s + L.vis_ids); // all visible: the list is the identity
	a.anchor = anchor; a.feat = anchor_feat; a.offset = offset; a.scaling = scaling; a.campos = campos;
	a.count = (uint32_t *)(s + L.count); a.maskbits = (uint32_t *)(s + L.maskbits); a.gauss_incl = (const uint32_t *)(s + L.gauss_incl);
	return 0;
}

int gsr_decode_stage2(int A, int feat_dim, int n_offsets, int64_t n_vis, int64_t P, const float *anchor, const float *anchor_feat,
                      const float *offset, const float *scaling, const float *campos, const float *const *mlp_params, void *scratch,
                      size_t scratch_bytes, const float *neural_opacity, float *xyz, float *color, float *opacity, float *uncertainty,
                      float *out_scaling, float *rot, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	// n_vis == -1: the counts of stage 1 have not reached the host yet.  The kernel then reads n_vis from the scratch buffer
	// (stage 1 ran with a visibility mask and left the visible list there) and P is the CAPACITY of the six outputs, which
	// A * n_offsets rows always satisfy; the caller narrows them once the counts are in.
	const bool counts_on_device = n_vis == -1;
	if (counts_on_device) {
		if (A <= 0 || P < (int64_t)A * n_offsets) return A == 0 ? 0 : GSR_E_BADARG;
		n_vis = A; // grid and validation bound; P <= n_vis * n_offsets holds with equality
		P = (int64_t)A * n_offsets;
	}
	if (A == 0 || n_vis == 0 || P == 0) return (A < 0 || n_vis < 0 || P < 0) ? GSR_E_BADARG : 0;
	DecodeArgs a{};
	const int rc = fill_decode_args(a, A, feat_dim, n_offsets, n_vis, P, anchor, anchor_feat, offset, scaling, campos, mlp_params, scratch, scratch_bytes);
	if (rc) return rc;
	if (counts_on_device) {
		const DecodeLayout L = decode_layout(A);
		a.vis_ids = (const uint32_t *)((char *)scratch + L.vis_ids);
		a.n_vis_dev = (const uint32_t *)((char *)scratch + L.vis_incl) + (A - 1);
	}
	if (!neural_opacity || !xyz || !color || !opacity || !uncertainty || !out_scaling || !rot) return GSR_E_BADARG;
	a.neural_opacity = const_cast<float *>(neural_opacity);
	a.out_xyz = xyz; a.out_color = color; a.out_opacity = opacity; a.out_uncertainty = uncertainty; a.out_scaling = out_scaling; a.out_rot = rot;
	StageTimer t(kDecodeFwd, stream);
	GSR_CUDA(decode_stage2(a, stream));
	return 0;
}

int gsr_decode_backward(int A, int feat_dim, int n_offsets, int64_t n_vis, int64_t P, const float *anchor, const float *anchor_feat,
                        const float *offset, const float *scaling, const float *campos, const float *const *mlp_params, void *scratch,
                        size_t scratch_bytes, const float *d_xyz, const float *d_color, const float *d_opacity, const float *d_uncertainty,
                        const float *d_scaling, const float *d_rot, const float *d_neural_opacity, float *g_anchor, float *g_feat,
                        float *g_offset, float *g_scaling, float *const *g_mlp_params, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (A == 0 || n_vis == 0) return (A < 0 || n_vis < 0) ? GSR_E_BADARG : 0;
	DecodeBwdArgs b{};
	const int rc = fill_decode_args(b.f, A, feat_dim, n_offsets, n_vis, P, anchor, anchor_feat, offset, scaling, campos, mlp_params, scratch, scratch_bytes);
	if (rc) return rc;
	if (!g_anchor || !g_feat || !g_offset || !g_scaling || !g_mlp_params || !aligned16(g_feat)) return GSR_E_BADARG;
	for (int m = 0; m < 4; m++) {
		b.g_w1[m] = g_mlp_params[4 * m + 0]; b.g_b1[m] = g_mlp_params[4 * m + 1]; b.g_w2[m] = g_mlp_params[4 * m + 2]; b.g_b2[m] = g_mlp_params[4 * m + 3];
		if (!b.g_w1[m] || !b.g_b1[m] || !b.g_w2[m] || !b.g_b2[m]) return GSR_E_BADARG;
	}
	b.d_xyz = d_xyz; b.d_color = d_color; b.d_opacity = d_opacity; b.d_uncertainty = d_uncertainty; b.d_scaling = d_scaling; b.d_rot = d_rot;
	b.d_neural_opacity = d_neural_opacity;
	b.g_anchor = g_anchor; b.g_feat = g_feat; b.g_offset = g_offset; b.g_scaling = g_scaling;
	StageTimer t(kDecodeBwd, stream);
	GSR_CUDA(decode_backward(b, stream));
	return 0;
}

// ---- densification statistics ------------------------------------------------------------------------------------------------
size_t gsr_training_statis_scratch_bytes(int A, int n_offsets) { return statis_scratch_bytes(A, n_offsets); }

int gsr_training_statis(int A, int n_offsets, int64_t n_vis, int64_t P, const uint8_t *anchor_visible_mask, const uint8_t *offset_selection_mask,
                        const uint8_t *update_filter, const float *neural_opacity, const float *viewspace_grad, float *opacity_accum,
                        float *anchor_demon, float *offset_gradient_accum, float *offset_denom, void *scratch, size_t scratch_bytes,
                        gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (A < 0 || n_offsets < 1 || n_vis < 0 || n_vis > A || P < 0 || P > n_vis * n_offsets) return GSR_E_BADARG;
	if (A == 0 || n_vis == 0) return 0;
	if (!anchor_visible_mask || !offset_selection_mask || !neural_opacity || !opacity_accum || !anchor_demon || !offset_gradient_accum || !offset_denom || !scratch)
		return GSR_E_BADARG;
	if (P > 0 && (!update_filter || !viewspace_grad)) return GSR_E_BADARG;
	if (scratch_bytes < statis_scratch_bytes(A, n_offsets) || !aligned16(scratch)) return GSR_E_WORKSPACE;
	GSR_CUDA(training_statis(A, n_offsets, n_vis, P, anchor_visible_mask, offset_selection_mask, update_filter, neural_opacity, viewspace_grad,
	                         opacity_accum, anchor_demon, offset_gradient_accum, offset_denom, (char *)scratch, stream));
	return 0;
}

// ---- fused L1 + SSIM image loss -------------------------------------------------------------------------------------------
int gsr_l1_ssim_forward(int planes, int height, int width, const float *taps11_host, const float *image, const float *target,
                        const float *mask, int mask_planes, double *sums, float *partials, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (planes <= 0 || height <= 0 || width <= 0 || !taps11_host || !image || !target || !sums) return GSR_E_BADARG;
	if (mask && mask_planes != 1 && mask_planes != planes) return GSR_E_BADARG;
	const size_t n = (size_t)planes * height * width;
	StageTimer t(kLossFwd, stream);
	GSR_CUDA(launch_l1_ssim_forward(planes, height, width, taps11_host, image, target, mask, mask_planes, sums, partials,
	                                partials ? partials + n : nullptr, partials ? partials + 2 * n : nullptr, stream));
	return 0;
}

int gsr_l1_ssim_backward(int planes, int height, int width, const float *taps11_host, const float *image, const float *target,
                         const float *mask, int mask_planes, const float *partials, const float *upstream, float *grad_image,
                         gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (planes <= 0 || height <= 0 || width <= 0 || !taps11_host || !image || !target || !partials || !upstream || !grad_image) return GSR_E_BADARG;
	if (mask && mask_planes != 1 && mask_planes != planes) return GSR_E_BADARG;
	const size_t n = (size_t)planes * height * width;
	StageTimer t(kLossBwd, stream);
	GSR_CUDA(launch_l1_ssim_backward(planes, height, width, taps11_host, image, target, mask, mask_planes, partials, partials + n,
	                                 partials + 2 * n, upstream, grad_image, stream));
	return 0;
}

int gsr_depth_align_l1_forward(int batch, int height, int width, const float *depth, const float *target, const float *fit_mask,
                               const float *loss_mask, double *state, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (batch <= 0 || height <= 0 || width <= 0 || !depth || !target || !state) return GSR_E_BADARG;
	StageTimer t(kLossFwd, stream);
	GSR_CUDA(launch_depth_align_l1_forward(batch, height * width, depth, target, fit_mask, loss_mask, state + 1, state + 1 + 5 * batch, state, stream));
	return 0;
}

int gsr_depth_align_l1_backward(int batch, int height, int width, const float *depth, const float *target, const float *fit_mask,
                                const float *loss_mask, const double *state, const float *upstream, float *grad_depth, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (batch <= 0 || height <= 0 || width <= 0 || !depth || !target || !state || !upstream || !grad_depth) return GSR_E_BADARG;
	StageTimer t(kLossBwd, stream);
	GSR_CUDA(launch_depth_align_l1_backward(batch, height * width, depth, target, fit_mask, loss_mask, state + 1, state + 1 + 5 * batch, upstream,
	                                        grad_depth, stream));
	return 0;
}

int gsr_depth_grad_forward(int batch, int height, int width, int n_scales, const float *prediction, const float *target, const float *mask,
                           const double *fit_state, double *gstate, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (batch <= 0 || height <= 0 || width <= 0 || n_scales < 1 || n_scales > 4 || !prediction || !target || !gstate) return GSR_E_BADARG;
	StageTimer t(kLossFwd, stream);
	GSR_CUDA(launch_depth_grad_forward(batch, height, width, n_scales, prediction, target, mask, fit_state ? fit_state + 1 : nullptr, gstate, stream));
	return 0;
}

int gsr_depth_grad_backward(int batch, int height, int width, int n_scales, const float *prediction, const float *target, const float *mask,
                            const float *fit_mask, const double *fit_state, const double *gstate, const float *upstream, float *grad,
                            int accumulate, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (batch <= 0 || height <= 0 || width <= 0 || n_scales < 1 || n_scales > 4 || !prediction || !target || !gstate || !upstream || !grad)
		return GSR_E_BADARG;
	StageTimer t(kLossBwd, stream);
	GSR_CUDA(launch_depth_grad_backward(batch, height, width, n_scales, prediction, target, mask, fit_mask, fit_state ? fit_state + 1 : nullptr,
	                                    gstate, upstream, grad, accumulate, stream));
	return 0;
}

int gsr_adam_step(int n_tensors, const gsr_adam_tensor *tensors_host, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (n_tensors < 0 || (n_tensors > 0 && !tensors_host)) return GSR_E_BADARG;
	// validate everything before the first launch: a bad descriptor must not leave the model half-updated
	for (int k = 0; k < n_tensors; k++) {
		const gsr_adam_tensor &a = tensors_host[k];
		if (a.numel < 0 || a.step < 1) return GSR_E_BADARG;
		if (a.numel > 0 && (!a.param || !a.grad || !a.exp_avg || !a.exp_avg_sq)) return GSR_E_BADARG;
		if (!(a.beta1 >= 0.0 && a.beta1 < 1.0) || !(a.beta2 >= 0.0 && a.beta2 < 1.0)) return GSR_E_BADARG;
	}
	StageTimer timer(kOptim, stream);
	AdamTensor tab[kAdamMaxTensors];   // no heap: the C ABI never throws
	int m = 0;
	for (int k = 0; k < n_tensors; k++) {
		const gsr_adam_tensor &a = tensors_host[k];
		if (a.numel > 0) {
			AdamTensor &t = tab[m++];
			t.param = a.param; t.grad = a.grad; t.exp_avg = a.exp_avg; t.exp_avg_sq = a.exp_avg_sq; t.n = a.numel;
			const double bc1 = 1.0 - std::pow(a.beta1, (double)a.step), bc2 = 1.0 - std::pow(a.beta2, (double)a.step);
			t.one_minus_beta1 = (float)(1.0 - a.beta1);
			t.beta2 = (float)a.beta2;
			t.one_minus_beta2 = (float)(1.0 - a.beta2);
			t.eps = (float)a.eps;
			t.weight_decay = (float)a.weight_decay;
			t.step_size = (float)(a.lr / bc1);
			t.bias_correction2_sqrt = (float)std::sqrt(bc2);
			t.pad = 0;
		}
		if (m == kAdamMaxTensors || (k == n_tensors - 1 && m > 0)) {
			GSR_CUDA(launch_adam_step(m, tab, stream));
			m = 0;
		}
	}
	return 0;
}

int gsr_debug_export(int P, int64_t num_rendered, int width, int height, const void *geom_buffer, const void *binning_buffer,
                     size_t binning_bytes, const void *image_buffer, float *xy, float *depths, float *conic_opacity, uint32_t *tiles_touched,
                     uint32_t *point_list, uint32_t *ranges, float *final_T, uint32_t *n_contrib, gsr_stream_t stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P <= 0 || width <= 0 || height <= 0) return GSR_E_BADARG;
	const GeomLayout GL = geom_layout(P);
	const ImageLayout IL = image_layout(width, height);
	const int64_t capacity = binning_buffer ? binning_capacity(P, width, height, binning_bytes) : 1;
	if (capacity < 1 || num_rendered > capacity) return GSR_E_WORKSPACE;
	const BinningLayout BL = binning_layout(P, capacity, width, height);
	const char *geom = (const char *)geom_buffer, *binning = (const char *)binning_buffer, *image = (const char *)image_buffer;
	const size_t N = (size_t)width * height;
	const size_t tiles = (size_t)((width + GSR_BLOCK_X - 1) / GSR_BLOCK_X) * ((height + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y);
	if (geom && (xy || depths || conic_opacity)) {
		export_records_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, (const float *)(geom + GL.rec), xy, depths, conic_opacity);
		GSR_CUDA(cudaGetLastError());
	}
	if (geom && tiles_touched) GSR_CUDA(cudaMemcpyAsync(tiles_touched, geom + GL.tiles_touched, (size_t)P * 4, cudaMemcpyDeviceToDevice, stream));
	if (binning && point_list && num_rendered > 0) {
		export_point_list_kernel<<<(unsigned)((num_rendered + 255) / 256), 256, 0, stream>>>(
		    num_rendered, (const uint32_t *)(binning + BL.val[point_list_index(width, height)]), (const uint32_t *)(binning + BL.header), point_list);
		GSR_CUDA(cudaGetLastError());
	}
	if (image && ranges) GSR_CUDA(cudaMemcpyAsync(ranges, image + IL.ranges, tiles * 8, cudaMemcpyDeviceToDevice, stream));
	if (image && final_T) GSR_CUDA(cudaMemcpyAsync(final_T, image + IL.final_T, N * 4, cudaMemcpyDeviceToDevice, stream));
	if (image && n_contrib) GSR_CUDA(cudaMemcpyAsync(n_contrib, image + IL.n_contrib, N * 4, cudaMemcpyDeviceToDevice, stream));
	return 0;
}

} // extern "C"
