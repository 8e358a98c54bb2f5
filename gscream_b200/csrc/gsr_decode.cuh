// gsr_decode.cuh — declarations of the fused anchor -> neural-Gaussian decode (gsr_decode.cu), shared with gsr_api.cu.
#pragma once
#include "gsr_common.cuh"

namespace gsr {

constexpr int kDecNA = 16;    // anchors per warp tile (the M dimension of the MLP products)
constexpr int kDecMaxK = 16;  // largest supported n_offsets

// The four MLPs of scene/gaussian_model.py:118-144 in torch.nn.Linear layout (weight[out][in], bias[out]):
// index 0 opacity, 1 uncertainty, 2 cov, 3 colour.
struct DecodeWeights {
	const float *w1[4], *b1[4], *w2[4], *b2[4];
};

struct DecodeLayout {  // caller-allocated scratch, sized by gsr_decode_scratch_bytes(A)
	size_t vis_incl, vis_ids, count, maskbits, gauss_incl, scan_tmp, total;
};
DecodeLayout decode_layout(int A);

struct DecodeArgs {
	int k;                       // n_offsets
	int n_vis;                   // visible anchors (host value; upper bound A in stage 1)
	const uint32_t *n_vis_dev;   // device copy of n_vis (null: n_vis is exact); stage 1, and stage 2 when launched before the host read the counts
	const uint32_t *vis_ids;     // [n_vis] visible anchor indices, ascending (null: identity)
	const float *anchor, *feat, *offset, *scaling, *campos;
	DecodeWeights wt;
	float *neural_opacity;       // [n_vis * k]   tanh output of the opacity MLP (stage 1 writes, stage 2 reads)
	uint8_t *mask;               // [n_vis * k]   neural_opacity > 0
	uint32_t *count, *maskbits;  // [n_vis]       per-anchor kept offsets: count and bit mask
	const uint32_t *gauss_incl;  // [n_vis]       inclusive scan of count
	float *out_xyz, *out_color, *out_opacity, *out_uncertainty, *out_scaling, *out_rot;
};

struct DecodeBwdArgs {
	DecodeArgs f;
	// upstream gradients (any may be null = zero): per kept Gaussian, and for the full neural_opacity vector
	const float *d_xyz, *d_color, *d_opacity, *d_uncertainty, *d_scaling, *d_rot, *d_neural_opacity;
	// outputs: g_anchor[A,3] g_feat[A,32] g_offset[A,k,3] g_scaling[A,6] must arrive zero-filled (rows of invisible anchors
	// and culled offsets are not touched); the sixteen weight / bias gradients are accumulated into (atomicAdd)
	float *g_anchor, *g_feat, *g_offset, *g_scaling;
	float *g_w1[4], *g_b1[4], *g_w2[4], *g_b2[4];
};

cudaError_t decode_stage1(int A, int k, const float *anchor, const float *feat, const uint8_t *visible_mask, const float *campos,
                          const DecodeWeights &wt, char *scratch, const DecodeLayout &L, float *neural_opacity, uint8_t *mask,
                          int64_t *counts_host, cudaStream_t stream);
cudaError_t decode_stage2(const DecodeArgs &a, cudaStream_t stream);
cudaError_t decode_backward(const DecodeBwdArgs &a, cudaStream_t stream);

// densification statistics (GaussianModel.training_statis)
size_t statis_scratch_bytes(int A, int k);
cudaError_t training_statis(int A, int k, int64_t n_vis, int64_t P, const uint8_t *anchor_visible, const uint8_t *offset_selected, const uint8_t *update_filter,
                            const float *neural_opacity, const float *viewspace_grad, float *opacity_accum, float *anchor_demon,
                            float *offset_gradient_accum, float *offset_denom, char *scratch, cudaStream_t stream);

} // namespace gsr
