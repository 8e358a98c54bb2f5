// gsr_blend.cuh — pieces shared by the forward and backward tile-blend kernels.
#pragma once
#include "gsr_common.cuh"

namespace gsr {

constexpr int kBatch = 256;          // Gaussians staged per round (one per thread)
constexpr int kWarpsPerTile = 8;     // 256 threads; warp w owns an 8x4 pixel block of the 16x16 tile
constexpr float kAlphaMin = 1.0f / 255.0f;

template <int C>
struct BlendTraits {
	// C <= 3: colours ride in the 64-B record (slots 10..12); otherwise one row of colors_precomp
	// (C*4 bytes, must be a multiple of 16 for cp.async.bulk) is gathered next to the record.
	static constexpr bool kFeatInRec = (C <= 3);
	static constexpr int kFeatFloats = kFeatInRec ? 0 : C;
	static constexpr uint32_t kBytesPerGaussian = GSR_REC_BYTES + kFeatFloats * 4;
	static constexpr size_t kStageBytes = (size_t)kBatch * kBytesPerGaussian;
	// dynamic shared memory: [rec 256x64B][feat 256xC*4B][warp lists 8x256B][masks 256B]
	static constexpr size_t kSmemBytes = kStageBytes + kWarpsPerTile * kBatch + kBatch;
};

// power = -0.5 (a dx^2 + c dy^2) - b dx dy, CR/forward.cu:524 / CR/backward.cu:524, with the rounding
// sequence of the reference build's SASS (same in renderCUDA forward and backward, C = 3 and 32):
//   s = fma(dx, a*dx, (c*dy)*dy) ; power = fma(s, -0.5, -((b*dx)*dy))
// pinned with explicit intrinsics: a 1-ulp change of alpha can flip the 1/255 and 1e-4 threshold tests
// and with them n_contrib.
__device__ __forceinline__ float gaussian_power(float a, float b, float c, float dx, float dy)
{
	const float s = __fmaf_rn(dx, __fmul_rn(a, dx), __fmul_rn(__fmul_rn(c, dy), dy));
	return __fmaf_rn(s, -0.5f, -__fmul_rn(__fmul_rn(b, dx), dy));
}

// Pixel owned by (warp, lane): warps tile the 16x16 block as 2 (x) by 4 (y) blocks of 8x4 pixels.
__device__ __forceinline__ void warp_block_origin(int warp, int &bx, int &by)
{
	bx = (warp & 1) * 8;
	by = (warp >> 1) * 4;
}

// One bit per warp: does the bounding box {|x - cx| <= hx, |y - cy| <= hy} of the Gaussian's
// alpha >= 1/255 region touch that warp's 8x4 pixel block?  (Pixel centres are integer coordinates,
// CR/forward.cu:466.)  hx < 0 encodes "never contributes"; +inf encodes "never cull".
__device__ __forceinline__ uint32_t warp_overlap_mask(float cx, float cy, float hx, float hy, float tile_x0, float tile_y0)
{
	const float lo_x = cx - hx, hi_x = cx + hx, lo_y = cy - hy, hi_y = cy + hy;
	uint32_t mx = 0, my = 0;
#pragma unroll
	for (int i = 0; i < 2; i++) {
		const float x0 = tile_x0 + 8.f * i;
		if (hi_x >= x0 && lo_x <= x0 + 7.f) mx |= 1u << i;
	}
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const float y0 = tile_y0 + 4.f * i;
		if (hi_y >= y0 && lo_y <= y0 + 3.f) my |= 1u << i;
	}
	if (!(hx >= 0.f)) return 0; // negative extent: opacity < 1/255, alpha can never reach the threshold
	uint32_t m = 0;
#pragma unroll
	for (int w = 0; w < 8; w++)
		if (((mx >> (w & 1)) & 1u) && ((my >> (w >> 1)) & 1u)) m |= 1u << w;
	return m;
}

// Build this warp's ordered list of staged Gaussians whose mask has the warp's bit set.
// Returns the list length.  s_mask[kBatch] was written by all threads before a __syncthreads().
__device__ __forceinline__ int build_warp_list(const uint8_t *s_mask, uint8_t *s_list_w, int warp, int lane, int count)
{
	int n = 0;
#pragma unroll
	for (int k = 0; k < kBatch / 32; k++) {
		const int j = k * 32 + lane;
		const bool hit = (j < count) && ((s_mask[j] >> warp) & 1u);
		const uint32_t b = __ballot_sync(0xffffffffu, hit);
		if (hit) s_list_w[n + __popc(b & ((1u << lane) - 1u))] = (uint8_t)j;
		n += __popc(b);
	}
	__syncwarp();
	return n;
}

} // namespace gsr
