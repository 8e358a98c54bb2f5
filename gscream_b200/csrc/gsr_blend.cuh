// gsr_blend.cuh — pieces shared by the forward and backward tile-blend kernels.
#pragma once
#include "gsr_common.cuh"

namespace gsr {

#ifndef GSR_BATCH
#define GSR_BATCH 128
#endif
constexpr int kBatch = GSR_BATCH;    // Gaussians staged per round
constexpr int kStages = 2;           // double-buffered slabs (cp.async groups)
constexpr int kIdStages = 3;         // ids / masks are triple-buffered so one barrier per round suffices
constexpr int kWarpsPerTile = 8;     // 256 threads; warp w owns an 8x4 pixel block of the 16x16 tile
constexpr float kAlphaMin = 1.0f / 255.0f;

template <int C>
struct BlendTraits {
	// C <= 3: colours ride in the 64-B record (slots 10..12); otherwise one row of colors_precomp
	// (C*4 bytes, a multiple of 16) is gathered next to the record.
	static constexpr bool kFeatInRec = (C <= 3);
	static constexpr int kFeatFloats = kFeatInRec ? 0 : C;
	static constexpr uint32_t kBytesPerGaussian = GSR_REC_BYTES + kFeatFloats * 4;
	static constexpr size_t kStageBytes = (size_t)kBatch * kBytesPerGaussian;
	// dynamic shared memory: kStages x [rec kBatch x 64 B][feat kBatch x C*4 B], then
	// [ids 3 x kBatch u32][masks 3 x kBatch u8][warp lists 8 x kBatch u8]
	static constexpr size_t kIdsOff = kStages * kStageBytes;
	static constexpr size_t kMaskOff = kIdsOff + (size_t)kIdStages * kBatch * 4;
	static constexpr size_t kListOff = kMaskOff + (size_t)kIdStages * kBatch;
	static constexpr size_t kSmemBytes = kListOff + (size_t)kWarpsPerTile * kBatch;
};

// 16-byte asynchronous global->shared copy (Ampere-style cp.async, SASS LDGSTS; L2 only, no L1 allocation).
// Measured against per-Gaussian bulk copies (cp.async.bulk / UBLKCP): a bulk copy takes its operands from
// uniform registers, so 32 lanes gathering 32 different rows serialise into 32 elect/R2UR/UBLKCP rounds
// (16 issue slots per Gaussian, 13 % of the forward kernel's instructions in profiles/r1_blend_v1), whereas
// one LDGSTS moves 512 B for the whole warp.  See DESIGN.md section "staging".
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Slab staging of one batch (`count` Gaussians) into shared-memory stage buffers.  Two interchangeable engines:
//   BULK = true : one cp.async.bulk (TMA, SASS UBLKCP) per record and per feature row, issued by the thread that
//                 owns the Gaussian, completing on the stage's mbarrier (expect_tx by thread 0).  Uses the async
//                 proxy: no LSU / L1 involvement, but each copy takes its operands from uniform registers, so a
//                 warp's 32 gathers serialise into 32 elect/R2UR/UBLKCP rounds (~16 issue slots per Gaussian).
//   BULK = false: 16-B cp.async (SASS LDGSTS) chunks assigned to threads in order: 12 instructions per thread per
//                 128-Gaussian batch at C = 32, but they travel through the LSU pipe that the inner loops' LDS also use.
// Both were measured (profiles/r1_*): forward is indifferent, backward prefers BULK by ~5 %.
// Defaults chosen from the A/B in profiles/r1_staging_ab.md: forward -> LDGSTS, backward -> bulk (TMA).
#ifndef GSR_FWD_BULK
#define GSR_FWD_BULK 0
#endif
#ifndef GSR_BWD_BULK
#define GSR_BWD_BULK 1
#endif

template <int C, bool kStageBulk>
__device__ __forceinline__ void stage_init(uint64_t *bars, int tid)
{
	if (kStageBulk) {
		if (tid == 0) {
			for (int s = 0; s < kStages; s++) mbar_init(&bars[s], 1);
			mbar_fence_init();
		}
	}
}
// `my_id` is the Gaussian owned by thread tid (< count); `s_ids` holds the same ids for the chunked engine.
template <int C, bool kStageBulk>
__device__ __forceinline__ void stage_issue(uint64_t *bar, float *s_rec, float *s_feat, const uint32_t *s_ids, uint32_t my_id, int count,
                                            const float *__restrict__ rec, const float *__restrict__ features, int tid)
{
	using TR = BlendTraits<C>;
	if (kStageBulk) {
		if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)count * TR::kBytesPerGaussian);
		if (tid < count) {
			bulk_g2s(s_rec + tid * GSR_REC_FLOATS, rec + (size_t)my_id * GSR_REC_FLOATS, GSR_REC_BYTES, bar);
			if constexpr (!TR::kFeatInRec) bulk_g2s(s_feat + tid * C, features + (size_t)my_id * C, C * 4, bar);
		}
	} else {
#pragma unroll
		for (int k = 0; k < (kBatch * 4) / 256; k++) {
			const int q = tid + 256 * k, row = q >> 2, part = q & 3;
			if (row < count) cp_async16(s_rec + row * GSR_REC_FLOATS + part * 4, rec + (size_t)s_ids[row] * GSR_REC_FLOATS + part * 4);
		}
		if constexpr (!TR::kFeatInRec) {
			constexpr int kChunksPerRow = C / 4;
#pragma unroll
			for (int k = 0; k < (kBatch * kChunksPerRow + 255) / 256; k++) {
				const int q = tid + 256 * k, row = q / kChunksPerRow, part = q % kChunksPerRow;
				if (row < count) cp_async16(s_feat + row * C + part * 4, features + (size_t)s_ids[row] * C + part * 4);
			}
		}
		cp_async_commit();
	}
}
// Wait until the batch staged as the `use`-th use of this stage buffer has landed (for this thread's view;
// the caller's block barrier publishes it to everyone in the chunked engine).
template <bool kStageBulk>
__device__ __forceinline__ void stage_wait(uint64_t *bar, int use)
{
	if (kStageBulk) mbar_wait(bar, (uint32_t)(use & 1));
	else cp_async_wait_all();
}
template <bool kStageBulk>
__device__ __forceinline__ void stage_drain()
{
	if (!kStageBulk) cp_async_wait_all();
}

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
