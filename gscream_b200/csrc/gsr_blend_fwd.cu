// gsr_blend_fwd.cu — forward tile blend: per-16x16-tile front-to-back alpha compositing of C colour /
// feature channels + depth + uncertainty.
//
// Replaces renderCUDA<C> forward (CR/forward.cu:441-568 of W-Ted/GScream's
// submodules/diff-gaussian-rasterization) with identical per-pixel semantics:
//   power = -0.5 (a dx^2 + c dy^2) - b dx dy ; skip if power > 0
//   alpha = min(0.99, opacity * exp(power))  ; skip if alpha < 1/255
//   if T (1 - alpha) < 1e-4 the pixel is done and this Gaussian is NOT blended
//   n_contrib = 1-based list position of the last blended Gaussian; colour gets T*bg, depth and
//   uncertainty do not.
// What is different (B200-first):
//   * the tile's slab (64-B projected records, plus a C*4-B feature row for C > 3) is gathered into
//     shared memory asynchronously (16-B cp.async chunks, double-buffered, one barrier per 128-Gaussian
//     round) — no register staging, features included (the reference re-reads colour and depth from global
//     memory for every contributing pixel, forward.cu:545-546).  Round 1 first used per-Gaussian bulk copies
//     (cp.async.bulk / UBLKCP + mbarrier); ncu showed their uniform-register operands serialise a warp's 32
//     gathers into 32 issue rounds (13 % of this kernel's instructions), see profiles/ and DESIGN.md;
//   * each warp owns an 8x4 pixel block and first compacts the staged batch down to the Gaussians whose
//     alpha >= 1/255 bounding box touches its block (conservative, computed in preprocess), so the
//     per-pair work is only spent where a contribution is possible.  The skipped pairs are exactly pairs
//     the reference `continue`s over, so results and n_contrib are unchanged;
//   * Gaussian ids two batches ahead and slabs one batch ahead are in flight while the current one is blended.
#include "gsr_blend.cuh"
#include "gsr_internal.cuh"

namespace gsr {

// Resident CTAs per SM the register allocation must allow.  Measured (profiles/r1_occupancy_ab.md): the forward
// kernel is latency bound, 4 CTAs/SM (64 registers, 12 B of spill at C = 32) beats 3 CTAs/SM (80 registers) by 10 %.
#ifndef GSR_FWD_MINBLOCKS
#define GSR_FWD_MINBLOCKS 4
#endif
template <int C>
__global__ void __launch_bounds__(256, GSR_FWD_MINBLOCKS) blend_forward_kernel(
    const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, int W, int H, int tiles_x,
    const float *__restrict__ rec, const float *__restrict__ features, const float *__restrict__ bg,
    float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
    float *__restrict__ out_color, float *__restrict__ out_depth, float *__restrict__ out_unc)
{
	using TR = BlendTraits<C>;
	constexpr bool kBulk = GSR_FWD_BULK != 0;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	uint32_t *s_ids = reinterpret_cast<uint32_t *>(smem_raw + TR::kIdsOff);  // [3][kBatch]
	uint8_t *s_mask = smem_raw + TR::kMaskOff;                               // [3][kBatch]
	uint8_t *s_list = smem_raw + TR::kListOff;                               // [8][kBatch]
	auto stage_rec = [&](int s) { return reinterpret_cast<float *>(smem_raw + (size_t)s * TR::kStageBytes); };
	auto stage_feat = [&](int s) { return reinterpret_cast<float *>(smem_raw + (size_t)s * TR::kStageBytes + (size_t)kBatch * GSR_REC_BYTES); };

	__shared__ __align__(8) uint64_t s_bar[kStages];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int tile = blockIdx.x;
	stage_init<C, kBulk>(s_bar, tid);
	const int tile_x0 = (tile % tiles_x) * GSR_BLOCK_X, tile_y0 = (tile / tiles_x) * GSR_BLOCK_Y;
	int bx, by;
	warp_block_origin(warp, bx, by);
	const int px = tile_x0 + bx + (lane & 7), py = tile_y0 + by + (lane >> 3);
	const bool inside = px < W && py < H;
	const float pixf_x = (float)px, pixf_y = (float)py;

	const uint2 range = ranges[tile];
	const int total = (int)(range.y - range.x);
	const int rounds = (total + kBatch - 1) / kBatch;

	float T = 1.0f;
	uint32_t last_contributor = 0;
	float acc[C];
#pragma unroll
	for (int ch = 0; ch < C; ch++) acc[ch] = 0.f;
	float D = 0.f, UNC = 0.f;
	bool done = !inside;

	// Staging, per batch b of kBatch list entries (threads < kBatch own one entry each):
	//   phase 1: id -> s_ids[b%3]; bounding extents from L2 -> per-warp overlap mask -> s_mask[b%3]
	//   phase 2 (all threads, after a barrier): 16-B cp.async chunks of the records / feature rows -> stage b%2
	auto phase1 = [&](int b, uint32_t id) {
		const int base = b * kBatch;
		uint32_t mask = 0;
		if (tid < kBatch) {
			if (base + tid < total) {
				const float *src = rec + (size_t)id * GSR_REC_FLOATS;
				const float2 cxy = __ldg(reinterpret_cast<const float2 *>(src));
				const float2 ext = __ldg(reinterpret_cast<const float2 *>(src + 8));
				mask = warp_overlap_mask(cxy.x, cxy.y, ext.x, ext.y, (float)tile_x0, (float)tile_y0);
				s_ids[(b % kIdStages) * kBatch + tid] = id;
			}
			s_mask[(b % kIdStages) * kBatch + tid] = (uint8_t)mask;
		}
	};
	auto load_id = [&](int b) -> uint32_t {
		const int i = b * kBatch + tid;
		return (tid < kBatch && i < total) ? point_list[range.x + i] : 0u;
	};

	// prologue: batch 0 fully issued, id of batch 1 in flight
	uint32_t next_id = load_id(0);
	if (rounds > 0) {
		phase1(0, next_id);
		__syncthreads();
		stage_issue<C, kBulk>(&s_bar[0], stage_rec(0), stage_feat(0), s_ids, next_id, min(kBatch, total), rec, features, tid);
	}
	next_id = load_id(1);

	for (int r = 0; r < rounds; r++) {
		if (r + 1 < rounds) phase1(r + 1, next_id);
		stage_wait<kBulk>(&s_bar[r & 1], r >> 1);
		// one barrier per round: batch r has landed for everyone, ids of batch r+1 are visible, and every warp
		// is past the previous batch's reads.  It doubles as the whole-tile early exit (CR/forward.cu:496-498).
		if (__syncthreads_count(done) == 256) break;
		if (r + 1 < rounds)
			stage_issue<C, kBulk>(&s_bar[(r + 1) & 1], stage_rec((r + 1) & 1), stage_feat((r + 1) & 1), s_ids + ((r + 1) % kIdStages) * kBatch, next_id,
			               min(kBatch, total - (r + 1) * kBatch), rec, features, tid);
		next_id = load_id(r + 2);

		const int base = r * kBatch;
		const int count = min(kBatch, total - base);
		const float *s_rec = stage_rec(r & 1);
		const float *s_feat = stage_feat(r & 1);
		uint8_t *my_list = s_list + warp * kBatch;
		if (__all_sync(0xffffffffu, done)) continue;
		const int n = build_warp_list(s_mask + (r % kIdStages) * kBatch, my_list, warp, lane, count);

		for (int k = 0; k < n; k++) {
			const int j = my_list[k];
			const float4 r0 = *reinterpret_cast<const float4 *>(s_rec + j * GSR_REC_FLOATS);     // x y a b
			const float4 r1 = *reinterpret_cast<const float4 *>(s_rec + j * GSR_REC_FLOATS + 4); // c o depth unc
			const float2 d = {r0.x - pixf_x, r0.y - pixf_y};
			const float power = gaussian_power(r0.z, r0.w, r1.x, d.x, d.y);
			if (done || power > 0.0f) continue;
			const float alpha = min(0.99f, __fmul_rn(r1.y, expf(power)));
			if (alpha < kAlphaMin) continue;
			const float test_T = __fmul_rn(T, __fsub_rn(1.f, alpha));
			if (test_T < 0.0001f) {
				done = true;
				continue;
			}
			const float w = alpha * T;
			if (TR::kFeatInRec) {
				const float4 r2 = *reinterpret_cast<const float4 *>(s_rec + j * GSR_REC_FLOATS + 8); // hx hy r g
				const float cb = s_rec[j * GSR_REC_FLOATS + 12];
				if (C > 0) acc[0] += r2.z * w;
				if (C > 1) acc[1 % C] += r2.w * w;
				if (C > 2) acc[2 % C] += cb * w;
			} else {
				const float4 *f4 = reinterpret_cast<const float4 *>(s_feat + j * C);
#pragma unroll
				for (int q = 0; q < C / 4; q++) {
					const float4 f = f4[q];
					acc[4 * q + 0] += f.x * w;
					acc[4 * q + 1] += f.y * w;
					acc[4 * q + 2] += f.z * w;
					acc[4 * q + 3] += f.w * w;
				}
			}
			D += r1.z * w;
			UNC += r1.w * w;
			T = test_T;
			last_contributor = (uint32_t)(base + j + 1);
		}
	}
	stage_drain<kBulk>(); // nothing may be in flight into shared memory when the CTA retires

	if (inside) {
		const size_t pix_id = (size_t)W * py + px;
		final_T[pix_id] = T;
		n_contrib[pix_id] = last_contributor;
		const size_t plane = (size_t)H * W;
#pragma unroll
		for (int ch = 0; ch < C; ch++) out_color[ch * plane + pix_id] = acc[ch] + T * bg[ch];
		out_depth[pix_id] = D;
		out_unc[pix_id] = UNC;
	}
}

template <int C>
static cudaError_t launch_fwd(int tiles, const uint2 *ranges, const uint32_t *point_list, int W, int H, int tiles_x, const float *rec,
                              const float *features, const float *bg, float *final_T, uint32_t *n_contrib, float *out_color,
                              float *out_depth, float *out_unc, cudaStream_t stream)
{
	using TR = BlendTraits<C>;
	static bool configured = false;
	if (!configured) {
		cudaError_t e = cudaFuncSetAttribute(blend_forward_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TR::kSmemBytes);
		if (e != cudaSuccess) return e;
		configured = true;
	}
	blend_forward_kernel<C><<<tiles, 256, TR::kSmemBytes, stream>>>(ranges, point_list, W, H, tiles_x, rec, features, bg, final_T, n_contrib,
	                                                               out_color, out_depth, out_unc);
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_blend_forward(int C, int W, int H, const uint2 *ranges, const uint32_t *point_list, const float *rec,
                                 const float *features, const float *bg, float *final_T, uint32_t *n_contrib, float *out_color,
                                 float *out_depth, float *out_unc, cudaStream_t stream)
{
	const int tiles_x = (W + GSR_BLOCK_X - 1) / GSR_BLOCK_X, tiles_y = (H + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y;
	const int tiles = tiles_x * tiles_y;
	if (tiles <= 0) return cudaSuccess;
	switch (C) {
	case 3: return launch_fwd<3>(tiles, ranges, point_list, W, H, tiles_x, rec, features, bg, final_T, n_contrib, out_color, out_depth, out_unc, stream);
	case 32: return launch_fwd<32>(tiles, ranges, point_list, W, H, tiles_x, rec, features, bg, final_T, n_contrib, out_color, out_depth, out_unc, stream);
	default: return cudaErrorInvalidValue;
	}
}

} // namespace gsr
