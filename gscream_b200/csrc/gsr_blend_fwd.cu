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
//     shared memory by per-Gaussian bulk-async copies (cp.async.bulk -> SASS UBLKCP, the non-tensor TMA
//     path) completing on an mbarrier — no register staging, features included (the reference re-reads
//     colour and depth from global memory for every contributing pixel, forward.cu:545-546);
//   * each warp owns an 8x4 pixel block and first compacts the staged batch down to the Gaussians whose
//     alpha >= 1/255 bounding box touches its block (conservative, computed in preprocess), so the
//     per-pair work is only spent where a contribution is possible.  The skipped pairs are exactly pairs
//     the reference `continue`s over, so results and n_contrib are unchanged;
//   * Gaussian ids for the next batch are prefetched while the current one is blended.
#include "gsr_blend.cuh"

namespace gsr {

template <int C>
__global__ void __launch_bounds__(256) blend_forward_kernel(
    const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, int W, int H, int tiles_x,
    const float *__restrict__ rec, const float *__restrict__ features, const float *__restrict__ bg,
    float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
    float *__restrict__ out_color, float *__restrict__ out_depth, float *__restrict__ out_unc)
{
	using TR = BlendTraits<C>;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	float *s_rec = reinterpret_cast<float *>(smem_raw);                                   // [256][16]
	float *s_feat = reinterpret_cast<float *>(smem_raw + (size_t)kBatch * GSR_REC_BYTES); // [256][C]
	uint8_t *s_list = smem_raw + TR::kStageBytes;                                         // [8][256]
	uint8_t *s_mask = s_list + kWarpsPerTile * kBatch;                                    // [256]
	__shared__ __align__(8) uint64_t s_bar;

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int tile = blockIdx.x;
	const int tile_x0 = (tile % tiles_x) * GSR_BLOCK_X, tile_y0 = (tile / tiles_x) * GSR_BLOCK_Y;
	int bx, by;
	warp_block_origin(warp, bx, by);
	const int px = tile_x0 + bx + (lane & 7), py = tile_y0 + by + (lane >> 3);
	const bool inside = px < W && py < H;
	const float pixf_x = (float)px, pixf_y = (float)py;

	const uint2 range = ranges[tile];
	const int total = (int)(range.y - range.x);
	const int rounds = (total + kBatch - 1) / kBatch;

	if (tid == 0) {
		mbar_init(&s_bar, 1);
		mbar_fence_init();
	}

	float T = 1.0f;
	uint32_t last_contributor = 0;
	float acc[C];
#pragma unroll
	for (int ch = 0; ch < C; ch++) acc[ch] = 0.f;
	float D = 0.f, UNC = 0.f;
	bool done = !inside;

	uint32_t next_id = (tid < total) ? point_list[range.x + tid] : 0u;
	__syncthreads(); // barrier init visible

	for (int r = 0; r < rounds; r++) {
		// whole tile finished? (CR/forward.cu:496-498)
		if (__syncthreads_count(done) == 256) break; // also: everyone is past the previous batch's smem reads
		const int base = r * kBatch;
		const int count = min(kBatch, total - base);

		// ---- stage this batch: one bulk-async gather per Gaussian, issued by its thread ----
		if (tid == 0) mbar_arrive_expect_tx(&s_bar, (uint32_t)count * TR::kBytesPerGaussian);
		uint32_t mask = 0;
		if (tid < count) {
			const uint32_t id = next_id;
			const float *src = rec + (size_t)id * GSR_REC_FLOATS;
			bulk_g2s(s_rec + tid * GSR_REC_FLOATS, src, GSR_REC_BYTES, &s_bar);
			if (!TR::kFeatInRec) bulk_g2s(s_feat + tid * C, features + (size_t)id * C, C * 4, &s_bar);
			// bounding extents straight from L2 (one 8-B + one 8-B load; conflict-free, unlike a strided smem read)
			const float2 cxy = __ldg(reinterpret_cast<const float2 *>(src));
			const float2 ext = __ldg(reinterpret_cast<const float2 *>(src + 8));
			mask = warp_overlap_mask(cxy.x, cxy.y, ext.x, ext.y, (float)tile_x0, (float)tile_y0);
		}
		s_mask[tid] = (uint8_t)mask;
		// prefetch the ids of the next batch
		next_id = (base + kBatch + tid < total) ? point_list[range.x + base + kBatch + tid] : 0u;
		__syncthreads(); // masks visible
		uint8_t *my_list = s_list + warp * kBatch;
		const int n = build_warp_list(s_mask, my_list, warp, lane, count);
		mbar_wait(&s_bar, (uint32_t)(r & 1)); // slab has landed

		if (__all_sync(0xffffffffu, done)) continue;
		for (int k = 0; k < n; k++) {
			const int j = my_list[k];
			const float4 r0 = *reinterpret_cast<const float4 *>(s_rec + j * GSR_REC_FLOATS);     // x y a b
			const float4 r1 = *reinterpret_cast<const float4 *>(s_rec + j * GSR_REC_FLOATS + 4); // c o depth unc
			// same expression as CR/forward.cu:521-525
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
