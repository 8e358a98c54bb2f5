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
//   * a tile's 8 warps never synchronise: each warp owns an 8x4 pixel block and feeds itself
//     (gsr_blend.cuh: WarpFeed) — it scans the tile's list, keeps only the instances whose alpha >= 1/255
//     bounding box touches its block (an 8-bit mask precomputed per instance by the binning stage), gathers their
//     32-B projected records and feature rows into its private double-buffered shared-memory stage with 16-B
//     cp.async one chunk ahead, and runs the reference's per-pixel recurrence over the landed chunk.  The skipped
//     pairs are exactly pairs the reference `continue`s over, so results and n_contrib are unchanged;
//   * a warp stops as soon as its 32 pixels are saturated (the reference stops per tile, CR/forward.cu:496-498);
//   * C = 32: the recurrence only produces the weights w = alpha T of a 16-entry chunk; the accumulation
//     out[p][ch] += sum_e w[p][e] f[e][ch] — 32 FFMA + 8 broadcast LDS.128 per (lane, entry) on the FP32 pipe, the
//     reference re-reads the features from global memory instead (forward.cu:545-546) — is a [32 x 16] x [16 x 32]
//     product per chunk on the tensor pipe (3xTF32 mma.sync, gsr_blend.cuh), accumulators in fragment layout.
// History of the staging engine (per-Gaussian bulk copies -> block-wide LDGSTS slabs -> warp-private feeds) with the
// measurements that drove it: DESIGN.md section 4, profiles/r1_staging_ab.md, profiles/r1_feed_ab.md; the tensor-pipe
// accumulation: profiles/r2_blend_mma.md.
#include "gsr_blend.cuh"
#include "gsr_internal.cuh"

namespace gsr {

constexpr int kWarpsPerCta = GSR_FWD_WARPS_PER_CTA;
constexpr int kCtasPerTile = kWarpsPerTile / kWarpsPerCta;

// Instances this warp examined but none of its 32 pixels blended: clear the warp's bit in the instance's mask, so that
// the backward pass (which scans the same list with the same masks) visits exactly the contributing (warp, instance) pairs
// — 30 % fewer than the bounding-box candidates.  Only this warp tests this bit, and only before clearing it.
template <class Feed>
__device__ __forceinline__ void clear_unblended(const Feed &feed, uint32_t *list, int warp, int lane, int m_cur, uint32_t blended)
{
	const uint32_t any_blended = __reduce_or_sync(0xffffffffu, blended);
	if (lane < m_cur && !((any_blended >> lane) & 1u))
		atomicAnd(list + feed.q_pos[(feed.done + lane) & (kRing - 1)], ~(1u << (24 + warp)));
}

// ---- small C (colours ride in the projected record): FP32 pipe ------------------------------------------------------------
// Resident CTAs per SM the register allocation must allow.  Measured (profiles/r1_occupancy_ab.md): the kernel is latency
// bound, 32 warps / SM (64 registers) beat 24 (80 registers) by 10 %.
#ifndef GSR_FWD_MINCTAS
#define GSR_FWD_MINCTAS (4 * kCtasPerTile)
#endif
template <int C>
__global__ void __launch_bounds__(32 * kWarpsPerCta, GSR_FWD_MINCTAS) blend_forward_kernel(
    const uint2 *__restrict__ ranges, uint32_t *point_list, const uint32_t *__restrict__ header, int W, int H, int tiles_x,
    const float *__restrict__ rec, const float *__restrict__ bg,
    float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
    float *__restrict__ out_color, float *__restrict__ out_depth, float *__restrict__ out_unc)
{
	using TR = BlendTraits<C>;
	static_assert(TR::kFeatInRec, "this kernel takes the colours from the record; C = 32 has its own kernel below");
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int tid = threadIdx.x, lwarp = tid >> 5, lane = tid & 31;
	const int tile = blockIdx.x / kCtasPerTile;
	const int warp = (blockIdx.x % kCtasPerTile) * kWarpsPerCta + lwarp; // this warp's 8x4 pixel block within the tile
	const int tile_x0 = (tile % tiles_x) * GSR_BLOCK_X, tile_y0 = (tile / tiles_x) * GSR_BLOCK_Y;
	int bx, by;
	warp_block_origin(warp, bx, by);
	const int px = tile_x0 + bx + (lane & 7), py = tile_y0 + by + (lane >> 3);
	const bool inside = px < W && py < H;
	const float pixf_x = (float)px, pixf_y = (float)py;

	const uint2 range = ranges[tile];
	const int packed = (int)__ldg(header + kHdrPacked); // format of the list entries, recorded by the instance emission
	float T = 1.0f;
	uint32_t last_contributor = 0; // 1-based list position of the last blended Gaussian
	uint32_t last_ring = 0;       // ... as 1 + ring index while its chunk is being blended
	float acc[C];
#pragma unroll
	for (int ch = 0; ch < C; ch++) acc[ch] = 0.f;
	float D = 0.f, UNC = 0.f;
	bool done = !inside;

	if (!__all_sync(0xffffffffu, done)) {
		WarpFeed<C, false> feed;
		feed.init(smem_raw + (size_t)lwarp * TR::kWarpBytes, point_list + range.x, (int)(range.y - range.x), rec, nullptr, warp, lane, packed != 0);
		feed.fill();
		int m_cur = feed.issue(0);
		for (int chunk = 0; m_cur > 0; chunk++) {
			feed.fill();
			const int m_next = feed.issue((chunk + 1) & 1);
			feed.wait();
			__syncwarp(); // every lane's copies of this chunk have landed
			const float *ent = feed.stage + (chunk & 1) * TR::kStageFloats;
			uint32_t blended = 0, ebit = 1u; // bit e: this pixel blended entry e of the chunk
			for (int e = 0; e < m_cur; e++, ent += TR::kEntryFloats, ebit <<= 1) {
				const float4 r0 = *reinterpret_cast<const float4 *>(ent);     // x y a b
				const float4 r1 = *reinterpret_cast<const float4 *>(ent + 4); // c o depth unc
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
				const float4 r2 = *reinterpret_cast<const float4 *>(ent + 8); // hx hy r g
				const float cb = ent[12];
				if (C > 0) acc[0] += r2.z * w;
				if (C > 1) acc[1 % C] += r2.w * w;
				if (C > 2) acc[2 % C] += cb * w;
				D += r1.z * w;
				UNC += r1.w * w;
				T = test_T;
				last_ring = feed.done + e + 1u;
				blended |= ebit;
			}
			if (packed) clear_unblended(feed, point_list + range.x, warp, lane, m_cur, blended);
			if (last_ring > feed.done) last_contributor = feed.q_pos[(last_ring - 1u) & (kRing - 1)] + 1u; // resolve before the ring moves on
			feed.done += m_cur;
			__syncwarp(); // the stage buffer and the ring slots of this chunk may be reused
			m_cur = m_next;
			if (__all_sync(0xffffffffu, done)) break; // this warp's 32 pixels are saturated
		}
		feed.drain(); // nothing may be in flight into shared memory when the warp retires
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

// ---- C = 32: weights on the FP32 pipe, accumulation on the tensor pipe -------------------------------------------------------
// Per 16-entry chunk: phase 1 (lane = pixel, entries in depth order) runs the recurrence and leaves w[p][e] = alpha T (0 where
// the pixel does not blend the entry) in shared memory, four entries per STS.128; phase 2 adds W F to the accumulators:
//   A = W   rows p = 16 mt + g (+8), k = entry 8 ks + t (+4)         (LDS.32 from the weight tile, conflict-free at stride 20)
//   B = F   k = entry 8 ks + t (+4), column n = g of tile nt = channel 8 nt + g   (LDS.32 from the staged feature rows)
//   C       lane (g, t) holds pixels 16 mt + g (+8), channels 8 nt + 2 t, + 1
// Entries past the end of a partial chunk carry zero weights, and their (stale) feature operands are zeroed as well.
constexpr int kWStride = 20; // floats per pixel row of the weight tile
struct Fwd32Smem {
	using TR = BlendTraits<32>;
	static constexpr int kWOff = TR::kWarpBytes;
	static constexpr int kWarpBytes = kWOff + 32 * kWStride * 4;
};
#ifndef GSR_FWD32_MINCTAS
#define GSR_FWD32_MINCTAS (20 / kWarpsPerCta)
#endif

__global__ void __launch_bounds__(32 * kWarpsPerCta, GSR_FWD32_MINCTAS) blend_forward_c32_kernel(
    const uint2 *__restrict__ ranges, uint32_t *point_list, const uint32_t *__restrict__ header, int W, int H, int tiles_x,
    const float *__restrict__ rec, const float *__restrict__ features, const float *__restrict__ bg,
    float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
    float *__restrict__ out_color, float *__restrict__ out_depth, float *__restrict__ out_unc)
{
	constexpr int C = 32;
	using TR = BlendTraits<C>;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int tid = threadIdx.x, lwarp = tid >> 5, lane = tid & 31;
	const int gq = lane >> 2, t = lane & 3; // fragment coordinates
	const int tile = blockIdx.x / kCtasPerTile;
	const int warp = (blockIdx.x % kCtasPerTile) * kWarpsPerCta + lwarp; // this warp's 8x4 pixel block within the tile
	int bx, by;
	warp_block_origin(warp, bx, by);
	const int x0 = (tile % tiles_x) * GSR_BLOCK_X + bx, y0 = (tile / tiles_x) * GSR_BLOCK_Y + by;
	const int px = x0 + (lane & 7), py = y0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const float pixf_x = (float)px, pixf_y = (float)py;

	const uint2 range = ranges[tile];
	const int packed = (int)__ldg(header + kHdrPacked); // format of the list entries, recorded by the instance emission
	float T = 1.0f;
	uint32_t last_contributor = 0; // 1-based list position of the last blended Gaussian
	uint32_t last_ring = 0;       // ... as 1 + ring index while its chunk is being blended
	float O[2][4][4];             // accumulators, C-fragment layout
#pragma unroll
	for (int mt = 0; mt < 2; mt++)
#pragma unroll
		for (int nt = 0; nt < 4; nt++)
#pragma unroll
			for (int i = 0; i < 4; i++) O[mt][nt][i] = 0.f;
	float D = 0.f, UNC = 0.f;
	bool done = !inside;

	if (!__all_sync(0xffffffffu, done)) {
		unsigned char *warp_smem = smem_raw + (size_t)lwarp * Fwd32Smem::kWarpBytes;
		float *s_w = reinterpret_cast<float *>(warp_smem + Fwd32Smem::kWOff); // [32 pixels][kWStride]
		WarpFeed<C, false> feed;
		feed.init(warp_smem, point_list + range.x, (int)(range.y - range.x), rec, features, warp, lane, packed != 0);
		feed.fill();
		int m_cur = feed.issue(0);
		for (int chunk = 0; m_cur > 0; chunk++) {
			feed.fill();
			const int m_next = feed.issue((chunk + 1) & 1);
			feed.wait();
			__syncwarp(); // every lane's copies of this chunk have landed
			const float *ent0 = feed.stage + (chunk & 1) * TR::kStageFloats;

			// ---- phase 1: the recurrence; weights of four entries per store ----
			uint32_t blended = 0; // bit e: this pixel blended entry e of the chunk
#pragma unroll
			for (int e0 = 0; e0 < kChunk; e0 += 4) {
				float w4[4] = {0.f, 0.f, 0.f, 0.f};
				if (e0 < m_cur) {
#pragma unroll
					for (int b = 0; b < 4; b++) {
						const int e = e0 + b;
						const float *ent = ent0 + e * TR::kEntryFloats; // (rows past m_cur hold stale but readable shared memory)
						const float4 r0 = *reinterpret_cast<const float4 *>(ent);     // x y a b
						const float4 r1 = *reinterpret_cast<const float4 *>(ent + 4); // c o depth unc
						const float dx = r0.x - pixf_x, dy = r0.y - pixf_y;
						const float power = gaussian_power(r0.z, r0.w, r1.x, dx, dy);
						const float alpha = min(0.99f, __fmul_rn(r1.y, expf(power)));
						const float test_T = __fmul_rn(T, __fsub_rn(1.f, alpha));
						const bool cand = (e < m_cur) && !done && !(power > 0.0f) && !(alpha < kAlphaMin);
						const bool ok = cand && !(test_T < 0.0001f);
						done = done || (cand && !ok);
						const float w = ok ? alpha * T : 0.f;
						if (ok) { // (predicated, not multiplied by a zero weight: a stale row may hold anything)
							D += r1.z * w;
							UNC += r1.w * w;
							T = test_T;
							last_ring = feed.done + e + 1u;
							blended |= 1u << e;
						}
						w4[b] = w;
					}
				}
				*reinterpret_cast<float4 *>(s_w + lane * kWStride + e0) = make_float4(w4[0], w4[1], w4[2], w4[3]);
			}
			__syncwarp(); // the weight tile is visible to every lane

			// ---- phase 2: out += W F on the tensor pipe ----
			if (__any_sync(0xffffffffu, blended != 0)) {
#pragma unroll
				for (int ks = 0; ks < 2; ks++) {
					const int ea = 8 * ks + t, eb = ea + 4; // this lane's two entries of the k-step
					uint32_t ah[2][4], al[2][4];
#pragma unroll
					for (int mt = 0; mt < 2; mt++) {
						const float *wr = s_w + (16 * mt + gq) * kWStride;
						split_tf32(wr[ea], ah[mt][0], al[mt][0]);
						split_tf32(wr[8 * kWStride + ea], ah[mt][1], al[mt][1]);
						split_tf32(wr[eb], ah[mt][2], al[mt][2]);
						split_tf32(wr[8 * kWStride + eb], ah[mt][3], al[mt][3]);
					}
					const float *fa = ent0 + ea * TR::kEntryFloats + TR::kRecParts * 4 + gq, *fb = fa + 4 * TR::kEntryFloats;
#pragma unroll
					for (int nt = 0; nt < 4; nt++) {
						const float f0 = ea < m_cur ? fa[8 * nt] : 0.f, f1 = eb < m_cur ? fb[8 * nt] : 0.f;
						uint32_t b0h, b0l, b1h, b1l;
						split_tf32(f0, b0h, b0l);
						split_tf32(f1, b1h, b1l);
#pragma unroll
						for (int mt = 0; mt < 2; mt++) mma3_tf32(O[mt][nt], ah[mt], al[mt], b0h, b1h, b0l, b1l);
					}
				}
			}
			if (packed) clear_unblended(feed, point_list + range.x, warp, lane, m_cur, blended);
			if (last_ring > feed.done) last_contributor = feed.q_pos[(last_ring - 1u) & (kRing - 1)] + 1u; // resolve before the ring moves on
			feed.done += m_cur;
			__syncwarp(); // the stage buffer, the ring slots and the weight tile of this chunk may be reused
			m_cur = m_next;
			if (__all_sync(0xffffffffu, done)) break; // this warp's 32 pixels are saturated
		}
		feed.drain(); // nothing may be in flight into shared memory when the warp retires
	}

	const size_t plane = (size_t)H * W;
	if (inside) {
		const size_t pix_id = (size_t)W * py + px;
		final_T[pix_id] = T;
		n_contrib[pix_id] = last_contributor;
		out_depth[pix_id] = D;
		out_unc[pix_id] = UNC;
	}
	// colour planes from the fragments: pixels p = 16 mt + g (+8) of the block, channels 8 nt + 2 t, + 1
#pragma unroll
	for (int mt = 0; mt < 2; mt++)
#pragma unroll
		for (int half = 0; half < 2; half++) {
			const int p = 16 * mt + gq + 8 * half;
			const float Tp = __shfl_sync(0xffffffffu, T, p);
			const int qx = x0 + (p & 7), qy = y0 + (p >> 3);
			if (qx < W && qy < H) {
				const size_t pix = (size_t)W * qy + qx;
#pragma unroll
				for (int nt = 0; nt < 4; nt++) {
					const int ch = 8 * nt + 2 * t;
					out_color[ch * plane + pix] = O[mt][nt][2 * half] + Tp * __ldg(bg + ch);
					out_color[(ch + 1) * plane + pix] = O[mt][nt][2 * half + 1] + Tp * __ldg(bg + ch + 1);
				}
			}
		}
}

cudaError_t launch_blend_forward(int C, int W, int H, const uint2 *ranges, const uint32_t *header, uint32_t *point_list, const float *rec,
                                 const float *features, const float *bg, float *final_T, uint32_t *n_contrib, float *out_color,
                                 float *out_depth, float *out_unc, cudaStream_t stream)
{
	const int tiles_x = (W + GSR_BLOCK_X - 1) / GSR_BLOCK_X, tiles_y = (H + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y;
	const int tiles = tiles_x * tiles_y;
	if (tiles <= 0) return cudaSuccess;
	cudaError_t e;
	switch (C) {
	case 3: {
		constexpr int smem = kWarpsPerCta * BlendTraits<3>::kWarpBytes;
		// (the attribute is per device and idempotent: set on every launch rather than cached in a per-process flag)
		if ((e = cudaFuncSetAttribute(blend_forward_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
		blend_forward_kernel<3><<<tiles * kCtasPerTile, 32 * kWarpsPerCta, smem, stream>>>(ranges, point_list, header, W, H, tiles_x, rec, bg, final_T, n_contrib,
		                                                                                out_color, out_depth, out_unc);
		break;
	}
	case 32: {
		constexpr int smem = kWarpsPerCta * Fwd32Smem::kWarpBytes;
		if ((e = cudaFuncSetAttribute(blend_forward_c32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
		blend_forward_c32_kernel<<<tiles * kCtasPerTile, 32 * kWarpsPerCta, smem, stream>>>(ranges, point_list, header, W, H, tiles_x, rec, features, bg, final_T,
		                                                                                 n_contrib, out_color, out_depth, out_unc);
		break;
	}
	default: return cudaErrorInvalidValue;
	}
	count_launch();
	return cudaGetLastError();
}

} // namespace gsr
