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
//   * one CTA per tile, but its 8 warps never synchronise: each warp owns an 8x4 pixel block and feeds itself
//     (gsr_blend.cuh: WarpFeed) — it scans the tile's list, keeps only the instances whose alpha >= 1/255
//     bounding box touches its block (an 8-bit mask precomputed per instance by the binning stage), gathers their
//     32-B projected records and feature rows into its private double-buffered shared-memory stage with 16-B
//     cp.async one chunk ahead, and runs the reference's per-pixel recurrence over the landed chunk.  The skipped
//     pairs are exactly pairs the reference `continue`s over, so results and n_contrib are unchanged;
//   * a warp stops as soon as its 32 pixels are saturated (the reference stops per tile, CR/forward.cu:496-498);
//   * features are read from shared memory (the reference re-reads colour and depth from global memory for
//     every contributing pixel, forward.cu:545-546).
// History of the staging engine (per-Gaussian bulk copies -> block-wide LDGSTS slabs -> warp-private feeds) with the
// measurements that drove it: DESIGN.md section 4, profiles/r1_staging_ab.md, profiles/r1_feed_ab.md.
#include "gsr_blend.cuh"
#include "gsr_internal.cuh"
#include "gsr_tf32.cuh"

namespace gsr {

constexpr int kWarpsPerCta = GSR_FWD_WARPS_PER_CTA;
constexpr int kCtasPerTile = kWarpsPerTile / kWarpsPerCta;

// Resident CTAs per SM the register allocation must allow.  Measured (profiles/r1_occupancy_ab.md): the forward
// kernel is latency bound, 4 CTAs/SM (64 registers) beats 3 CTAs/SM (80 registers) by 10 %.
// the per-entry bit of the `blended` mask as a shifted loop variable instead of 1u << e: -1 % (0.880 vs 0.892 ms)
#ifndef GSR_FWD_EBIT
#define GSR_FWD_EBIT 1
#endif
// feature rows by per-entry TMA bulk copies instead of 16-B cp.async (C = 32): A/B in profiles/r1_feed_ab.md section 6
#ifndef GSR_FWD_FEED_BULK
#define GSR_FWD_FEED_BULK 0
#endif
#ifndef GSR_FWD_MINBLOCKS
#define GSR_FWD_MINBLOCKS 4
#endif
// keep the per-entry loop rolled (ptxas unrolls it by two, which costs registers): A/B switch
#ifndef GSR_FWD_UNROLL1
#define GSR_FWD_UNROLL1 0
#endif
// resident CTAs per SM asked of ptxas (register budget = 65536 / (threads x CTAs)); -DGSR_FWD_MINCTAS=n overrides for A/B
#ifndef GSR_FWD_MINCTAS
#define GSR_FWD_MINCTAS (GSR_FWD_MINBLOCKS * kCtasPerTile)
#endif
template <int C>
__global__ void __launch_bounds__(32 * kWarpsPerCta, GSR_FWD_MINCTAS) blend_forward_kernel(
    const uint2 *__restrict__ ranges, uint32_t *point_list, int packed, int W, int H, int tiles_x,
    const float *__restrict__ rec, const float *__restrict__ features, const float *__restrict__ bg,
    float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
    float *__restrict__ out_color, float *__restrict__ out_depth, float *__restrict__ out_unc)
{
	using TR = BlendTraits<C>;
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
	float T = 1.0f;
	uint32_t last_contributor = 0; // 1-based list position of the last blended Gaussian
	uint32_t last_ring = 0;       // ... as 1 + ring index while its chunk is being blended
	float acc[C];
#pragma unroll
	for (int ch = 0; ch < C; ch++) acc[ch] = 0.f;
	float D = 0.f, UNC = 0.f;
	// packed accumulators (GSR_FFMA2, C > 3): channel pairs (2i, 2i+1) and (depth, uncertainty)
	constexpr bool kPacked = (GSR_FFMA2 != 0) && !TR::kFeatInRec;
	constexpr int kPairs = kPacked ? C / 2 : 1;
	uint64_t acc2[kPairs], du2 = 0ull;
#pragma unroll
	for (int i = 0; i < kPairs; i++) acc2[i] = 0ull;
	bool done = !inside;

	if (!__all_sync(0xffffffffu, done)) {
		using Feed = WarpFeed<C, false, (GSR_FWD_FEED_BULK != 0) && (C > 3)>;
		Feed feed;
		feed.init(smem_raw + (size_t)lwarp * (TR::kWarpBytes + Feed::kExtraBytes), point_list + range.x, (int)(range.y - range.x), rec, features, warp, lane, packed != 0);
		feed.fill();
		int m_cur = feed.issue(0);
		int chunk = 0;
		for (; m_cur > 0; chunk++) {
			feed.fill();
			const int m_next = feed.issue((chunk + 1) & 1);
			feed.wait(chunk, m_cur);
			__syncwarp(); // every lane's copies of this chunk have landed
			const float *ent = feed.stage + (chunk & 1) * TR::kStageFloats;
			uint32_t blended = 0; // bit e: this pixel blended entry e of the chunk
#if GSR_FWD_EBIT
			uint32_t ebit = 1u;
#if GSR_FWD_UNROLL1
#pragma unroll 1
#endif
			for (int e = 0; e < m_cur; e++, ent += TR::kEntryFloats, ebit <<= 1) {
#else
			for (int e = 0; e < m_cur; e++, ent += TR::kEntryFloats) {
#endif
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
				if (TR::kFeatInRec) {
					const float4 r2 = *reinterpret_cast<const float4 *>(ent + 8); // hx hy r g
					const float cb = ent[12];
					if (C > 0) acc[0] += r2.z * w;
					if (C > 1) acc[1 % C] += r2.w * w;
					if (C > 2) acc[2 % C] += cb * w;
				} else if (kPacked) {
					const uint64_t ww = pack2(w, w);
					const float4 *f4 = reinterpret_cast<const float4 *>(ent + TR::kRecParts * 4);
#pragma unroll
					for (int q = 0; q < C / 4; q++) {
						const float4 f = f4[q];
						acc2[(2 * q) % kPairs] = fma2(pack2(f.x, f.y), ww, acc2[(2 * q) % kPairs]);
						acc2[(2 * q + 1) % kPairs] = fma2(pack2(f.z, f.w), ww, acc2[(2 * q + 1) % kPairs]);
					}
					du2 = fma2(pack2(r1.z, r1.w), ww, du2);
				} else {
					const float4 *f4 = reinterpret_cast<const float4 *>(ent + TR::kRecParts * 4);
#pragma unroll
					for (int q = 0; q < C / 4; q++) {
						const float4 f = f4[q];
						acc[4 * q + 0] += f.x * w;
						acc[4 * q + 1] += f.y * w;
						acc[4 * q + 2] += f.z * w;
						acc[4 * q + 3] += f.w * w;
					}
				}
				if (!kPacked) {
					D += r1.z * w;
					UNC += r1.w * w;
				}
				T = test_T;
				last_ring = feed.done + e + 1u;
#if GSR_FWD_EBIT
				blended |= ebit;
#else
				blended |= 1u << e;
#endif
			}
			// Instances this warp examined but none of its 32 pixels blended: clear the warp's bit in the instance's mask, so that
			// the backward pass (which scans the same list with the same masks) visits exactly the contributing (warp, instance) pairs
			// — 30 % fewer than the bounding-box candidates.  Only this warp tests this bit, and only before clearing it.
			if (packed) {
				const uint32_t any_blended = __reduce_or_sync(0xffffffffu, blended);
				if (lane < m_cur && !((any_blended >> lane) & 1u))
					atomicAnd(point_list + range.x + feed.q_pos[(feed.done + lane) & (kRing - 1)], ~(1u << (24 + warp)));
			}
			if (last_ring > feed.done) last_contributor = feed.q_pos[(last_ring - 1u) & (kRing - 1)] + 1u; // resolve before the ring moves on
			feed.done += m_cur;
			__syncwarp(); // the stage buffer and the ring slots of this chunk may be reused
			m_cur = m_next;
			if (__all_sync(0xffffffffu, done)) { chunk++; break; } // this warp's 32 pixels are saturated; chunk `chunk` may be in flight
		}
		feed.drain(chunk, m_cur); // nothing may be in flight into shared memory when the warp retires
	}

	if (kPacked) {
#pragma unroll
		for (int i = 0; i < (kPacked ? C / 2 : 0); i++) unpack2(acc2[i], acc[(2 * i) % C], acc[(2 * i + 1) % C]);
		unpack2(du2, D, UNC);
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

// ---- C = 32 variant with the colour accumulation on the tensor pipe (-DGSR_FWD_MMA=1) ------------------------------------
// out[p][ch] += sum_j w[p][j] f[j][ch] over a sub-chunk of 8 staged Gaussians is a [32 x 8] x [8 x 32] product per warp: the scalar
// kernel spends 32 FFMA + 8 LDS.128 per (pixel lane, contributing Gaussian) on it, 36 % of its instructions, and it is issue
// bound (90 % issue-slot utilisation, profiles/r1_blend_v5_summary.md).  Here the per-pixel recurrence only produces
// w = alpha * T (0 for non-contributing pairs) into an 8 x 32 shared tile, and 24 mma.sync.m16n8k8 TF32 (3xTF32 split) per
// sub-chunk add the product to accumulators held in C-fragment layout:
//   A = w   rows p = 16 mt + g (+8), k = Gaussian t (+4) of the sub-chunk
//   B = f   k = Gaussian t (+4), col n = g of tile nt <-> channel 4 g + nt (one LDS.128 per Gaussian covers the four tiles)
//   C       lane (g, t) holds pixels 16 mt + g (+8), channels 8 t + nt and 8 t + 4 + nt
#ifndef GSR_FWD_MMA
#define GSR_FWD_MMA 0
#endif
constexpr int kSub = 8;                                   // Gaussians per MMA k-step
constexpr int kWStride = 40;                              // floats per row of the 8 x 32 weight tile (bank-conflict-free A loads)
constexpr int kFwdMmaWarpBytes = BlendTraits<32>::kWarpBytes + kSub * kWStride * 4;
#ifndef GSR_FWD_MMA_MINWARPS
#define GSR_FWD_MMA_MINWARPS 26
#endif

__global__ void __launch_bounds__(32 * kWarpsPerCta, GSR_FWD_MMA_MINWARPS / kWarpsPerCta) blend_forward_mma_kernel(
    const uint2 *__restrict__ ranges, uint32_t *point_list, int packed, int W, int H, int tiles_x,
    const float *__restrict__ rec, const float *__restrict__ features, const float *__restrict__ bg,
    float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
    float *__restrict__ out_color, float *__restrict__ out_depth, float *__restrict__ out_unc)
{
	constexpr int C = 32;
	using TR = BlendTraits<C>;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int tid = threadIdx.x, lwarp = tid >> 5, lane = tid & 31;
	const int tile = blockIdx.x / kCtasPerTile;
	const int warp = (blockIdx.x % kCtasPerTile) * kWarpsPerCta + lwarp;
	const int tile_x0 = (tile % tiles_x) * GSR_BLOCK_X, tile_y0 = (tile / tiles_x) * GSR_BLOCK_Y;
	int bx, by;
	warp_block_origin(warp, bx, by);
	const int px = tile_x0 + bx + (lane & 7), py = tile_y0 + by + (lane >> 3);
	const bool inside = px < W && py < H;
	const float pixf_x = (float)px, pixf_y = (float)py;
	const int fg = lane >> 2, ft = lane & 3;

	const uint2 range = ranges[tile];
	float T = 1.0f;
	uint32_t last_contributor = 0, last_ring = 0;
	float acc[2][4][4]; // [mt][nt][c0..c3]
#pragma unroll
	for (int mt = 0; mt < 2; mt++)
#pragma unroll
		for (int nt = 0; nt < 4; nt++)
#pragma unroll
			for (int i = 0; i < 4; i++) acc[mt][nt][i] = 0.f;
	float D = 0.f, UNC = 0.f;
	bool done = !inside;

	if (!__all_sync(0xffffffffu, done)) {
		unsigned char *wsm = smem_raw + (size_t)lwarp * kFwdMmaWarpBytes;
		float *s_w = reinterpret_cast<float *>(wsm + TR::kWarpBytes); // [8][40]: alpha * T of the sub-chunk, row = Gaussian, col = pixel
		WarpFeed<C, false> feed;
		feed.init(wsm, point_list + range.x, (int)(range.y - range.x), rec, features, warp, lane, packed != 0);
		feed.fill();
		int m_cur = feed.issue(0);
		for (int chunk = 0; m_cur > 0; chunk++) {
			feed.fill();
			const int m_next = feed.issue((chunk + 1) & 1);
			cp_async_wait_but_one();
			__syncwarp();
			const float *ent0 = feed.stage + (chunk & 1) * TR::kStageFloats;
			uint32_t blended = 0;
			for (int sub = 0; sub < m_cur; sub += kSub) {
				const float *ent = ent0 + sub * TR::kEntryFloats;
				bool any_w = false;
#pragma unroll
				for (int e = 0; e < kSub; e++, ent += TR::kEntryFloats) {
					float w = 0.f;
					if (sub + e < m_cur) {
						const float4 r0 = *reinterpret_cast<const float4 *>(ent);     // x y a b
						const float4 r1 = *reinterpret_cast<const float4 *>(ent + 4); // c o depth unc
						const float2 d = {r0.x - pixf_x, r0.y - pixf_y};
						const float power = gaussian_power(r0.z, r0.w, r1.x, d.x, d.y);
						if (!(done || power > 0.0f)) {
							const float alpha = min(0.99f, __fmul_rn(r1.y, expf(power)));
							if (!(alpha < kAlphaMin)) {
								const float test_T = __fmul_rn(T, __fsub_rn(1.f, alpha));
								if (test_T < 0.0001f) {
									done = true;
								} else {
									w = alpha * T;
									D += r1.z * w;
									UNC += r1.w * w;
									T = test_T;
									last_ring = feed.done + sub + e + 1u;
									blended |= 1u << (sub + e);
								}
							}
						}
					}
					s_w[e * kWStride + lane] = w;
					any_w |= (w != 0.f);
				}
				__syncwarp();
				if (__any_sync(0xffffffffu, any_w)) {
					// B fragments: Gaussians t and t+4 of the sub-chunk, channels 4g..4g+3 (the four n-tiles)
					const float *fb = ent0 + (sub + ft) * TR::kEntryFloats + TR::kRecParts * 4 + 4 * fg;
					const float4 f0 = *reinterpret_cast<const float4 *>(fb), f1 = *reinterpret_cast<const float4 *>(fb + 4 * TR::kEntryFloats);
					const float b0v[4] = {f0.x, f0.y, f0.z, f0.w}, b1v[4] = {f1.x, f1.y, f1.z, f1.w};
					uint32_t b0h[4], b0l[4], b1h[4], b1l[4];
#pragma unroll
					for (int nt = 0; nt < 4; nt++) {
						tf32_split(b0v[nt], b0h[nt], b0l[nt]);
						tf32_split(b1v[nt], b1h[nt], b1l[nt]);
					}
#pragma unroll
					for (int mt = 0; mt < 2; mt++) {
						uint32_t ah[4], al[4];
						tf32_split(s_w[ft * kWStride + 16 * mt + fg], ah[0], al[0]);
						tf32_split(s_w[ft * kWStride + 16 * mt + fg + 8], ah[1], al[1]);
						tf32_split(s_w[(ft + 4) * kWStride + 16 * mt + fg], ah[2], al[2]);
						tf32_split(s_w[(ft + 4) * kWStride + 16 * mt + fg + 8], ah[3], al[3]);
#pragma unroll
						for (int nt = 0; nt < 4; nt++) {
							mma_tf32(acc[mt][nt], al[0], al[1], al[2], al[3], b0h[nt], b1h[nt]);
							mma_tf32(acc[mt][nt], ah[0], ah[1], ah[2], ah[3], b0l[nt], b1l[nt]);
							mma_tf32(acc[mt][nt], ah[0], ah[1], ah[2], ah[3], b0h[nt], b1h[nt]);
						}
					}
				}
				__syncwarp(); // the weight tile may be overwritten
			}
			if (packed) {
				const uint32_t any_blended = __reduce_or_sync(0xffffffffu, blended);
				if (lane < m_cur && !((any_blended >> lane) & 1u))
					atomicAnd(point_list + range.x + feed.q_pos[(feed.done + lane) & (kRing - 1)], ~(1u << (24 + warp)));
			}
			if (last_ring > feed.done) last_contributor = feed.q_pos[(last_ring - 1u) & (kRing - 1)] + 1u;
			feed.done += m_cur;
			__syncwarp();
			m_cur = m_next;
			if (__all_sync(0xffffffffu, done)) break;
		}
		cp_async_wait_all();
	}

	const size_t plane = (size_t)H * W;
	if (inside) {
		const size_t pix_id = (size_t)W * py + px;
		final_T[pix_id] = T;
		n_contrib[pix_id] = last_contributor;
		out_depth[pix_id] = D;
		out_unc[pix_id] = UNC;
	}
	// colour planes from the C fragments: pixels p = 16 mt + g (+8) of the warp's 8x4 block, channels 8t + nt and 8t + 4 + nt
#pragma unroll
	for (int mt = 0; mt < 2; mt++) {
#pragma unroll
		for (int half = 0; half < 2; half++) {
			const int p = 16 * mt + fg + 8 * half;
			const float Tp = __shfl_sync(0xffffffffu, T, p);
			const int qx = tile_x0 + bx + (p & 7), qy = tile_y0 + by + (p >> 3);
			if (qx < W && qy < H) {
				const size_t pix = (size_t)W * qy + qx;
#pragma unroll
				for (int nt = 0; nt < 4; nt++) {
					const int ch0 = 8 * ft + nt, ch1 = ch0 + 4;
					out_color[ch0 * plane + pix] = acc[mt][nt][2 * half] + Tp * __ldg(bg + ch0);
					out_color[ch1 * plane + pix] = acc[mt][nt][2 * half + 1] + Tp * __ldg(bg + ch1);
				}
			}
		}
	}
}

template <int C>
static size_t fwd_smem_bytes()
{
	return (size_t)kWarpsPerCta * (BlendTraits<C>::kWarpBytes + ((GSR_FWD_FEED_BULK != 0 && C > 3) ? 16 : 0));
}

template <int C>
static cudaError_t launch_fwd(int tiles, const uint2 *ranges, uint32_t *point_list, int packed, int W, int H, int tiles_x, const float *rec,
                              const float *features, const float *bg, float *final_T, uint32_t *n_contrib, float *out_color,
                              float *out_depth, float *out_unc, cudaStream_t stream)
{
	using TR = BlendTraits<C>;
	static bool configured = false;
	if (!configured) {
		cudaError_t e = cudaFuncSetAttribute(blend_forward_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem_bytes<C>());
		if (e != cudaSuccess) return e;
		configured = true;
	}
	blend_forward_kernel<C><<<tiles * kCtasPerTile, 32 * kWarpsPerCta, fwd_smem_bytes<C>(), stream>>>(ranges, point_list, packed, W, H, tiles_x, rec, features, bg, final_T, n_contrib,
	                                                               out_color, out_depth, out_unc);
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_blend_forward(int C, int P, int W, int H, const uint2 *ranges, uint32_t *point_list, const float *rec,
                                 const float *features, const float *bg, float *final_T, uint32_t *n_contrib, float *out_color,
                                 float *out_depth, float *out_unc, cudaStream_t stream)
{
	const int tiles_x = (W + GSR_BLOCK_X - 1) / GSR_BLOCK_X, tiles_y = (H + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y;
	const int tiles = tiles_x * tiles_y;
	if (tiles <= 0) return cudaSuccess;
	const int packed = point_list_packed(P) ? 1 : 0;
#if GSR_FWD_MMA
	if (C == 32) {
		const size_t smem = (size_t)kWarpsPerCta * kFwdMmaWarpBytes;
		static bool configured = false;
		if (!configured) {
			cudaError_t e = cudaFuncSetAttribute(blend_forward_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (e != cudaSuccess) return e;
			configured = true;
		}
		blend_forward_mma_kernel<<<tiles * kCtasPerTile, 32 * kWarpsPerCta, smem, stream>>>(ranges, point_list, packed, W, H, tiles_x, rec, features, bg,
		                                                                                  final_T, n_contrib, out_color, out_depth, out_unc);
		count_launch();
		return cudaGetLastError();
	}
#endif
	switch (C) {
	case 3: return launch_fwd<3>(tiles, ranges, point_list, packed, W, H, tiles_x, rec, features, bg, final_T, n_contrib, out_color, out_depth, out_unc, stream);
	case 32: return launch_fwd<32>(tiles, ranges, point_list, packed, W, H, tiles_x, rec, features, bg, final_T, n_contrib, out_color, out_depth, out_unc, stream);
	default: return cudaErrorInvalidValue;
	}
}

} // namespace gsr
