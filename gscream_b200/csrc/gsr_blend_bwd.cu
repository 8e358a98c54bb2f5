// gsr_blend_bwd.cu — backward tile blend: per-pixel -> per-Gaussian gradient scatter.
//
// Replaces renderCUDA<C> backward (CR/backward.cu:409-604 of W-Ted/GScream's
// submodules/diff-gaussian-rasterization).  Same traversal (back to front from the last contributor,
// T recovered by division, straight-through min(0.99,.), background term on colour channels only), but
// restructured for the SM instead of translated:
//
//   * The reference keeps three per-thread arrays of C floats (accum_rec, last_color, dL_dpixel) and
//     issues C+8 global float atomics per contributing (pixel, Gaussian) pair (179 registers and
//     40 atomics per pair at C = 32).  Here the per-channel recurrence
//         accum_rec[ch] <- last_alpha*last_color[ch] + (1-last_alpha)*accum_rec[ch]
//         dL_dalpha     += (c[ch] - accum_rec[ch]) * dL_dpixel[ch]
//     is collapsed, by linearity in ch, into ONE scalar recurrence on X = sum_ch accum_rec[ch]*g[ch]:
//         X <- last_alpha*last_dot + (1-last_alpha)*X,   dL_dalpha = (dot - X) * T,   dot = f_j . g_p
//     so a pixel needs only its gradient row g_p (C+2 registers) and three scalars.
//   * Per-Gaussian sums over the warp's 32 pixels are formed on chip before touching global memory:
//     the 8 geometric/scalar terms by a transpose-reduce butterfly (9 shuffles instead of 40), and, for
//     C = 32, the C colour terms by switching roles — lane l owns channel l and holds the gradient
//     COLUMN of its warp's 32 pixels in registers, the per-pixel weights alpha*T go through 128 B of
//     shared memory, and the warp issues a single coalesced 128-B red.global.add per Gaussian.
//     That is one RED instruction per (warp, Gaussian) where the reference issues 32 x C scalar atomics.
//   * Same warp-private feed as the forward kernel (gsr_blend.cuh: per-warp list scan by the instance masks, private
//     double-buffered cp.async gather, no block barrier), run back to front and started at the warp's own deepest last
//     contributor — nothing behind it can receive gradient.
//
// Scalar terms are accumulated into gacc[P][8] = {dmean2D.x, dmean2D.y, dconic.x, dconic.y, dconic.w,
// dopacity, ddepth, duncertainty}; colours into dL_dcolors[P][C].  Both must be zero (or hold the
// running sum) on entry.
#include "gsr_blend.cuh"
#include "gsr_internal.cuh"

#ifndef GSR_BWD_RCP
#define GSR_BWD_RCP 1
#endif
#ifndef GSR_BWD_ACC4
#define GSR_BWD_ACC4 0
#endif
// C = 32: evaluate the two 32x32 products per (warp, chunk) on the tensor pipe (gsr_blend_bwd_mma.cu) instead of FFMA
#ifndef GSR_BWD_MMA
#define GSR_BWD_MMA 0
#endif
// C = 32: keep the lane = channel copy of the gradient block (the column each lane needs for the colour sums) in shared
// memory instead of 32 registers per lane: ~96 registers -> 20 warps per SM instead of 16, for 8 more LDS.128 per pair
#ifndef GSR_BWD_GCOL_SMEM
#define GSR_BWD_GCOL_SMEM 0
#endif
// load the next entry's record (two LDS.128) one iteration ahead of its use: measured slower (2.32 vs 2.18 ms, 8 B of spill)
#ifndef GSR_BWD_PREFETCH
#define GSR_BWD_PREFETCH 0
#endif
// two partial sums in the lane = channel colour sums: -1 % (2.153 vs 2.178 ms); four (GSR_BWD_ACC4) are slower
#ifndef GSR_BWD_PB2
#define GSR_BWD_PB2 1
#endif


namespace gsr {

constexpr int kWarpsPerCta = GSR_BWD_WARPS_PER_CTA;
constexpr int kCtasPerTile = kWarpsPerTile / kWarpsPerCta;

// Transpose-reduce: every lane contributes N values; afterwards v[0] on lane l is the warp-wide total
// of value index vidx<N>(l).  N/2 + N/4 + ... + 1 exchanges, then plain xor-adds for the remaining strides.
template <int N>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[N], int lane)
{
	int s = 16;
#pragma unroll
	for (int n = N / 2; n >= 1; n >>= 1, s >>= 1) {
		const bool upper = (lane & s) != 0;
#pragma unroll
		for (int i = 0; i < n; i++) {
			const float send = upper ? v[i] : v[i + n];
			const float keep = upper ? v[i + n] : v[i];
			v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
		}
	}
#pragma unroll
	for (; s >= 1; s >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], s);
}
template <int N>
__device__ __forceinline__ int vidx(int lane)
{
	int idx = 0, s = 16;
#pragma unroll
	for (int n = N / 2; n >= 1; n >>= 1, s >>= 1)
		if (lane & s) idx += n;
	return idx;
}
template <int N>
__device__ __forceinline__ bool vowner(int lane)
{
	// the lowest lane among those holding the same total
	int rest = 32 / N - 1; // mask of the low bits not consumed by the exchange steps
	return (lane & rest) == 0;
}

// C = 32 keeps a 34-float gradient row and a 32-float gradient column per lane: 124 registers, 2 CTAs/SM (forcing 3
// spills 108 B and is 18 % slower, profiles/r1_occupancy_ab.md).  Small C fits 3.
// feature rows by per-entry TMA bulk copies instead of 16-B cp.async (C = 32): A/B in profiles/r1_feed_ab.md section 6
#ifndef GSR_BWD_FEED_BULK
#define GSR_BWD_FEED_BULK 0
#endif
#ifndef GSR_BWD_MINBLOCKS
#define GSR_BWD_MINBLOCKS(C) ((C) <= 8 ? 3 : 2)
#endif
#if GSR_BWD_GCOL_SMEM
#ifndef GSR_BWD_GCOL_WARPS
#define GSR_BWD_GCOL_WARPS 20
#endif
#define GSR_BWD_MINCTAS(C) ((C) == 32 ? (GSR_BWD_GCOL_WARPS / kWarpsPerCta) : GSR_BWD_MINBLOCKS(C) * kCtasPerTile)
#elif defined(GSR_BWD_MINCTAS32)
#define GSR_BWD_MINCTAS(C) ((C) == 32 ? GSR_BWD_MINCTAS32 : GSR_BWD_MINBLOCKS(C) * kCtasPerTile)
#else
#define GSR_BWD_MINCTAS(C) (GSR_BWD_MINBLOCKS(C) * kCtasPerTile)
#endif
template <int C>
__global__ void __launch_bounds__(32 * kWarpsPerCta, GSR_BWD_MINCTAS(C)) blend_backward_kernel(
    const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, int packed, int W, int H, int tiles_x,
    const float *__restrict__ rec, const float *__restrict__ features, const float *__restrict__ bg,
    const float *__restrict__ final_Ts, const uint32_t *__restrict__ n_contrib,
    const float *__restrict__ dL_dpixels, const float *__restrict__ dL_dpixel_depths, const float *__restrict__ dL_dpixel_uncs,
    float *__restrict__ gacc, float *__restrict__ dL_dcolors)
{
	using TR = BlendTraits<C>;
	constexpr bool kLaneChannel = (C == 32); // colour sums by role switch; otherwise through the butterfly
	constexpr bool kGcolSmem = kLaneChannel && (GSR_BWD_GCOL_SMEM != 0);
	constexpr int kGcolStride = 36;
	constexpr int NV = kLaneChannel ? 8 : 16;
	static_assert(kLaneChannel || C <= 8, "butterfly path carries at most 8 colour channels");

	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ __align__(16) float s_w[kWarpsPerCta][32];

	const int tid = threadIdx.x, lwarp = tid >> 5, lane = tid & 31;
	const int tile = blockIdx.x / kCtasPerTile;
	const int warp = (blockIdx.x % kCtasPerTile) * kWarpsPerCta + lwarp; // this warp's 8x4 pixel block within the tile
	const int tile_x0 = (tile % tiles_x) * GSR_BLOCK_X, tile_y0 = (tile / tiles_x) * GSR_BLOCK_Y;
	int bx, by;
	warp_block_origin(warp, bx, by);
	const int px = tile_x0 + bx + (lane & 7), py = tile_y0 + by + (lane >> 3);
	const bool inside = px < W && py < H;
	const float pixf_x = (float)px, pixf_y = (float)py;
	const size_t plane = (size_t)H * W;
	const size_t pix_id = (size_t)W * py + px;

	const uint2 range = ranges[tile];
	const float T_final = inside ? final_Ts[pix_id] : 0.f;
	const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

	// deepest last contributor of the warp: nothing behind it can receive gradient from these 32 pixels
	int warp_last = last_contributor;
#pragma unroll
	for (int s = 16; s >= 1; s >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, s));
	warp_last = min(warp_last, (int)(range.y - range.x));
	if (warp_last == 0) return; // warps are independent: no barrier follows

	// this pixel's upstream gradient row (colour channels, depth, uncertainty)
	float g[C];
	float gd = 0.f, gu = 0.f, bg_dot = 0.f;
#pragma unroll
	for (int ch = 0; ch < C; ch++) {
		g[ch] = inside ? dL_dpixels[ch * plane + pix_id] : 0.f;
		bg_dot += bg[ch] * g[ch];
	}
	if (inside) {
		gd = dL_dpixel_depths[pix_id];
		gu = dL_dpixel_uncs[pix_id];
	}
	// role switch (C == 32): lane l also holds channel l's gradient for the warp's 32 pixels
	float gcol[(kLaneChannel && !kGcolSmem) ? 32 : 1];
	float *s_gc = reinterpret_cast<float *>(smem_raw + (size_t)kWarpsPerCta * (TR::kWarpBytes + ((GSR_BWD_FEED_BULK != 0 && C > 3) ? 16 : 0))) + lwarp * 32 * kGcolStride; // [ch][36]
	if (kGcolSmem) {
#pragma unroll
		for (int ch = 0; ch < C; ch++) s_gc[ch * kGcolStride + lane] = g[ch];
		__syncwarp();
	} else if (kLaneChannel) {
		const float *src = dL_dpixels + (size_t)lane * plane;
		const int x0 = tile_x0 + bx, y0 = tile_y0 + by;
		const bool vec_ok = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(dL_dpixels) & 15) == 0) && (x0 + 8 <= W);
#pragma unroll
		for (int rr = 0; rr < 4; rr++) {
			const int y = y0 + rr;
			if (vec_ok && y < H) {
				const float4 *p4 = reinterpret_cast<const float4 *>(src + (size_t)W * y + x0);
				const float4 a4 = __ldg(p4), b4 = __ldg(p4 + 1);
				gcol[rr * 8 + 0] = a4.x; gcol[rr * 8 + 1] = a4.y; gcol[rr * 8 + 2] = a4.z; gcol[rr * 8 + 3] = a4.w;
				gcol[rr * 8 + 4] = b4.x; gcol[rr * 8 + 5] = b4.y; gcol[rr * 8 + 6] = b4.z; gcol[rr * 8 + 7] = b4.w;
			} else {
#pragma unroll
				for (int cc = 0; cc < 8; cc++) {
					const int x = x0 + cc;
					gcol[rr * 8 + cc] = (x < W && y < H) ? __ldg(src + (size_t)W * y + x) : 0.f;
				}
			}
		}
	}

	// packed copies (GSR_FFMA2, C == 32): the gradient row and column as register pairs for fma.rn.f32x2
	constexpr bool kPacked = (GSR_FFMA2 != 0) && kLaneChannel && !kGcolSmem;
	constexpr int kPairs = kPacked ? C / 2 : 1;
	uint64_t g2[kPairs], gcol2[kPairs];
	if (kPacked) {
#pragma unroll
		for (int i = 0; i < kPairs; i++) {
			g2[i] = pack2(g[(2 * i) % C], g[(2 * i + 1) % C]);
			gcol2[i] = pack2(gcol[(2 * i) % (kPacked ? 32 : 1)], gcol[(2 * i + 1) % (kPacked ? 32 : 1)]);
		}
	}

	float T = T_final;
	float X = 0.f, last_alpha = 0.f, last_dot = 0.f;
	const float ddelx_dx = 0.5 * W, ddely_dy = 0.5 * H;
	const float neg_Tfinal_bg = -T_final * bg_dot; // background term: (-T_final / (1 - alpha)) * sum_ch bg[ch] g[ch]

	// back to front (CR/backward.cu:500): the feed scans list positions warp_last-1 .. 0
	using Feed = WarpFeed<C, true, (GSR_BWD_FEED_BULK != 0) && (C > 3)>;
	Feed feed;
	feed.init(smem_raw + (size_t)lwarp * (TR::kWarpBytes + Feed::kExtraBytes), point_list + range.x, warp_last, rec, features, warp, lane, packed != 0);
	feed.fill();
	int m_cur = feed.issue(0);
	int chunk = 0;
	for (; m_cur > 0; chunk++) {
		feed.fill();
		const int m_next = feed.issue((chunk + 1) & 1);
		feed.wait(chunk, m_cur);
		__syncwarp(); // every lane's copies of this chunk have landed
		const float *ent = feed.stage + (chunk & 1) * TR::kStageFloats;
#if GSR_BWD_PREFETCH
		float4 r0n = *reinterpret_cast<const float4 *>(ent), r1n = *reinterpret_cast<const float4 *>(ent + 4);
#endif
		for (int e = 0; e < m_cur; e++, ent += TR::kEntryFloats) {
			const uint32_t slot = (feed.done + e) & (kRing - 1);
			const int pos = (int)feed.q_pos[slot]; // 0-based list position
#if GSR_BWD_PREFETCH
			const float4 r0 = r0n, r1 = r1n;     // this entry's record was loaded during the previous iteration
			if (e + 1 < m_cur) {
				r0n = *reinterpret_cast<const float4 *>(ent + TR::kEntryFloats);
				r1n = *reinterpret_cast<const float4 *>(ent + TR::kEntryFloats + 4);
			}
#else
			const float4 r0 = *reinterpret_cast<const float4 *>(ent);     // x y a b
			const float4 r1 = *reinterpret_cast<const float4 *>(ent + 4); // c o depth unc
#endif
			const float2 d = {r0.x - pixf_x, r0.y - pixf_y};
			const float power = gaussian_power(r0.z, r0.w, r1.x, d.x, d.y);
			const bool maybe = (pos < last_contributor) && !(power > 0.0f);
			const float G = expf(power);
			const float alpha = min(0.99f, __fmul_rn(r1.y, G));
			const bool valid = maybe && !(alpha < kAlphaMin);
			if (!__any_sync(0xffffffffu, valid)) continue;

			float v[NV];
#pragma unroll
			for (int i = 0; i < NV; i++) v[i] = 0.f;
			float w = 0.f;
			if (valid) {
				// T <- T / (1 - alpha) (CR/backward.cu:533), as T * rcp(1 - alpha); the reciprocal also serves the
				// background term
#if GSR_BWD_RCP
				const float rinv = __frcp_rn(__fsub_rn(1.f, alpha));
				T = T * rinv;
#else
				const float one_minus = __fsub_rn(1.f, alpha);
				T = __fdiv_rn(T, one_minus);
#endif
				w = alpha * T;
				// dot = f_j . g_p over colour channels, depth and uncertainty
#if GSR_BWD_ACC4
				float d0 = r1.z * gd, d1 = r1.w * gu, d2 = 0.f, d3 = 0.f;
#else
				float d0 = r1.z * gd + r1.w * gu;
				float &d1 = d0, &d2 = d0, &d3 = d0;
#endif
				if (TR::kFeatInRec) {
					const float4 r2 = *reinterpret_cast<const float4 *>(ent + 8);
					const float cb = ent[12];
					if (C > 0) d2 += r2.z * g[0];
					if (C > 1) d3 += r2.w * g[1 % C];
					if (C > 2) d0 += cb * g[2 % C];
				} else if (kPacked) {
					// two chains of packed FMAs (even / odd 16-B parts), four partial sums folded at the end
					const float4 *f4 = reinterpret_cast<const float4 *>(ent + TR::kRecParts * 4);
					uint64_t da = pack2(d0, 0.f), db = 0ull;
#pragma unroll
					for (int q = 0; q < C / 4; q++) {
						const float4 f = f4[q];
						da = fma2(pack2(f.x, f.y), g2[(2 * q) % kPairs], da);
						db = fma2(pack2(f.z, f.w), g2[(2 * q + 1) % kPairs], db);
					}
					float lo, hi;
					unpack2(add2(da, db), lo, hi);
					d0 = lo + hi;
				} else {
					const float4 *f4 = reinterpret_cast<const float4 *>(ent + TR::kRecParts * 4);
#pragma unroll
					for (int q = 0; q < C / 4; q++) {
						const float4 f = f4[q];
						d0 += f.x * g[4 * q + 0];
						d1 += f.y * g[4 * q + 1];
						d2 += f.z * g[4 * q + 2];
						d3 += f.w * g[4 * q + 3];
					}
				}
#if GSR_BWD_ACC4
				const float dot = (d0 + d1) + (d2 + d3);
#else
				const float dot = d0;
#endif
				X = last_alpha * last_dot + (1.f - last_alpha) * X;
				last_dot = dot;
				float dL_dalpha = (dot - X) * T;
				last_alpha = alpha;
#if GSR_BWD_RCP
				dL_dalpha += neg_Tfinal_bg * rinv;
#else
				if (bg_dot != 0.f) dL_dalpha += (-T_final / one_minus) * bg_dot;
#endif

				const float dL_dG = r1.y * dL_dalpha;
				const float gdx = G * d.x, gdy = G * d.y;
				const float dG_ddelx = -gdx * r0.z - gdy * r0.w;
				const float dG_ddely = -gdy * r1.x - gdx * r0.w;
				v[0] = dL_dG * dG_ddelx * ddelx_dx;
				v[1] = dL_dG * dG_ddely * ddely_dy;
				v[2] = -0.5f * gdx * d.x * dL_dG;
				v[3] = -0.5f * gdx * d.y * dL_dG;
				v[4] = -0.5f * gdy * d.y * dL_dG;
				v[5] = G * dL_dalpha;
				v[6] = w * gd;
				v[7] = w * gu;
				if (!kLaneChannel) {
#pragma unroll
					for (int ch = 0; ch < C; ch++) v[8 + ch] = w * g[ch];
				}
			}
			const uint32_t id = feed.q_id[slot];

			// The butterfly's 9 dependent shuffles and the colour sums below are independent, but ptxas keeps them apart whatever the
			// source order (weights stored first, hand-interleaved levels, butterfly deferred into the next entry's geometry block:
			// measured 2.158 / 2.158 / 2.476 ms against 2.148 ms, profiles/r1_bwd_order_ab.md; code at commit 63eca77).
			warp_transpose_reduce<NV>(v, lane);
			if (vowner<NV>(lane)) {
				const int q = vidx<NV>(lane);
				if (q < 8) red_add(gacc + (size_t)id * 8 + q, v[0]);
				else if (q < 8 + C) red_add(dL_dcolors + (size_t)id * C + (q - 8), v[0]);
			}
			if (kLaneChannel) {
				s_w[lwarp][lane] = w;
				__syncwarp();
#if GSR_BWD_ACC4
				float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#elif GSR_BWD_PB2
				float s0 = 0.f, s1 = 0.f;
				float &s2 = s0, &s3 = s1;
#else
				float s0 = 0.f;
				float &s1 = s0, &s2 = s0, &s3 = s0;
#endif
				const float4 *w4 = reinterpret_cast<const float4 *>(s_w[lwarp]);
				if (kPacked) {
					uint64_t sa = 0ull, sb = 0ull;
#pragma unroll
					for (int q = 0; q < 8; q++) {
						const float4 ww = w4[q];
						sa = fma2(pack2(ww.x, ww.y), gcol2[(2 * q) % kPairs], sa);
						sb = fma2(pack2(ww.z, ww.w), gcol2[(2 * q + 1) % kPairs], sb);
					}
					float lo, hi;
					unpack2(add2(sa, sb), lo, hi);
					red_add(dL_dcolors + (size_t)id * C + lane, lo + hi); // 32 lanes -> one coalesced 128-B RED
					__syncwarp();
					continue;
				}
				const float4 *c4 = reinterpret_cast<const float4 *>(s_gc + lane * kGcolStride);
#pragma unroll
				for (int q = 0; q < 8; q++) {
					const float4 ww = w4[q];
					if (kGcolSmem) {
						const float4 gc = c4[q];
						s0 += ww.x * gc.x;
						s1 += ww.y * gc.y;
						s2 += ww.z * gc.z;
						s3 += ww.w * gc.w;
					} else {
						s0 += ww.x * gcol[(4 * q + 0) % (kGcolSmem ? 1 : 32)];
						s1 += ww.y * gcol[(4 * q + 1) % (kGcolSmem ? 1 : 32)];
						s2 += ww.z * gcol[(4 * q + 2) % (kGcolSmem ? 1 : 32)];
						s3 += ww.w * gcol[(4 * q + 3) % (kGcolSmem ? 1 : 32)];
					}
				}
#if GSR_BWD_ACC4
				red_add(dL_dcolors + (size_t)id * C + lane, (s0 + s1) + (s2 + s3)); // 32 lanes -> one coalesced 128-B RED
#elif GSR_BWD_PB2
				red_add(dL_dcolors + (size_t)id * C + lane, s0 + s1);
#else
				red_add(dL_dcolors + (size_t)id * C + lane, s0);
#endif
				__syncwarp();
			}
		}
		feed.done += m_cur;
		__syncwarp(); // the stage buffer and the ring slots of this chunk may be reused
		m_cur = m_next;
	}
	feed.drain(chunk, 0);
}

template <int C>
static size_t bwd_smem_bytes()
{
	return (size_t)kWarpsPerCta * (BlendTraits<C>::kWarpBytes + ((GSR_BWD_FEED_BULK != 0 && C > 3) ? 16 : 0) + ((C == 32 && GSR_BWD_GCOL_SMEM) ? 32 * 36 * 4 : 0));
}

template <int C>
static cudaError_t launch_bwd(int tiles, const uint2 *ranges, const uint32_t *point_list, int packed, int W, int H, int tiles_x, const float *rec,
                              const float *features, const float *bg, const float *final_Ts, const uint32_t *n_contrib,
                              const float *dL_dpixels, const float *dL_dpixel_depths, const float *dL_dpixel_uncs, float *gacc,
                              float *dL_dcolors, cudaStream_t stream)
{
	using TR = BlendTraits<C>;
	static bool configured = false;
	if (!configured) {
		cudaError_t e = cudaFuncSetAttribute(blend_backward_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem_bytes<C>());
		if (e != cudaSuccess) return e;
		configured = true;
	}
	blend_backward_kernel<C><<<tiles * kCtasPerTile, 32 * kWarpsPerCta, bwd_smem_bytes<C>(), stream>>>(ranges, point_list, packed, W, H, tiles_x, rec, features, bg, final_Ts, n_contrib,
	                                                                dL_dpixels, dL_dpixel_depths, dL_dpixel_uncs, gacc, dL_dcolors);
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_blend_backward(int C, int P, int W, int H, const uint2 *ranges, const uint32_t *point_list, const float *rec,
                                  const float *features, const float *bg, const float *final_Ts, const uint32_t *n_contrib,
                                  const float *dL_dpixels, const float *dL_dpixel_depths, const float *dL_dpixel_uncs, float *gacc,
                                  float *dL_dcolors, cudaStream_t stream)
{
	const int tiles_x = (W + GSR_BLOCK_X - 1) / GSR_BLOCK_X, tiles_y = (H + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y;
	const int tiles = tiles_x * tiles_y;
	if (tiles <= 0) return cudaSuccess;
	const int packed = point_list_packed(P) ? 1 : 0;
#if GSR_BWD_MMA
	if (C == 32)
		return launch_blend_backward_mma(P, W, H, ranges, point_list, rec, features, bg, final_Ts, n_contrib, dL_dpixels, dL_dpixel_depths, dL_dpixel_uncs, gacc, dL_dcolors, stream);
#endif
	switch (C) {
	case 3: return launch_bwd<3>(tiles, ranges, point_list, packed, W, H, tiles_x, rec, features, bg, final_Ts, n_contrib, dL_dpixels, dL_dpixel_depths, dL_dpixel_uncs, gacc, dL_dcolors, stream);
	case 32: return launch_bwd<32>(tiles, ranges, point_list, packed, W, H, tiles_x, rec, features, bg, final_Ts, n_contrib, dL_dpixels, dL_dpixel_depths, dL_dpixel_uncs, gacc, dL_dcolors, stream);
	default: return cudaErrorInvalidValue;
	}
}

} // namespace gsr
