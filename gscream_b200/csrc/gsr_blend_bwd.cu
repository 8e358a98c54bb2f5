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
//   * Every term the reference adds atomically is linear in two per-(pixel, Gaussian) scalars,
//         s = G * dL_dalpha      and      w = alpha * T,
//     with coefficients that depend only on the Gaussian's record and the pixel's position / gradient row.  A warp
//     therefore works on a landed chunk of its feed (gsr_blend.cuh) in two phases:
//       phase 1 (lane = pixel, sequential in depth): the recurrence; leaves s and w of every (entry, pixel) of the chunk
//                in 4 KB of shared memory.  Nothing in it waits for a reduction.
//       phase 2 (entries independent): the per-Gaussian sums over the warp's 32 pixels.  The 8 geometric / scalar terms are
//                rebuilt from s, w and the staged record and reduced by a transpose-reduce butterfly over TWO entries at a
//                time (16 shuffles per pair, the first level free of selects because the upper half-warp loads the pair
//                swapped); the C = 32 colour terms by switching roles — lane l owns channel l and holds the gradient COLUMN
//                of the warp's 32 pixels in registers — ending in one coalesced 128-B red.global.add per (warp, Gaussian)
//                where the reference issues 32 x C scalar atomics.
//     Round 1 ran both per entry, the butterfly's 9 dependent shuffles and the colour sums' FMA chain on the recurrence's
//     critical path: 63 % issue-slot utilisation at 16 warps per SM (profiles/r1_blend_v7_summary.md).  Separating the phases
//     lets ptxas interleave independent entries' chains (profiles/r2_blend_bwd.md).
//   * Same warp-private feed as the forward kernel (per-warp list scan by the instance masks, private double-buffered
//     cp.async gather, no block barrier), run back to front and started at the warp's own deepest last contributor.
//
// Scalar terms are accumulated into gacc[P][8] = {dmean2D.x, dmean2D.y, dconic.x, dconic.y, dconic.w,
// dopacity, ddepth, duncertainty}; colours into dL_dcolors[P][C].  Both must be zero (or hold the
// running sum) on entry.
#include "gsr_blend.cuh"
#include "gsr_internal.cuh"

namespace gsr {

constexpr int kWarpsPerCta = GSR_BWD_WARPS_PER_CTA;
constexpr int kCtasPerTile = kWarpsPerTile / kWarpsPerCta;

// Transpose-reduce over the lane-index bits S, S/2, ..: every lane contributes N values; afterwards v[0] on lane l is
// the total (over the lanes that differ from l in those bits and all lower ones) of value index vidx<N, S>(l).
// N/2 + N/4 + ... + 1 exchanges, then plain xor-adds for the remaining strides.
template <int N, int S>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[N], int lane)
{
	int s = S;
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
template <int N, int S>
__device__ __forceinline__ int vidx(int lane)
{
	int idx = 0, s = S;
#pragma unroll
	for (int n = N / 2; n >= 1; n >>= 1, s >>= 1)
		if (lane & s) idx += n;
	return idx;
}

// The eight geometric / scalar terms of one (pixel, Gaussian) pair (CR/backward.cu:557-601) from s = G dL_dalpha and
// w = alpha T:  dL_dG G = opacity * s, so
//   dmean2D.x = -(o s) (dx a + dy b) W/2, dmean2D.y = -(o s) (dy c + dx b) H/2, dconic = -1/2 (o s) {dx dx, dx dy, dy dy},
//   dopacity = s, ddepth = w g_depth, duncertainty = w g_unc.
__device__ __forceinline__ void pair_terms(const float *ent, float s, float w, float pixf_x, float pixf_y, float gd, float gu,
                                           float half_w, float half_h, float *v)
{
	const float4 r0 = *reinterpret_cast<const float4 *>(ent);     // x y a b
	const float2 r1 = *reinterpret_cast<const float2 *>(ent + 4); // c o
	const float dx = r0.x - pixf_x, dy = r0.y - pixf_y;
	const float u = r1.y * s;
	const float udx = u * dx, udy = u * dy;
	v[0] = -(udx * r0.z + udy * r0.w) * half_w;
	v[1] = -(udy * r1.x + udx * r0.w) * half_h;
	v[2] = -0.5f * udx * dx;
	v[3] = -0.5f * udx * dy;
	v[4] = -0.5f * udy * dy;
	v[5] = s;
	v[6] = w * gd;
	v[7] = w * gu;
}

// C = 32 keeps a 34-float gradient row and a 32-float gradient column per lane: 2 CTAs of 8 warps' worth of registers per
// SM (16 warps); C <= 8 fits 3.
#ifndef GSR_BWD_MINBLOCKS
#define GSR_BWD_MINBLOCKS(C) ((C) <= 8 ? 3 : 2)
#endif
#define GSR_BWD_MINCTAS(C) (GSR_BWD_MINBLOCKS(C) * kCtasPerTile)

// entries of a chunk whose state-independent parts are computed side by side in phase 1
#ifndef GSR_BWD_SUB
#define GSR_BWD_SUB 2
#endif
constexpr int kSub = GSR_BWD_SUB;
static_assert(kChunk % kSub == 0, "sub-batches tile the chunk");

// 1 / d for d in [0.01, 1]: MUFU.RCP + one Newton step, no range check (__frcp_rn's slow path is a call behind a branch)
__device__ __forceinline__ float rcp_1ulp(float d)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
	const float e = __fmaf_rn(-d, r, 1.f);
	return __fmaf_rn(r, e, r);
}

template <int C>
struct BwdSmem {
	using TR = BlendTraits<C>;
	static constexpr int kHandoffBytes = 2 * kChunk * 32 * 4; // s[kChunk][32], w[kChunk][32]
	static constexpr int kWarpBytes = TR::kWarpBytes + kHandoffBytes;
};

template <int C>
__global__ void __launch_bounds__(32 * kWarpsPerCta, GSR_BWD_MINCTAS(C)) blend_backward_kernel(
    const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, const uint32_t *__restrict__ header, int W, int H, int tiles_x,
    const float *__restrict__ rec, const float *__restrict__ features, const float *__restrict__ bg,
    const float *__restrict__ final_Ts, const uint32_t *__restrict__ n_contrib,
    const float *__restrict__ dL_dpixels, const float *__restrict__ dL_dpixel_depths, const float *__restrict__ dL_dpixel_uncs,
    float *__restrict__ gacc, float *__restrict__ dL_dcolors)
{
	using TR = BlendTraits<C>;
	static_assert(C <= 8, "the butterfly carries at most 8 colour channels; C = 32 has its own kernel below");

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
	const size_t plane = (size_t)H * W;
	const size_t pix_id = (size_t)W * py + px;

	const uint2 range = ranges[tile];
	const int packed = (int)__ldg(header + kHdrPacked); // format of the list entries, recorded by the instance emission
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
	float T = T_final;
	float X = 0.f, last_alpha = 0.f, last_dot = 0.f;
	const float half_w = 0.5f * (float)W, half_h = 0.5f * (float)H;
	const float neg_Tfinal_bg = -T_final * bg_dot; // background term: (-T_final / (1 - alpha)) * sum_ch bg[ch] g[ch]

	unsigned char *warp_smem = smem_raw + (size_t)lwarp * BwdSmem<C>::kWarpBytes;
	float *s_s = reinterpret_cast<float *>(warp_smem + TR::kWarpBytes); // [kChunk][32]
	float *s_w = s_s + kChunk * 32;                                     // [kChunk][32]

	// back to front (CR/backward.cu:500): the feed scans list positions warp_last-1 .. 0
	using Feed = WarpFeed<C, true>;
	Feed feed;
	feed.init(warp_smem, point_list + range.x, warp_last, rec, features, warp, lane, packed != 0);
	feed.fill();
	int m_cur = feed.issue(0);
	int chunk = 0;
	for (; m_cur > 0; chunk++) {
		feed.fill();
		const int m_next = feed.issue((chunk + 1) & 1);
		feed.wait();
		__syncwarp(); // every lane's copies of this chunk have landed
		const float *ent0 = feed.stage + (chunk & 1) * TR::kStageFloats;

		// ---- phase 1: the per-pixel recurrence over the chunk's entries, in depth order ----
		// kSub entries at a time: (a) everything that does not depend on the pixel's running state — alpha, 1 / (1 - alpha), G
		// and the dot product, kSub independent chains for the scheduler to interleave — then (b) the short carried chain
		// (T, X).  An entry that does not touch the pixel is encoded as alpha = 0, rinv = 1, G = 0: the recurrence below then
		// leaves T unchanged, hands X on unchanged (0 * dot + 1 * X') and produces s = w = 0, without a branch.
		uint32_t live = 0; // bit e: some pixel of the warp received gradient from entry e
		for (int e0 = 0; e0 < m_cur; e0 += kSub) {
			float al[kSub], ri[kSub], Gs[kSub], dt[kSub];
#pragma unroll
			for (int b = 0; b < kSub; b++) {
				const int e = e0 + b;
				const float *ent = ent0 + e * TR::kEntryFloats; // (rows past m_cur hold stale but readable shared memory)
				const int pos = (int)feed.q_pos[(feed.done + e) & (kRing - 1)]; // 0-based list position
				const float4 r0 = *reinterpret_cast<const float4 *>(ent);     // x y a b
				const float4 r1 = *reinterpret_cast<const float4 *>(ent + 4); // c o depth unc
				const float dx = r0.x - pixf_x, dy = r0.y - pixf_y;
				const float power = gaussian_power(r0.z, r0.w, r1.x, dx, dy);
				const float G = expf(power);
				const float alpha = min(0.99f, __fmul_rn(r1.y, G));
				const bool valid = (e < m_cur) && (pos < last_contributor) && !(power > 0.0f) && !(alpha < kAlphaMin);
				// dot = f_j . g_p over colour channels, depth and uncertainty
				float d0 = r1.z * gd + r1.w * gu;
				if (TR::kFeatInRec) {
					const float4 r2 = *reinterpret_cast<const float4 *>(ent + 8);
					const float cb = ent[12];
					if (C > 0) d0 += r2.z * g[0];
					if (C > 1) d0 += r2.w * g[1 % C];
					if (C > 2) d0 += cb * g[2 % C];
				} else {
					float d1 = 0.f, d2 = 0.f, d3 = 0.f;
					const float4 *f4 = reinterpret_cast<const float4 *>(ent + TR::kRecParts * 4);
#pragma unroll
					for (int q = 0; q < C / 4; q++) {
						const float4 f = f4[q];
						d0 += f.x * g[(4 * q + 0) % C];
						d1 += f.y * g[(4 * q + 1) % C];
						d2 += f.z * g[(4 * q + 2) % C];
						d3 += f.w * g[(4 * q + 3) % C];
					}
					d0 = (d0 + d1) + (d2 + d3);
				}
				al[b] = valid ? alpha : 0.f;
				ri[b] = valid ? rcp_1ulp(__fsub_rn(1.f, alpha)) : 1.f; // T / (1 - alpha) (CR/backward.cu:533) as T * rcp; also serves the background term
				Gs[b] = valid ? G : 0.f;
				dt[b] = valid ? d0 : 0.f;
				if (__any_sync(0xffffffffu, valid)) live |= 1u << e;
			}
#pragma unroll
			for (int b = 0; b < kSub; b++) {
				T *= ri[b];
				const float Xn = last_alpha * last_dot + (1.f - last_alpha) * X;
				const float dL_dalpha = (dt[b] - Xn) * T + neg_Tfinal_bg * ri[b];
				s_s[(e0 + b) * 32 + lane] = Gs[b] * dL_dalpha;
				s_w[(e0 + b) * 32 + lane] = al[b] * T;
				X = Xn;
				last_alpha = al[b];
				last_dot = dt[b];
			}
		}
		__syncwarp(); // s and w of the chunk are visible to every lane

		// ---- phase 2: per-Gaussian sums over the warp's 32 pixels; entries are independent of each other ----
		{
			while (live) {
				const int e = __ffs(live) - 1;
				live &= live - 1;
				float v[16];
				const float w = s_w[e * 32 + lane];
				pair_terms(ent0 + e * TR::kEntryFloats, s_s[e * 32 + lane], w, pixf_x, pixf_y, gd, gu, half_w, half_h, v);
#pragma unroll
				for (int ch = 0; ch < 8; ch++) v[8 + ch] = ch < C ? w * g[ch % C] : 0.f;
				warp_transpose_reduce<16, 16>(v, lane);
				if ((lane & 1) == 0) {
					const uint32_t id = feed.q_id[(feed.done + e) & (kRing - 1)];
					const int q = vidx<16, 16>(lane);
					if (q < 8) red_add(gacc + (size_t)id * 8 + q, v[0]);
					else if (q < 8 + C) red_add(dL_dcolors + (size_t)id * C + (q - 8), v[0]);
				}
			}
		}
		feed.done += m_cur;
		__syncwarp(); // the stage buffer, the ring slots and the s / w rows of this chunk may be reused
		m_cur = m_next;
	}
	feed.drain();
}

// ---- C = 32 -----------------------------------------------------------------------------------------------------------------
// With 32 feature channels the two dense products per (block, chunk) dominate everything else:
//     dot[p][e]  = sum_ch g[p][ch] f[e][ch]        (32 pixels x 16 entries x 32 channels)   -> feeds the recurrence
//     dcol[e][ch] = sum_p  w[e][p] g[p][ch]        (16 entries x 32 channels x 32 pixels)   -> dL_dcolors
// On the FP32 pipe they cost 64 FFMA + 16 broadcast LDS.128 per (lane, entry); round 2's first build measured the shared-memory
// data pipe at 72 % next to 72 % issue utilisation — a broadcast LDS.128 is two wavefronts for 16 useful bytes
// (profiles/r2_blend_bwd.md).  Here both run on the tensor pipe as mma.sync m16n8k8 TF32 with a 3xTF32 split (x = hi + lo,
// hi = top 19 bits; hi*hi + lo*hi + hi*lo, error ~2^-21 per product): operands are read as fragments — every lane a different
// word, one wavefront per 32 words — and the gradient block lives in registers twice, as the A operand of the first product
// (gA: 4 pixels x 8 channels per lane) and the B operand of the second (gB: 8 pixels x 4 channels per lane).
// The contraction index of all three products is permuted (k-step ks, slot (t, half) <-> 16 (ks >> 1) + 4 t + 2 (ks & 1) + half)
// so that one LDS.128 fetches a lane's operands of two k-steps.
// The eight scalar sums go the same way.  With block-centred pixel coordinates (cx, cy) and (ax, ay) = Gaussian centre relative
// to the block centre, dx = ax - cx, so every geometric term is a fixed linear map (per Gaussian) of the six moments
//     M = sum_p s[p] {1, cx, cy, cx^2, cx cy, cy^2}
// — a [16 x 32] x [32 x 8] product whose B operand is the same for every chunk and exact in TF32 — and ddepth / duncertainty are
// columns 32, 33 of the colour product.  |ax| exceeds |dx| by at most 3.5, which bounds the cancellation in the map.
// pixel (or channel) index of k-step ks, slot (t, half)
__device__ __forceinline__ int kperm(int ks, int t, int half) { return 16 * (ks >> 1) + 4 * t + 2 * (ks & 1) + half; }

constexpr int kDotStride = 20;     // floats per pixel row of the dot tile (16 entries + pad: conflict-free LDS.128 by lane = pixel)
constexpr int kSub32 = 4;          // entries per recurrence sub-batch (one LDS.128 of dots)
struct Bwd32Smem {
	using TR = BlendTraits<32>;
	static constexpr int kS = TR::kWarpBytes, kW = kS + kChunk * 32 * 4, kDot = kW + kChunk * 32 * 4;
	static constexpr int kWarpBytes = kDot + 32 * kDotStride * 4;
};
#ifndef GSR_BWD32_MINCTAS
#define GSR_BWD32_MINCTAS 6
#endif

__global__ void __launch_bounds__(32 * kWarpsPerCta, GSR_BWD32_MINCTAS) blend_backward_c32_kernel(
    const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, const uint32_t *__restrict__ header, int W, int H, int tiles_x,
    const float *__restrict__ rec, const float *__restrict__ features, const float *__restrict__ bg,
    const float *__restrict__ final_Ts, const uint32_t *__restrict__ n_contrib,
    const float *__restrict__ dL_dpixels, const float *__restrict__ dL_dpixel_depths, const float *__restrict__ dL_dpixel_uncs,
    float *__restrict__ gacc, float *__restrict__ dL_dcolors)
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
	const size_t plane = (size_t)H * W;
	const size_t pix_id = (size_t)W * py + px;

	const uint2 range = ranges[tile];
	const int packed = (int)__ldg(header + kHdrPacked); // format of the list entries, recorded by the instance emission
	const float T_final = inside ? final_Ts[pix_id] : 0.f;
	const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

	// deepest last contributor of the warp: nothing behind it can receive gradient from these 32 pixels
	int warp_last = last_contributor;
#pragma unroll
	for (int s = 16; s >= 1; s >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, s));
	warp_last = min(warp_last, (int)(range.y - range.x));
	if (warp_last == 0) return; // warps are independent: no barrier follows

	// pixel p of the block (p = 8 row + col) -> offset in a plane, or -1 outside the image
	auto pix_off = [&](int p) -> long long {
		const int x = x0 + (p & 7), y = y0 + (p >> 3);
		return (x < W && y < H) ? (long long)W * y + x : -1;
	};
	// the block's 32 x 32 upstream colour gradient, as the A operand of the dot product (rows = pixels, k = channels) ...
	float gA[2][4][4];
#pragma unroll
	for (int mt = 0; mt < 2; mt++)
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const long long off = pix_off(16 * mt + gq + 8 * (j & 1));
#pragma unroll
			for (int ks = 0; ks < 4; ks++) gA[mt][ks][j] = off >= 0 ? __ldg(dL_dpixels + (size_t)kperm(ks, t, j >> 1) * plane + off) : 0.f;
		}
	// ... and as the B operand of the colour sums (k = pixels, columns = channels): channel 8 nt + gq, pixels 16 kk + 4 t .. + 3
	float gB[4][4][2];
	{
		const bool vec_ok = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(dL_dpixels) & 15) == 0) && (x0 + 8 <= W);
#pragma unroll
		for (int nt = 0; nt < 4; nt++)
#pragma unroll
			for (int kk = 0; kk < 2; kk++) {
				const float *src = dL_dpixels + (size_t)(8 * nt + gq) * plane;
				const int y = y0 + 2 * kk + (t >> 1), xx = x0 + 4 * (t & 1);
				float v[4];
				if (vec_ok && y < H) {
					const float4 q = __ldg(reinterpret_cast<const float4 *>(src + (size_t)W * y + xx));
					v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
				} else {
#pragma unroll
					for (int j = 0; j < 4; j++) v[j] = (xx + j < W && y < H) ? __ldg(src + (size_t)W * y + xx + j) : 0.f;
				}
				gB[2 * kk][nt][0] = v[0]; gB[2 * kk][nt][1] = v[1];
				gB[2 * kk + 1][nt][0] = v[2]; gB[2 * kk + 1][nt][1] = v[3];
			}
	}
	// columns 32, 33 of that operand: the depth and uncertainty gradients (lanes gq = 0, 1; the other columns of the tile are zero)
	float gY[4][2];
#pragma unroll
	for (int ks = 0; ks < 4; ks++)
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const long long off = pix_off(kperm(ks, t, h));
			gY[ks][h] = (gq < 2 && off >= 0) ? __ldg((gq == 0 ? dL_dpixel_depths : dL_dpixel_uncs) + off) : 0.f;
		}

	// operand of the moment product: column gq is the monomial {1, cx, cy, cx^2, cx cy, cy^2, 0, 0}[gq] of the slot's pixel in
	// block-centred coordinates
	uint32_t gM[4][2];
#pragma unroll
	for (int ks = 0; ks < 4; ks++)
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const int p = kperm(ks, t, h);
			const float cx = (float)(p & 7) - 3.5f, cy = (float)(p >> 3) - 1.5f;
			const float ex = (gq == 1 || gq == 4) ? cx : gq == 3 ? cx * cx : 1.f;      // cx^i
			const float ey = (gq == 2 || gq == 4) ? cy : gq == 5 ? cy * cy : 1.f;      // cy^j
			gM[ks][h] = __float_as_uint(gq < 6 ? ex * ey : 0.f);
		}

	// lane = pixel quantities
	const float gd = inside ? dL_dpixel_depths[pix_id] : 0.f, gu = inside ? dL_dpixel_uncs[pix_id] : 0.f;
	float bg_dot = 0.f;
	{
		const float bgv = bg[lane];
		if (__any_sync(0xffffffffu, bgv != 0.f)) {
#pragma unroll 4
			for (int ch = 0; ch < C; ch++) bg_dot += __shfl_sync(0xffffffffu, bgv, ch) * (inside ? __ldg(dL_dpixels + ch * plane + pix_id) : 0.f);
		}
	}
	float T = T_final;
	float X = 0.f, last_alpha = 0.f, last_dot = 0.f;
	const float half_w = 0.5f * (float)W, half_h = 0.5f * (float)H;
	const float neg_Tfinal_bg = -T_final * bg_dot; // background term: (-T_final / (1 - alpha)) * sum_ch bg[ch] g[ch]
	const float blk_cx = (float)x0 + 3.5f, blk_cy = (float)y0 + 1.5f;

	unsigned char *warp_smem = smem_raw + (size_t)lwarp * Bwd32Smem::kWarpBytes;
	float *s_s = reinterpret_cast<float *>(warp_smem + Bwd32Smem::kS);   // [kChunk][32]
	float *s_w = reinterpret_cast<float *>(warp_smem + Bwd32Smem::kW);   // [kChunk][32]
	float *s_dot = reinterpret_cast<float *>(warp_smem + Bwd32Smem::kDot); // [32][kDotStride]

	// back to front (CR/backward.cu:500): the feed scans list positions warp_last-1 .. 0
	using Feed = WarpFeed<C, true>;
	Feed feed;
	feed.init(warp_smem, point_list + range.x, warp_last, rec, features, warp, lane, packed != 0);
	feed.fill();
	int m_cur = feed.issue(0);
	int chunk = 0;
	for (; m_cur > 0; chunk++) {
		feed.fill();
		const int m_next = feed.issue((chunk + 1) & 1);
		feed.wait();
		__syncwarp(); // every lane's copies of this chunk have landed
		const float *ent0 = feed.stage + (chunk & 1) * TR::kStageFloats;

		// ---- phase 0: dot[p][e] for the whole chunk on the tensor pipe (rows past m_cur: stale operands, results unused) ----
		{
			float D[2][2][4];
#pragma unroll
			for (int mt = 0; mt < 2; mt++)
#pragma unroll
				for (int nt = 0; nt < 2; nt++)
#pragma unroll
					for (int i = 0; i < 4; i++) D[mt][nt][i] = 0.f;
#pragma unroll
			for (int kk = 0; kk < 2; kk++) {
				uint32_t bh[2][4], bl[2][4];
#pragma unroll
				for (int nt = 0; nt < 2; nt++) {
					const float4 f = *reinterpret_cast<const float4 *>(ent0 + (8 * nt + gq) * TR::kEntryFloats + TR::kRecParts * 4 + 16 * kk + 4 * t);
					split_tf32(f.x, bh[nt][0], bl[nt][0]);
					split_tf32(f.y, bh[nt][1], bl[nt][1]);
					split_tf32(f.z, bh[nt][2], bl[nt][2]);
					split_tf32(f.w, bh[nt][3], bl[nt][3]);
				}
#pragma unroll
				for (int sub = 0; sub < 2; sub++) {
#pragma unroll
					for (int mt = 0; mt < 2; mt++) {
						uint32_t ah[4], al[4];
#pragma unroll
						for (int j = 0; j < 4; j++) split_tf32(gA[mt][2 * kk + sub][j], ah[j], al[j]);
#pragma unroll
						for (int nt = 0; nt < 2; nt++) mma3_tf32(D[mt][nt], ah, al, bh[nt][2 * sub], bh[nt][2 * sub + 1], bl[nt][2 * sub], bl[nt][2 * sub + 1]);
					}
				}
			}
#pragma unroll
			for (int mt = 0; mt < 2; mt++)
#pragma unroll
				for (int nt = 0; nt < 2; nt++) {
					*reinterpret_cast<float2 *>(s_dot + (16 * mt + gq) * kDotStride + 8 * nt + 2 * t) = make_float2(D[mt][nt][0], D[mt][nt][1]);
					*reinterpret_cast<float2 *>(s_dot + (16 * mt + gq + 8) * kDotStride + 8 * nt + 2 * t) = make_float2(D[mt][nt][2], D[mt][nt][3]);
				}
		}
		__syncwarp();

		// ---- phase 1 (lane = pixel): the recurrence over the chunk's entries, in depth order ----
		// kSub32 entries at a time: (a) everything that does not depend on the pixel's running state — alpha, 1 / (1 - alpha), G —
		// then (b) the short carried chain (T, X).  An entry that does not touch the pixel is encoded as alpha = 0, rinv = 1,
		// G = 0: the recurrence then leaves T unchanged, hands X on unchanged (0 * dot + 1 * X') and produces s = w = 0.
		uint32_t live = 0; // bit e: some pixel of the warp received gradient from entry e
		for (int e0 = 0; e0 < m_cur; e0 += kSub32) {
			const float4 d4 = *reinterpret_cast<const float4 *>(s_dot + lane * kDotStride + e0);
			const float dcol[4] = {d4.x, d4.y, d4.z, d4.w};
			float al[kSub32], ri[kSub32], Gs[kSub32], dt[kSub32];
#pragma unroll
			for (int b = 0; b < kSub32; b++) {
				const int e = e0 + b;
				const float *ent = ent0 + e * TR::kEntryFloats;
				const int pos = (int)feed.q_pos[(feed.done + e) & (kRing - 1)]; // 0-based list position
				const float4 r0 = *reinterpret_cast<const float4 *>(ent);     // x y a b
				const float4 r1 = *reinterpret_cast<const float4 *>(ent + 4); // c o depth unc
				const float dx = r0.x - pixf_x, dy = r0.y - pixf_y;
				const float power = gaussian_power(r0.z, r0.w, r1.x, dx, dy);
				const float G = expf(power);
				const float alpha = min(0.99f, __fmul_rn(r1.y, G));
				const bool valid = (e < m_cur) && (pos < last_contributor) && !(power > 0.0f) && !(alpha < kAlphaMin);
				al[b] = valid ? alpha : 0.f;
				ri[b] = valid ? rcp_1ulp(__fsub_rn(1.f, alpha)) : 1.f; // T / (1 - alpha) (CR/backward.cu:533) as T * rcp; also serves the background term
				Gs[b] = valid ? G : 0.f;
				dt[b] = valid ? dcol[b] + (r1.z * gd + r1.w * gu) : 0.f; // f_j . g_p over colour channels, depth and uncertainty
				if (__any_sync(0xffffffffu, valid)) live |= 1u << e;
			}
#pragma unroll
			for (int b = 0; b < kSub32; b++) {
				T *= ri[b];
				const float Xn = last_alpha * last_dot + (1.f - last_alpha) * X;
				const float dL_dalpha = (dt[b] - Xn) * T + neg_Tfinal_bg * ri[b];
				s_s[(e0 + b) * 32 + lane] = Gs[b] * dL_dalpha;
				s_w[(e0 + b) * 32 + lane] = al[b] * T;
				X = Xn;
				last_alpha = al[b];
				last_dot = dt[b];
			}
		}
		__syncwarp(); // s and w of the chunk are visible to every lane

		// ---- phase 2: per-Gaussian sums over the warp's 32 pixels on the tensor pipe (rows = entries gq, gq + 8) ----
		if (live) {
			float Dc[4][4], Dy[4], Dm[4];
#pragma unroll
			for (int i = 0; i < 4; i++) {
				Dy[i] = 0.f;
				Dm[i] = 0.f;
#pragma unroll
				for (int nt = 0; nt < 4; nt++) Dc[nt][i] = 0.f;
			}
#pragma unroll
			for (int kk = 0; kk < 2; kk++) {
				const float4 w0 = *reinterpret_cast<const float4 *>(s_w + gq * 32 + 16 * kk + 4 * t);
				const float4 w1 = *reinterpret_cast<const float4 *>(s_w + (gq + 8) * 32 + 16 * kk + 4 * t);
				const float4 s0 = *reinterpret_cast<const float4 *>(s_s + gq * 32 + 16 * kk + 4 * t);
				const float4 s1 = *reinterpret_cast<const float4 *>(s_s + (gq + 8) * 32 + 16 * kk + 4 * t);
				const float wv[2][4] = {{w0.x, w1.x, w0.y, w1.y}, {w0.z, w1.z, w0.w, w1.w}}; // [sub][a0..a3]
				const float sv[2][4] = {{s0.x, s1.x, s0.y, s1.y}, {s0.z, s1.z, s0.w, s1.w}};
#pragma unroll
				for (int sub = 0; sub < 2; sub++) {
					const int ks = 2 * kk + sub;
					uint32_t ah[4], al[4];
#pragma unroll
					for (int j = 0; j < 4; j++) split_tf32(wv[sub][j], ah[j], al[j]);
#pragma unroll
					for (int nt = 0; nt < 4; nt++) {
						uint32_t b0h, b0l, b1h, b1l;
						split_tf32(gB[ks][nt][0], b0h, b0l);
						split_tf32(gB[ks][nt][1], b1h, b1l);
						mma3_tf32(Dc[nt], ah, al, b0h, b1h, b0l, b1l);
					}
					{
						uint32_t b0h, b0l, b1h, b1l;
						split_tf32(gY[ks][0], b0h, b0l);
						split_tf32(gY[ks][1], b1h, b1l);
						mma3_tf32(Dy, ah, al, b0h, b1h, b0l, b1l);
					}
#pragma unroll
					for (int j = 0; j < 4; j++) split_tf32(sv[sub][j], ah[j], al[j]);
					// moments: the operand (gM) is exact in TF32, so no low part
					mma_tf32(Dm, al, gM[ks][0], gM[ks][1]);
					mma_tf32(Dm, ah, gM[ks][0], gM[ks][1]);
				}
			}
			// rows gq (c0, c1) and gq + 8 (c2, c3) of the results belong to entries gq and gq + 8
#pragma unroll
			for (int r = 0; r < 2; r++) {
				const int e = gq + 8 * r;
				const bool on = e < m_cur && ((live >> e) & 1u);
				const uint32_t id = feed.q_id[(feed.done + e) & (kRing - 1)];
				if (on) {
#pragma unroll
					for (int nt = 0; nt < 4; nt++) red_add_v2(dL_dcolors + (size_t)id * C + 8 * nt + 2 * t, Dc[nt][2 * r], Dc[nt][2 * r + 1]);
				}
				// the quad's four lanes hold the row's moments pairwise (t = 0: M0 Mx, 1: My Mxx, 2: Mxy Myy) and (t = 0) the depth /
				// uncertainty sums; every lane fetches all of them and forms its own pair of the eight outputs
				const float M0 = __shfl_sync(0xffffffffu, Dm[2 * r], 0, 4), Mx = __shfl_sync(0xffffffffu, Dm[2 * r + 1], 0, 4);
				const float My = __shfl_sync(0xffffffffu, Dm[2 * r], 1, 4), Mxx = __shfl_sync(0xffffffffu, Dm[2 * r + 1], 1, 4);
				const float Mxy = __shfl_sync(0xffffffffu, Dm[2 * r], 2, 4), Myy = __shfl_sync(0xffffffffu, Dm[2 * r + 1], 2, 4);
				const float Sd = __shfl_sync(0xffffffffu, Dy[2 * r], 0, 4), Su = __shfl_sync(0xffffffffu, Dy[2 * r + 1], 0, 4);
				const float4 r0 = *reinterpret_cast<const float4 *>(ent0 + e * TR::kEntryFloats);     // x y a b
				const float2 r1 = *reinterpret_cast<const float2 *>(ent0 + e * TR::kEntryFloats + 4); // c o
				const float ax = r0.x - blk_cx, ay = r0.y - blk_cy; // dx = ax - cx, dy = ay - cy
				const float Sx = ax * M0 - Mx, Sy = ay * M0 - My;
				const float Sxx = ax * (ax * M0 - 2.f * Mx) + Mxx;
				const float Sxy = ax * (ay * M0 - My) - ay * Mx + Mxy;
				const float Syy = ay * (ay * M0 - 2.f * My) + Myy;
				const float o = r1.y;
				// (CR/backward.cu:557-601) dmean2D = -(o s)(dx a + dy b) W/2, -(o s)(dy c + dx b) H/2; dconic = -1/2 (o s){dx dx, dx dy, dy dy}
				float o0, o1;
				if (t == 0) { o0 = -o * (r0.z * Sx + r0.w * Sy) * half_w; o1 = -o * (r1.x * Sy + r0.w * Sx) * half_h; }
				else if (t == 1) { o0 = -0.5f * o * Sxx; o1 = -0.5f * o * Sxy; }
				else if (t == 2) { o0 = -0.5f * o * Syy; o1 = M0; }
				else { o0 = Sd; o1 = Su; }
				if (on) red_add_v2(gacc + (size_t)id * 8 + 2 * t, o0, o1);
			}
		}
		feed.done += m_cur;
		__syncwarp(); // the stage buffer, the ring slots and the s / w rows of this chunk may be reused
		m_cur = m_next;
	}
	feed.drain();
}

static cudaError_t launch_bwd32(int tiles, const uint2 *ranges, const uint32_t *point_list, const uint32_t *__restrict__ header, int W, int H, int tiles_x, const float *rec,
                                const float *features, const float *bg, const float *final_Ts, const uint32_t *n_contrib,
                                const float *dL_dpixels, const float *dL_dpixel_depths, const float *dL_dpixel_uncs, float *gacc,
                                float *dL_dcolors, cudaStream_t stream)
{
	constexpr int smem = kWarpsPerCta * Bwd32Smem::kWarpBytes;
	cudaError_t e = cudaFuncSetAttribute(blend_backward_c32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	if (e != cudaSuccess) return e;
	blend_backward_c32_kernel<<<tiles * kCtasPerTile, 32 * kWarpsPerCta, smem, stream>>>(ranges, point_list, header, W, H, tiles_x, rec, features, bg, final_Ts, n_contrib,
	                                                                                  dL_dpixels, dL_dpixel_depths, dL_dpixel_uncs, gacc, dL_dcolors);
	count_launch();
	return cudaGetLastError();
}

template <int C>
static cudaError_t launch_bwd(int tiles, const uint2 *ranges, const uint32_t *point_list, const uint32_t *__restrict__ header, int W, int H, int tiles_x, const float *rec,
                              const float *features, const float *bg, const float *final_Ts, const uint32_t *n_contrib,
                              const float *dL_dpixels, const float *dL_dpixel_depths, const float *dL_dpixel_uncs, float *gacc,
                              float *dL_dcolors, cudaStream_t stream)
{
	constexpr int smem = kWarpsPerCta * BwdSmem<C>::kWarpBytes;
	// (the attribute is per device and idempotent: set it on every launch rather than cache a per-process flag)
	cudaError_t e = cudaFuncSetAttribute(blend_backward_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	if (e != cudaSuccess) return e;
	blend_backward_kernel<C><<<tiles * kCtasPerTile, 32 * kWarpsPerCta, smem, stream>>>(ranges, point_list, header, W, H, tiles_x, rec, features, bg, final_Ts, n_contrib,
	                                                                                 dL_dpixels, dL_dpixel_depths, dL_dpixel_uncs, gacc, dL_dcolors);
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_blend_backward(int C, int W, int H, const uint2 *ranges, const uint32_t *header, const uint32_t *point_list, const float *rec,
                                  const float *features, const float *bg, const float *final_Ts, const uint32_t *n_contrib,
                                  const float *dL_dpixels, const float *dL_dpixel_depths, const float *dL_dpixel_uncs, float *gacc,
                                  float *dL_dcolors, cudaStream_t stream)
{
	const int tiles_x = (W + GSR_BLOCK_X - 1) / GSR_BLOCK_X, tiles_y = (H + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y;
	const int tiles = tiles_x * tiles_y;
	if (tiles <= 0) return cudaSuccess;
	switch (C) {
	case 3: return launch_bwd<3>(tiles, ranges, point_list, header, W, H, tiles_x, rec, features, bg, final_Ts, n_contrib, dL_dpixels, dL_dpixel_depths, dL_dpixel_uncs, gacc, dL_dcolors, stream);
	case 32: return launch_bwd32(tiles, ranges, point_list, header, W, H, tiles_x, rec, features, bg, final_Ts, n_contrib, dL_dpixels, dL_dpixel_depths, dL_dpixel_uncs, gacc, dL_dcolors, stream);
	default: return cudaErrorInvalidValue;
	}
}

} // namespace gsr
