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
//     therefore works on a landed 16-entry chunk of its feed (gsr_blend.cuh) in phases:
//       phase 0 (C = 32): dot[p][e] = g_p . f_e for the whole chunk as one [32 x 32] x [32 x 16] product on the tensor pipe;
//       phase 1 (lane = pixel, sequential in depth): the recurrence, branch-free, four entries side by side; leaves s and w
//                of every (entry, pixel) of the chunk in 4 KB of shared memory.  Nothing in it waits for a reduction;
//       phase 2 (entries independent): every per-Gaussian sum over the warp's 32 pixels as tensor-pipe products of the s / w
//                tiles with operands that are fixed per warp — the gradient block (colour sums), the depth / uncertainty
//                gradients, and the monomials of the block-centred pixel coordinates (the geometric terms as moments) —
//                ending in red.global.add.v2.f32 straight from the accumulator fragments, 10 per chunk, where the reference
//                issues 32 x (C + 8) scalar atomics per (warp, Gaussian).
//     Details and the measurements that led here (round 1: everything per entry on the FP32 pipe, a 9-shuffle butterfly on
//     the recurrence's critical path; then the shared-memory data-pipe finding) are with the kernel below and in
//     profiles/r2_blend_mma.md.
//   * Same warp-private feed as the forward kernel (per-warp list scan by the instance masks, private double-buffered
//     cp.async gather, no block barrier), run back to front and started at the warp's own deepest last contributor.
//
// Scalar terms are accumulated into gacc[P][8] = {dmean2D.x, dmean2D.y, dconic.x, dconic.y, dconic.w,
// dopacity, ddepth, duncertainty}; colours into dL_dcolors[P][C].  Both must be zero (or hold the
// running sum) on entry.
#include "gsr_blend.cuh"
#include "gsr_internal.cuh"
#include <type_traits>

namespace gsr {

constexpr int kWarpsPerCta = GSR_BWD_WARPS_PER_CTA;
constexpr int kCtasPerTile = kWarpsPerTile / kWarpsPerCta;

__device__ __forceinline__ float exp2_approx(float x)
{
	float r;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
// 1 / d for d in [0.01, 1]: MUFU.RCP + one Newton step, no range check (__frcp_rn's slow path is a call behind a branch)
__device__ __forceinline__ float rcp_1ulp(float d)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
	const float e = __fmaf_rn(-d, r, 1.f);
	return __fmaf_rn(r, e, r);
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
template <int C>
struct BwdSmem {
	using TR = BlendTraits<C>;
	static constexpr int kS = TR::kWarpBytes, kW = kS + kChunk * 32 * 4, kDot = kW + kChunk * 32 * 4;
	static constexpr int kWarpBytes = kDot + (C == 32 ? 32 * kDotStride * 4 : 0);
};
// resident CTAs (of kWarpsPerCta warps) per SM asked of ptxas.  C = 32: 164 registers (the gradient block twice, 72 registers of
// operands) -> 12 warps per SM; 144 / 128 registers spill and are 13 % / 60 % slower (profiles/r2_blend_mma.md)
#ifndef GSR_BWD32_MINCTAS
#define GSR_BWD32_MINCTAS 6
#endif
#ifndef GSR_BWD3_MINCTAS
#define GSR_BWD3_MINCTAS 10
#endif

template <int C>
__global__ void __launch_bounds__(32 * kWarpsPerCta, C == 32 ? GSR_BWD32_MINCTAS : GSR_BWD3_MINCTAS) blend_backward_kernel(
    const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, const uint32_t *__restrict__ header, int W, int H, int tiles_x,
    const float *__restrict__ rec, const float *__restrict__ features, const float *__restrict__ bg,
    const float *__restrict__ final_Ts, const uint32_t *__restrict__ n_contrib,
    const float *__restrict__ dL_dpixels, const float *__restrict__ dL_dpixel_depths, const float *__restrict__ dL_dpixel_uncs,
    float *__restrict__ gacc, float *__restrict__ dL_dcolors)
{
	constexpr bool kWide = (C == 32); // feature rows beside the record, dot products and colour sums on the tensor pipe
	static_assert(kWide || C == 3, "instantiated for the two channel counts of the C ABI");
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
	// ... and the shallowest one (0 when a pixel of the block received nothing at all)
	int warp_min_last = last_contributor;
#pragma unroll
	for (int s = 16; s >= 1; s >>= 1) warp_min_last = min(warp_min_last, __shfl_xor_sync(0xffffffffu, warp_min_last, s));

	// pixel p of the block (p = 8 row + col) -> offset in a plane, or -1 outside the image
	auto pix_off = [&](int p) -> long long {
		const int x = x0 + (p & 7), y = y0 + (p >> 3);
		return (x < W && y < H) ? (long long)W * y + x : -1;
	};
	// the block's 32 x 32 upstream colour gradient, as the A operand of the dot product (rows = pixels, k = channels) ...
	float gA[kWide ? 2 : 1][4][4];
	if (kWide) {
#pragma unroll
		for (int mt = 0; mt < 2; mt++)
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const long long off = pix_off(16 * mt + gq + 8 * (j & 1));
#pragma unroll
				for (int ks = 0; ks < 4; ks++) gA[mt % (kWide ? 2 : 1)][ks][j] = off >= 0 ? __ldg(dL_dpixels + (size_t)kperm(ks, t, j >> 1) * plane + off) : 0.f;
			}
	}
	// ... and as the B operand of the colour sums (k = pixels, columns = channels): channel 8 nt + gq, pixels 16 kk + 4 t .. + 3
	float gB[4][kWide ? 4 : 1][2];
	if (kWide) {
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
				gB[2 * kk][nt % (kWide ? 4 : 1)][0] = v[0]; gB[2 * kk][nt % (kWide ? 4 : 1)][1] = v[1];
				gB[2 * kk + 1][nt % (kWide ? 4 : 1)][0] = v[2]; gB[2 * kk + 1][nt % (kWide ? 4 : 1)][1] = v[3];
			}
	}
	// one more 8-column tile of that operand: the depth and uncertainty gradients (columns gq = 0, 1) and, for C = 3, the three
	// colour gradients (columns 2..4); the other columns are zero
	float gY[4][2];
	{
		const float *col = gq == 0 ? dL_dpixel_depths : gq == 1 ? dL_dpixel_uncs : (!kWide && gq < 2 + C) ? dL_dpixels + (size_t)(gq - 2) * plane : nullptr;
#pragma unroll
		for (int ks = 0; ks < 4; ks++)
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const long long off = pix_off(kperm(ks, t, h));
				gY[ks][h] = (col != nullptr && off >= 0) ? __ldg(col + off) : 0.f;
			}
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
	float grow[kWide ? 1 : C]; // C = 3: this pixel's colour gradient row (the dot product is three FFMA)
	float bg_dot = 0.f;
	if (!kWide) {
#pragma unroll
		for (int ch = 0; ch < (kWide ? 0 : C); ch++) {
			grow[ch] = inside ? dL_dpixels[ch * plane + pix_id] : 0.f;
			bg_dot += bg[ch] * grow[ch];
		}
	} else {
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

	unsigned char *warp_smem = smem_raw + (size_t)lwarp * BwdSmem<C>::kWarpBytes;
	float *s_s = reinterpret_cast<float *>(warp_smem + BwdSmem<C>::kS);   // [kChunk][32]
	float *s_w = reinterpret_cast<float *>(warp_smem + BwdSmem<C>::kW);   // [kChunk][32]
	float *s_dot = reinterpret_cast<float *>(warp_smem + BwdSmem<C>::kDot); // [32][kDotStride]

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

		// ---- phase 0 (C = 32): dot[p][e] for the whole chunk on the tensor pipe (rows past m_cur: stale operands, results unused) ----
		if (kWide) {
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
						for (int j = 0; j < 4; j++) split_tf32(gA[mt % (kWide ? 2 : 1)][2 * kk + sub][j], ah[j], al[j]);
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
			__syncwarp();
		}

		// ---- phase 1 (lane = pixel): the recurrence over the chunk's entries, in depth order ----
		// kSub32 entries at a time: (a) everything that does not depend on the pixel's running state — alpha, 1 / (1 - alpha), G —
		// then (b) the short carried chain (T, X).  An entry that does not touch the pixel is encoded as alpha = 0, rinv = 1,
		// G = 0: the recurrence then leaves T unchanged, hands X on unchanged (0 * dot + 1 * X') and produces s = w = 0.
		bool live = false; // some pixel of the warp received gradient from some entry of the chunk
		// positions run downwards: once the chunk's first entry lies below every pixel's last contributor, no entry of the
		// chunk needs the per-pixel position test any more (the common case after the first chunks)
		const bool below_all = (int)feed.q_pos[feed.done & (kRing - 1)] < warp_min_last;
		auto sub_batch = [&](int e0, auto check_pos) {
			float dcol[4] = {0.f, 0.f, 0.f, 0.f};
			if (kWide) {
				const float4 d4 = *reinterpret_cast<const float4 *>(s_dot + lane * kDotStride + e0);
				dcol[0] = d4.x; dcol[1] = d4.y; dcol[2] = d4.z; dcol[3] = d4.w;
			}
			float al[kSub32], ri[kSub32], Gs[kSub32], dt[kSub32];
			bool any_valid = false, near = false;
#pragma unroll
			for (int b = 0; b < kSub32; b++) {
				const int e = e0 + b;
				const float *ent = ent0 + e * TR::kEntryFloats;
				const float4 r0 = *reinterpret_cast<const float4 *>(ent);     // x y a b
				const float4 r1 = *reinterpret_cast<const float4 *>(ent + 4); // c o depth unc
				const float dx = r0.x - pixf_x, dy = r0.y - pixf_y;
				const float power = gaussian_power(r0.z, r0.w, r1.x, dx, dy);
				// G by ex2.approx (2 instructions instead of expf's 8; ~3e-7 relative apart).  The alpha >= 1/255 decision must be the
				// forward's, which used expf: pairs within 4e-9 of the threshold are redone exactly below.
				const float G = exp2_approx(power * 1.4426950408889634f);
				const float alpha = min(0.99f, __fmul_rn(r1.y, G));
				near = near || fabsf(alpha - kAlphaMin) < 4e-9f;
				bool valid = (e < m_cur) && !(power > 0.0f) && !(alpha < kAlphaMin);
				if (decltype(check_pos)::value) valid = valid && (int)feed.q_pos[(feed.done + e) & (kRing - 1)] < last_contributor; // 0-based list position
				else valid = valid && last_contributor > 0;
				al[b] = valid ? alpha : 0.f;
				ri[b] = valid ? rcp_1ulp(__fsub_rn(1.f, alpha)) : 1.f; // T / (1 - alpha) (CR/backward.cu:533) as T * rcp; also serves the background term
				Gs[b] = valid ? G : 0.f;
				float dot = r1.z * gd + r1.w * gu; // f_j . g_p over colour channels, depth and uncertainty
				if (kWide) {
					dot += dcol[b];
				} else { // the colours ride in the record (slots 10..12)
					const float4 r2 = *reinterpret_cast<const float4 *>(ent + 8);
					dot += r2.z * grow[0] + r2.w * grow[1 % (kWide ? 1 : C)] + ent[12] * grow[2 % (kWide ? 1 : C)];
				}
				dt[b] = valid ? dot : 0.f;
				any_valid = any_valid || valid;
			}
			if (__any_sync(0xffffffffu, near)) { // rare: some pair sits at the 1/255 threshold — redo the sub-batch exactly (expf)
				any_valid = false;
#pragma unroll 1
				for (int b = 0; b < kSub32; b++) {
					const int e = e0 + b;
					const float *ent = ent0 + e * TR::kEntryFloats;
					const float4 r0 = *reinterpret_cast<const float4 *>(ent);
					const float4 r1 = *reinterpret_cast<const float4 *>(ent + 4);
					const float power = gaussian_power(r0.z, r0.w, r1.x, r0.x - pixf_x, r0.y - pixf_y);
					const float G = expf(power);
					const float alpha = min(0.99f, __fmul_rn(r1.y, G));
					const bool valid = (e < m_cur) && ((int)feed.q_pos[(feed.done + e) & (kRing - 1)] < last_contributor) && !(power > 0.0f) && !(alpha < kAlphaMin);
					float dot = r1.z * gd + r1.w * gu;
					if (kWide) dot += s_dot[lane * kDotStride + e];
					else {
						const float4 r2 = *reinterpret_cast<const float4 *>(ent + 8);
						dot += r2.z * grow[0] + r2.w * grow[1 % (kWide ? 1 : C)] + ent[12] * grow[2 % (kWide ? 1 : C)];
					}
					// (static indexing keeps the four arrays in registers)
#pragma unroll
					for (int bb = 0; bb < kSub32; bb++)
						if (bb == b) {
							al[bb] = valid ? alpha : 0.f;
							ri[bb] = valid ? rcp_1ulp(__fsub_rn(1.f, alpha)) : 1.f;
							Gs[bb] = valid ? G : 0.f;
							dt[bb] = valid ? dot : 0.f;
						}
					any_valid = any_valid || valid;
				}
			}
			if (__any_sync(0xffffffffu, any_valid)) live = true;
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
		};
		if (below_all) {
			for (int e0 = 0; e0 < m_cur; e0 += kSub32) sub_batch(e0, std::false_type{});
		} else {
			for (int e0 = 0; e0 < m_cur; e0 += kSub32) sub_batch(e0, std::true_type{});
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
					if (kWide) {
#pragma unroll
						for (int nt = 0; nt < 4; nt++) {
							uint32_t b0h, b0l, b1h, b1l;
							split_tf32(gB[ks][nt % (kWide ? 4 : 1)][0], b0h, b0l);
							split_tf32(gB[ks][nt % (kWide ? 4 : 1)][1], b1h, b1l);
							mma3_tf32(Dc[nt], ah, al, b0h, b1h, b0l, b1l);
						}
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
				const bool on = e < m_cur; // (an entry none of the block's pixels touched adds zeros)
				const uint32_t id = feed.q_id[(feed.done + e) & (kRing - 1)];
				if (on && kWide) {
#pragma unroll
					for (int nt = 0; nt < 4; nt++) red_add_v2(dL_dcolors + (size_t)id * C + 8 * nt + 2 * t, Dc[nt][2 * r], Dc[nt][2 * r + 1]);
				}
				if (on && !kWide) { // columns 2..4 of the last tile: lane t = 1 holds channels 0, 1 and lane t = 2 channel 2
					if (t == 1) {
						red_add(dL_dcolors + (size_t)id * C + 0, Dy[2 * r]);
						red_add(dL_dcolors + (size_t)id * C + 1, Dy[2 * r + 1]);
					} else if (t == 2) {
						red_add(dL_dcolors + (size_t)id * C + 2, Dy[2 * r]);
					}
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

template <int C>
static cudaError_t launch_bwd(int tiles, const uint2 *ranges, const uint32_t *point_list, const uint32_t *header, int W, int H, int tiles_x, const float *rec,
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
	case 32: return launch_bwd<32>(tiles, ranges, point_list, header, W, H, tiles_x, rec, features, bg, final_Ts, n_contrib, dL_dpixels, dL_dpixel_depths, dL_dpixel_uncs, gacc, dL_dcolors, stream);
	default: return cudaErrorInvalidValue;
	}
}

} // namespace gsr
