// gsr_blend_bwd_mma.cu — backward tile blend, C = 32, with the two per-(warp, chunk) 32x32 products on the tensor pipe.
//
// Same semantics, traversal and feed as blend_backward_kernel<32> (gsr_blend_bwd.cu; reference: renderCUDA backward,
// CR/backward.cu:409-604).  What changes is where the two dense contractions of the per-pixel -> per-Gaussian
// scatter are evaluated.  For one warp (32 pixels p, gradient block g[p][ch]) and one chunk of 16 staged Gaussians j:
//     dots[j][p]   = sum_ch f[j][ch] * g[p][ch]          (needed by the per-pixel recurrence: dL/dalpha)
//     dcol[j][ch]  = sum_p  w[j][p]  * g[p][ch]          (the colour gradient; w = alpha * T from the recurrence)
// are a [16x32]x[32x32] GEMM each.  The scalar kernel spends 32 FFMA + 8 LDS.128 per (pixel lane, Gaussian) on each of
// them — 87 of its ~236 instructions per contributing (warp, Gaussian) pair — and keeps the gradient block twice in
// registers (row per lane + column per lane, 66 registers), which pins it at 128 registers / 16 warps per SM.
// Here both run as mma.sync.m16n8k8 TF32 with the 3xTF32 split (x = hi + lo, a*b ~ a_lo*b_hi + a_hi*b_lo + a_hi*b_hi,
// fp32 accumulate; relative error ~2^-21 per product, below the 1e-5 parity tolerance — fp32 FFMA is ~3.9x slower per
// MAC than this path on B200 even after the 3x split, profiles/r1_mma_probe.md), the gradient block lives once in
// shared memory (4 KB per warp, XOR-swizzled so that both fragment shapes read it conflict-free), and the dots / weights
// pass between the pixel-lane recurrence and the fragments through a 16x32 shared tile.
// tcgen05 is not usable here: each warp multiplies its own private operands (its pixel block's gradients, its own
// compacted Gaussian list); there is no CTA-wide 64/128-row tile to hand to the single-thread UMMA issue model.
//
// Fragment maps (g = lane >> 2, t = lane & 3; PTX m16n8k8 .tf32 row.col):
//   dots : A = f   rows j = g, g+8          k-step kt, col t / t+4  <-> ch 8t+2kt / 8t+2kt+1   (lane reads f[j][8t..8t+7])
//          B = g^T col n = g of tile nt     <-> pixel 8(g>>1)+2nt+(g&1)                        (lane reads g[p][8t..8t+7])
//          C       cols 2t, 2t+1 of tile nt <-> pixels 8t+2nt, 8t+2nt+1  -> tile[j][8t+2nt..+1]
//   dcol : A = w   rows j = g, g+8          k-step kt, col t / t+4  <-> pixel 8t+2kt / 8t+2kt+1 (lane reads tile[j][8t..8t+7])
//          B = g   col n = g of tile nt     <-> ch 4g+nt                                       (lane reads g[p][4g..4g+3])
//          C       cols 2t, 2t+1 of tile nt <-> ch 8t+nt, 8t+4+nt -> two red.global.add.v4.f32 per row
#include "gsr_blend.cuh"
#include "gsr_internal.cuh"
#include "gsr_tf32.cuh"

namespace gsr {

namespace {

constexpr int kWarpsPerCta = GSR_BWD_WARPS_PER_CTA;
constexpr int kCtasPerTile = kWarpsPerTile / kWarpsPerCta;
constexpr int kC = 32;
constexpr int kTileStride = 36;                       // floats per row of the 16x32 dots / weights tile
constexpr int kMmaWarpBytes = BlendTraits<kC>::kWarpBytes + kChunk * kTileStride * 4 + 32 * 32 * 4;
#ifndef GSR_BWD_MMA_MINWARPS
#define GSR_BWD_MMA_MINWARPS 16
#endif

// gradient block g[p][ch] in shared memory: row p = 32 floats, 16-B units XOR-swizzled by the row
__device__ __forceinline__ int sg_swz(int p) { return (((p >> 3) & 3) << 1) | (p & 1); }
__device__ __forceinline__ int sg_index(int p, int ch) { return p * 32 + ((((ch >> 2) ^ sg_swz(p)) << 2) | (ch & 3)); }

// Transpose-reduce of 8 values per lane (see gsr_blend_bwd.cu): afterwards v[0] on lane l is the warp total of value vidx8(l).
__device__ __forceinline__ void warp_transpose_reduce8(float (&v)[8], int lane)
{
	int s = 16;
#pragma unroll
	for (int n = 4; n >= 1; n >>= 1, s >>= 1) {
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

__global__ void __launch_bounds__(32 * kWarpsPerCta, GSR_BWD_MMA_MINWARPS / kWarpsPerCta) blend_backward_mma_kernel(
    const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, int packed, int W, int H, int tiles_x,
    const float *__restrict__ rec, const float *__restrict__ features, const float *__restrict__ bg,
    const float *__restrict__ final_Ts, const uint32_t *__restrict__ n_contrib,
    const float *__restrict__ dL_dpixels, const float *__restrict__ dL_dpixel_depths, const float *__restrict__ dL_dpixel_uncs,
    float *__restrict__ gacc, float *__restrict__ dL_dcolors)
{
	using TR = BlendTraits<kC>;
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
	const size_t plane = (size_t)H * W;
	const size_t pix_id = (size_t)W * py + px;

	const uint2 range = ranges[tile];
	const float T_final = inside ? final_Ts[pix_id] : 0.f;
	const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;
	int warp_last = last_contributor;
#pragma unroll
	for (int s = 16; s >= 1; s >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, s));
	warp_last = min(warp_last, (int)(range.y - range.x));
	if (warp_last == 0) return; // warps are independent: no barrier follows

	unsigned char *wsm = smem_raw + (size_t)lwarp * kMmaWarpBytes;
	float *s_tile = reinterpret_cast<float *>(wsm + TR::kWarpBytes);                 // [16][36]: dots, then alpha*T
	float *s_g = s_tile + kChunk * kTileStride;                                      // [32][32] swizzled gradient block

	// this warp's 32x32 upstream-gradient block -> shared memory (lane = pixel); depth / uncertainty terms stay scalar
	float gd = 0.f, gu = 0.f, bg_dot = 0.f;
#pragma unroll 8
	for (int ch = 0; ch < kC; ch++) {
		const float v = inside ? __ldg(dL_dpixels + ch * plane + pix_id) : 0.f;
		bg_dot = fmaf(__ldg(bg + ch), v, bg_dot);
		s_g[sg_index(lane, ch)] = v;
	}
	if (inside) {
		gd = dL_dpixel_depths[pix_id];
		gu = dL_dpixel_uncs[pix_id];
	}

	float T = T_final;
	float X = 0.f, last_alpha = 0.f, last_dot = 0.f;
	const float ddelx_dx = 0.5 * W, ddely_dy = 0.5 * H;
	const float neg_Tfinal_bg = -T_final * bg_dot;
	const int fg = lane >> 2, ft = lane & 3; // fragment coordinates

	// B fragments of the dots product, chunk invariant: g[p][8t..8t+7] for the four pixels p = 8(g>>1) + 2nt + (g&1), raw fp32
	// (split into TF32 hi / lo at use: 32 registers instead of 64)
	float gfrag[4][8];
	__syncwarp();
#pragma unroll
	for (int nt = 0; nt < 4; nt++) {
		const int p = 8 * (fg >> 1) + 2 * nt + (fg & 1);
		const int sw = sg_swz(p);
		const float4 g0 = *reinterpret_cast<const float4 *>(s_g + p * 32 + (((2 * ft) ^ sw) << 2));
		const float4 g1 = *reinterpret_cast<const float4 *>(s_g + p * 32 + (((2 * ft + 1) ^ sw) << 2));
		gfrag[nt][0] = g0.x; gfrag[nt][1] = g0.y; gfrag[nt][2] = g0.z; gfrag[nt][3] = g0.w;
		gfrag[nt][4] = g1.x; gfrag[nt][5] = g1.y; gfrag[nt][6] = g1.z; gfrag[nt][7] = g1.w;
	}

	WarpFeed<kC, true> feed;
	feed.init(wsm, point_list + range.x, warp_last, rec, features, warp, lane, packed != 0);
	feed.fill();
	int m_cur = feed.issue(0);
	for (int chunk = 0; m_cur > 0; chunk++) {
		feed.fill();
		const int m_next = feed.issue((chunk + 1) & 1);
		cp_async_wait_but_one();
		__syncwarp(); // this chunk has landed (and s_g is visible on the first pass)
		const float *ent0 = feed.stage + (chunk & 1) * TR::kStageFloats;

		// ---- dots[j][p] = f[j] . g[p] over the 32 colour channels -------------------------------------------------
		{
			float c[4][4];
#pragma unroll
			for (int nt = 0; nt < 4; nt++)
#pragma unroll
				for (int i = 0; i < 4; i++) c[nt][i] = 0.f;
			const float *ra = ent0 + fg * TR::kEntryFloats + TR::kRecParts * 4 + 8 * ft; // f[g][8t..], f[g+8][8t..]
			const float *rb = ra + 8 * TR::kEntryFloats;
#pragma unroll
			for (int kt = 0; kt < 4; kt++) {
				const float2 fa = *reinterpret_cast<const float2 *>(ra + 2 * kt), fb = *reinterpret_cast<const float2 *>(rb + 2 * kt);
				uint32_t a0h, a0l, a1h, a1l, a2h, a2l, a3h, a3l;
				tf32_split(fa.x, a0h, a0l);
				tf32_split(fb.x, a1h, a1l);
				tf32_split(fa.y, a2h, a2l);
				tf32_split(fb.y, a3h, a3l);
#pragma unroll
				for (int nt = 0; nt < 4; nt++) {
					uint32_t b0h, b0l, b1h, b1l;
					tf32_split_trunc(gfrag[nt][2 * kt], b0h, b0l);
					tf32_split_trunc(gfrag[nt][2 * kt + 1], b1h, b1l);
					mma_tf32(c[nt], a0l, a1l, a2l, a3l, b0h, b1h);
					mma_tf32(c[nt], a0h, a1h, a2h, a3h, b0l, b1l);
					mma_tf32(c[nt], a0h, a1h, a2h, a3h, b0h, b1h);
				}
			}
#pragma unroll
			for (int nt = 0; nt < 4; nt++) {
				*reinterpret_cast<float2 *>(s_tile + fg * kTileStride + 8 * ft + 2 * nt) = make_float2(c[nt][0], c[nt][1]);
				*reinterpret_cast<float2 *>(s_tile + (fg + 8) * kTileStride + 8 * ft + 2 * nt) = make_float2(c[nt][2], c[nt][3]);
			}
		}
		__syncwarp();

		// ---- per-pixel recurrence, back to front (lane = pixel) ---------------------------------------------------------
		uint32_t anymask = 0;
		const float *ent = ent0;
		for (int e = 0; e < m_cur; e++, ent += TR::kEntryFloats) {
			const uint32_t slot = (feed.done + e) & (kRing - 1);
			const int pos = (int)feed.q_pos[slot];
			const float4 r0 = *reinterpret_cast<const float4 *>(ent);     // x y a b
			const float4 r1 = *reinterpret_cast<const float4 *>(ent + 4); // c o depth unc
			const float2 d = {r0.x - pixf_x, r0.y - pixf_y};
			const float power = gaussian_power(r0.z, r0.w, r1.x, d.x, d.y);
			const bool maybe = (pos < last_contributor) && !(power > 0.0f);
			const float G = expf(power);
			const float alpha = min(0.99f, __fmul_rn(r1.y, G));
			const bool valid = maybe && !(alpha < kAlphaMin);
			float *cell = s_tile + e * kTileStride + lane; // dots[e][lane] in, alpha*T out
			if (!__any_sync(0xffffffffu, valid)) {
				*cell = 0.f;
				continue;
			}
			anymask |= 1u << e;
			float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
			float w = 0.f;
			if (valid) {
				const float rinv = __frcp_rn(__fsub_rn(1.f, alpha)); // T <- T / (1 - alpha), CR/backward.cu:533
				T = T * rinv;
				w = alpha * T;
				const float dot = fmaf(r1.w, gu, fmaf(r1.z, gd, *cell));
				X = last_alpha * last_dot + (1.f - last_alpha) * X;
				last_dot = dot;
				float dL_dalpha = (dot - X) * T;
				last_alpha = alpha;
				dL_dalpha += neg_Tfinal_bg * rinv;
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
			}
			*cell = w;
			warp_transpose_reduce8(v, lane);
			if ((lane & 3) == 0) {
				const int q = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
				red_add(gacc + (size_t)feed.q_id[slot] * 8 + q, v[0]);
			}
		}
		for (int e = m_cur; e < kChunk; e++) s_tile[e * kTileStride + lane] = 0.f; // padding rows carry no weight
		__syncwarp();

		// ---- dcol[j][ch] = sum_p w[j][p] g[p][ch] ---------------------------------------------------------------------
		if (anymask) {
			uint32_t ahi[8], alo[8], bhi_[8], blo_[8]; // rows g and g+8 of w, pixels 8t..8t+7, split
			{
				const float *ra = s_tile + fg * kTileStride + 8 * ft;
				const float *rb = ra + 8 * kTileStride;
				const float4 a0 = *reinterpret_cast<const float4 *>(ra), a1 = *reinterpret_cast<const float4 *>(ra + 4);
				const float4 b0 = *reinterpret_cast<const float4 *>(rb), b1 = *reinterpret_cast<const float4 *>(rb + 4);
				const float fa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
				const float fb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
				for (int i = 0; i < 8; i++) {
					tf32_split(fa[i], ahi[i], alo[i]);
					tf32_split(fb[i], bhi_[i], blo_[i]);
				}
			}
			float c[4][4];
#pragma unroll
			for (int nt = 0; nt < 4; nt++)
#pragma unroll
				for (int i = 0; i < 4; i++) c[nt][i] = 0.f;
#pragma unroll
			for (int kt = 0; kt < 4; kt++) {
				const int p0 = 8 * ft + 2 * kt;                          // swizzle of rows p0 / p0+1: 2t / 2t+1
				const float4 g0 = *reinterpret_cast<const float4 *>(s_g + p0 * 32 + ((fg ^ (2 * ft)) << 2));
				const float4 g1 = *reinterpret_cast<const float4 *>(s_g + (p0 + 1) * 32 + ((fg ^ (2 * ft + 1)) << 2));
				const float gv0[4] = {g0.x, g0.y, g0.z, g0.w}, gv1[4] = {g1.x, g1.y, g1.z, g1.w};
#pragma unroll
				for (int nt = 0; nt < 4; nt++) {
					uint32_t h0, l0, h1, l1;
					tf32_split_trunc(gv0[nt], h0, l0);
					tf32_split_trunc(gv1[nt], h1, l1);
					mma_tf32(c[nt], alo[2 * kt], blo_[2 * kt], alo[2 * kt + 1], blo_[2 * kt + 1], h0, h1);
					mma_tf32(c[nt], ahi[2 * kt], bhi_[2 * kt], ahi[2 * kt + 1], bhi_[2 * kt + 1], l0, l1);
					mma_tf32(c[nt], ahi[2 * kt], bhi_[2 * kt], ahi[2 * kt + 1], bhi_[2 * kt + 1], h0, h1);
				}
			}
			// rows fg and fg+8 of the chunk; lane holds channels 8t..8t+3 (c[nt][0]) and 8t+4..8t+7 (c[nt][1])
			if ((anymask >> fg) & 1u) {
				float *dst = dL_dcolors + (size_t)feed.q_id[(feed.done + fg) & (kRing - 1)] * kC + 8 * ft;
				red_add_v4(dst, c[0][0], c[1][0], c[2][0], c[3][0]);
				red_add_v4(dst + 4, c[0][1], c[1][1], c[2][1], c[3][1]);
			}
			if ((anymask >> (fg + 8)) & 1u) {
				float *dst = dL_dcolors + (size_t)feed.q_id[(feed.done + fg + 8) & (kRing - 1)] * kC + 8 * ft;
				red_add_v4(dst, c[0][2], c[1][2], c[2][2], c[3][2]);
				red_add_v4(dst + 4, c[0][3], c[1][3], c[2][3], c[3][3]);
			}
		}
		feed.done += m_cur;
		__syncwarp(); // stage buffer, ring slots and the tile may be reused
		m_cur = m_next;
	}
	cp_async_wait_all();
}

} // namespace

cudaError_t launch_blend_backward_mma(int P, int W, int H, const uint2 *ranges, const uint32_t *point_list, const float *rec,
                                      const float *features, const float *bg, const float *final_Ts, const uint32_t *n_contrib,
                                      const float *dL_dpixels, const float *dL_dpixel_depths, const float *dL_dpixel_uncs, float *gacc,
                                      float *dL_dcolors, cudaStream_t stream)
{
	const int tiles_x = (W + GSR_BLOCK_X - 1) / GSR_BLOCK_X, tiles_y = (H + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y;
	const int tiles = tiles_x * tiles_y;
	if (tiles <= 0) return cudaSuccess;
	const size_t smem = (size_t)kWarpsPerCta * kMmaWarpBytes;
	static bool configured = false;
	if (!configured) {
		cudaError_t e = cudaFuncSetAttribute(blend_backward_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return e;
		configured = true;
	}
	blend_backward_mma_kernel<<<tiles * kCtasPerTile, 32 * kWarpsPerCta, smem, stream>>>(
	    ranges, point_list, point_list_packed(P) ? 1 : 0, W, H, tiles_x, rec, features, bg, final_Ts, n_contrib, dL_dpixels, dL_dpixel_depths,
	    dL_dpixel_uncs, gacc, dL_dcolors);
	count_launch();
	return cudaGetLastError();
}

} // namespace gsr
