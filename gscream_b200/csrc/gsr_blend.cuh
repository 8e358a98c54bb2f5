// gsr_blend.cuh — pieces shared by the forward and backward tile-blend kernels.
//
// Staging scheme (v2, "warp-private feeds").  A 16x16 tile is blended by 8 warps, warp w owning an 8x4 pixel block.
// Each instance of the tile's depth-sorted list carries, in the top byte of its point_list entry, an 8-bit mask of the
// warps whose pixel block its alpha >= 1/255 bounding box touches (written by the instance emission, gsr_binning.cu).
// Every warp runs its own pipeline over the list, with no block-level barrier anywhere:
//   scan    : 64 list entries per step (two coalesced 4-B loads per lane, prefetched two steps ahead), ballot-compacted
//             into the warp's ring of (Gaussian id, list position) — only entries whose mask names this warp;
//   gather  : the next 16 ring entries' projected records (and feature rows) are copied into one of the warp's two
//             private stage buffers with 16-B cp.async (LDGSTS), one chunk ahead of the arithmetic;
//   blend   : the landed 16-entry chunk is processed in phases: the per-pixel recurrence (lane = pixel, entries in order) on the
//             FP32 pipe, the dense per-(block, chunk) products at C = 32 as 3xTF32 mma.sync on the tensor pipe (helpers below;
//             gsr_blend_fwd.cu / gsr_blend_bwd.cu).
// v1 staged every instance of the tile once per CTA behind a per-round __syncthreads(); ncu showed 18-19 % of the warp
// samples waiting at that barrier for the slowest warp of the round (profiles/r1_blend_v3_summary.md).  v2 trades ~1.4x more
// L2->SM gather traffic (an instance is fetched by each warp that needs it) for fully decoupled warps, a per-warp early
// exit in the forward and a per-warp start position in the backward.  A/B: profiles/r1_feed_ab.md.
#pragma once
#include "gsr_common.cuh"

namespace gsr {

constexpr int kWarpsPerTile = 8;     // warp w of a tile owns an 8x4 pixel block of the 16x16 tile
// Warps never synchronise with each other, so the CTA is only a residency unit: a tile is blended by 8 / WPC CTAs of WPC
// warps.  Small CTAs return their SM slots as soon as their own warps are done instead of waiting for the slowest warp
// of the tile.  Measured (profiles/r1_feed_ab.md): the forward kernel is fastest with one warp per CTA (32 resident CTAs
// per SM), the backward kernel with two.
#ifndef GSR_FWD_WARPS_PER_CTA
#define GSR_FWD_WARPS_PER_CTA 1
#endif
#ifndef GSR_BWD_WARPS_PER_CTA
#define GSR_BWD_WARPS_PER_CTA 2
#endif
static_assert(kWarpsPerTile % GSR_FWD_WARPS_PER_CTA == 0 && kWarpsPerTile % GSR_BWD_WARPS_PER_CTA == 0, "warps per CTA must divide 8");
constexpr int kChunk = 16;           // ring entries gathered / blended per pipeline step
constexpr int kScan = 64;            // list entries scanned per refill (two per lane)
constexpr int kRing = 128;           // ring capacity (>= kScan + 2 * kChunk + kChunk)
static_assert((kRing & (kRing - 1)) == 0 && kRing >= kScan + 3 * kChunk, "ring: power of two, holds one scan step plus the queued and in-flight chunks");
constexpr uint32_t kIdMask = 0x00FFFFFFu;  // low 24 bits of a packed point_list entry: the Gaussian id
constexpr float kAlphaMin = 1.0f / 255.0f;

template <int C>
struct BlendTraits {
	// C <= 3: colours ride in the 64-B record (slots 10..12), the whole record is staged.  Otherwise the blend loops need
	// only the first 32 B of the record (x y a b | c opacity depth uncertainty) plus the C*4-B row of colors_precomp.
	static constexpr bool kFeatInRec = (C <= 3);
	static constexpr int kRecParts = kFeatInRec ? 4 : 2;      // 16-B parts
	static constexpr int kFeatParts = kFeatInRec ? 0 : C / 4;
	static constexpr int kParts = kRecParts + kFeatParts;
	static constexpr int kEntryFloats = kParts * 4;
	static constexpr int kStageFloats = kChunk * kEntryFloats;
	// per warp: two stage buffers, then the ring (ids, positions)
	static constexpr int kWarpBytes = 2 * kStageFloats * 4 + kRing * 8;
};

// 16-byte asynchronous global->shared copy (cp.async, SASS LDGSTS; L2 only, no L1 allocation).
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// One warp's feed over its tile's list.  kReverse: scan back to front starting at list position n-1 (backward pass).
// (A per-entry cp.async.bulk variant of the feature gather was measured in round 1 and lost: profiles/r1_feed_ab.md section 6.)
// All members are warp-uniform except `lane`-dependent temporaries.
template <int C, bool kReverse>
struct WarpFeed {
	using TR = BlendTraits<C>;
	const uint32_t *list;      // point_list + range.x
	const float *rec, *feat;
	float *stage;              // [2][kChunk][kEntryFloats]
	uint32_t *q_id, *q_pos;    // [kRing]
	int n, next_scan;          // list positions to scan; next scan step (in units of kScan)
	uint32_t tail, issued, done;
	uint32_t pre0a, pre0b, pre1a, pre1b;  // entries of the next two scan steps (two per lane each)
	int warp, lane;
	bool packed;

	__device__ __forceinline__ int pos_of(int ordinal) const { return kReverse ? n - 1 - ordinal : ordinal; }
	// plain (not read-only-path) load: the forward kernel clears mask bits of this list while it runs (see blend_forward_kernel)
	__device__ __forceinline__ uint32_t load_entry(int ordinal) const { return ordinal < n ? list[pos_of(ordinal)] : 0u; }

	__device__ __forceinline__ void init(unsigned char *warp_smem, const uint32_t *list_, int n_, const float *rec_, const float *feat_,
	                                     int warp_, int lane_, bool packed_)
	{
		stage = reinterpret_cast<float *>(warp_smem);
		q_id = reinterpret_cast<uint32_t *>(warp_smem + 2 * TR::kStageFloats * 4);
		q_pos = q_id + kRing;
		list = list_; n = n_; rec = rec_; feat = feat_; warp = warp_; lane = lane_; packed = packed_;
		next_scan = 0;
		tail = issued = done = 0;
		pre0a = load_entry(lane); pre0b = load_entry(32 + lane);
		pre1a = load_entry(kScan + lane); pre1b = load_entry(kScan + 32 + lane);
	}
	__device__ __forceinline__ bool exhausted() const { return next_scan * kScan >= n; }

	// consume one prefetched scan step into the ring, start the load of the step after the next
	__device__ __forceinline__ void refill()
	{
		const int base = next_scan * kScan;
#pragma unroll
		for (int c = 0; c < 2; c++) {
			const uint32_t v = c == 0 ? pre0a : pre0b;
			const int ordinal = base + 32 * c + lane;
			const bool hit = ordinal < n && (!packed || ((v >> (24 + warp)) & 1u));
			const uint32_t ball = __ballot_sync(0xffffffffu, hit);
			if (hit) {
				const uint32_t idx = (tail + __popc(ball & ((1u << lane) - 1u))) & (kRing - 1);
				q_id[idx] = packed ? (v & kIdMask) : v;
				q_pos[idx] = (uint32_t)pos_of(ordinal);
			}
			tail += __popc(ball);
		}
		pre0a = pre1a; pre0b = pre1b;
		next_scan++;
		pre1a = load_entry((next_scan + 1) * kScan + lane);
		pre1b = load_entry((next_scan + 1) * kScan + 32 + lane);
	}
	// keep two chunks queued ahead of the gather whenever the list allows
	__device__ __forceinline__ void fill()
	{
		while ((int)(tail - issued) < 2 * kChunk && !exhausted()) refill();
		__syncwarp();
	}
	// gather the next (up to) kChunk ring entries into stage buffer `s`; always commits one cp.async group.
	// Two uniform passes (records, then feature rows) with power-of-two lane -> (entry, 16-B part) maps, so that a round is
	// LDS id / LEA / LDGSTS: a single mixed pass cost 32 instructions per round in index and base-pointer selection
	// (12.7 % of the forward kernel's instructions, profiles/r1_blend_v5_summary.md).
	__device__ __forceinline__ int issue(int s)
	{
		const int m = min(kChunk, (int)(tail - issued));
		float *dst = stage + s * TR::kStageFloats;
		{
			constexpr int kPer = 32 / TR::kRecParts; // entries per round
			const int le = lane / TR::kRecParts, part = lane % TR::kRecParts;
#pragma unroll
			for (int r = 0; r < kChunk / kPer; r++) {
				const int e = r * kPer + le;
				if (e < m)
					cp_async16(dst + e * TR::kEntryFloats + part * 4, rec + (size_t)q_id[(issued + e) & (kRing - 1)] * GSR_REC_FLOATS + part * 4);
			}
		}
		if constexpr (!TR::kFeatInRec) {
			constexpr int kPer = 32 / TR::kFeatParts;
			static_assert(32 % TR::kFeatParts == 0, "feature row must split into a power-of-two number of 16-B parts");
			const int le = lane / TR::kFeatParts, part = lane % TR::kFeatParts;
#pragma unroll
			for (int r = 0; r < kChunk / kPer; r++) {
				const int e = r * kPer + le;
				if (e < m)
					cp_async16(dst + e * TR::kEntryFloats + (TR::kRecParts + part) * 4, feat + (size_t)q_id[(issued + e) & (kRing - 1)] * C + part * 4);
			}
		}
		cp_async_commit();
		issued += m;
		return m;
	}
	// the older of the (at most two) chunks in flight has landed for this lane
	__device__ __forceinline__ void wait() { cp_async_wait_but_one(); }
	// before the warp retires: nothing may still be in flight into its shared memory
	__device__ __forceinline__ void drain() { cp_async_wait_all(); }
};

// ---- 3xTF32 on the legacy tensor path (mma.sync m16n8k8, SASS HMMA.1688.F32.TF32) ----------------------------------------
// The dense per-(pixel block, chunk) products of the C = 32 blend kernels run here: x = hi + lo with hi = the top 19 bits of x
// (what the tensor core reads of a 32-bit operand), lo = x - hi exact; hi*hi + lo*hi + hi*lo leaves ~2^-21 relative error per
// product, far inside the 1e-5 parity gate (tests/test_gpu_parity.py).  tcgen05 is no option for a single warp with private
// operands: it needs a CTA-wide 64/128-row tile in shared memory issued by one thread.
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
	// (volatile: splits of loop-invariant register operands must stay inside the chunk loop — hoisted, they would double the
	// registers those operands take)
	asm volatile("and.b32 %0, %1, 0xffffe000;" : "=r"(hi) : "r"(__float_as_uint(x)));
	lo = __float_as_uint(x - __uint_as_float(hi));
}
// D += A B, m16n8k8; with g = lane >> 2, t = lane & 3:  A row-major a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);
// B b0 (k = t, n = g) b1 (k = t+4, n = g);  C/D c0 c1 (g, 2t), (g, 2t+1), c2 c3 (g+8, 2t), (g+8, 2t+1)
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
	asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
	             : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
	             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma3_tf32(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t b0h, uint32_t b1h, uint32_t b0l, uint32_t b1l)
{
	mma_tf32(c, al, b0h, b1h);
	mma_tf32(c, ah, b0l, b1l);
	mma_tf32(c, ah, b0h, b1h);
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

} // namespace gsr
