// gsr_common.cuh — private definitions shared by the sm_100a kernels of libgsr_b200.so.
//
// Vocabulary follows the reference (W-Ted/GScream, submodules/diff-gaussian-rasterization):
// P Gaussians, R = num_rendered tile instances, 16x16-pixel tiles, per-tile ranges into the
// depth-sorted point_list.  CR/ = the reference's cuda_rasterizer/ directory.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#define GSR_BLOCK_X 16   // CR/config.h:16
#define GSR_BLOCK_Y 16   // CR/config.h:17
#define GSR_TILE_PIX 256

// Per-Gaussian projected record written by the preprocess kernel and gathered (bulk-async)
// by the blend kernels.  16 floats = 64 B so that one cp.async.bulk moves one Gaussian.
//   [0] x  [1] y  [2] conic.a  [3] conic.b | [4] conic.c  [5] opacity  [6] depth  [7] uncertainty
//   [8] hx [9] hy (half extents of the alpha >= 1/255 ellipse's bounding box, conservative)
//   [10..12] rgb (only when C <= 3: colours ride in the record)   [13] radius (integral)
//   [14] pow_min: lower bound on `power` below which alpha < 1/255 for sure  [15] spare
#define GSR_REC_FLOATS 16
#define GSR_REC_BYTES 64

namespace gsr {

// ---- private scratch layouts (opaque to callers; sized by gsr_*_bytes) -------------------
constexpr size_t kAlign = 256;
__host__ __device__ inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct GeomLayout {     // "geomBuffer"
	size_t rec;          // float[P][16]
	size_t tiles_touched;// u32[P]
	size_t depth_key[2]; // u32[P] x2: [0] unsorted depth keys, [1] sorted
	size_t depth_val[2]; // u32[P] x2: [0] iota, [1] Gaussian indices in depth order
	size_t offsets;      // u32[P]   inclusive scan of tiles_touched in depth order
	size_t gacc;         // float[P][8] backward accumulators (see gsr_preprocess.cu)
	size_t clamped;      // u8[3P]   SH clamp flags (CR/forward.cu:69-71)
	size_t rgb;          // float[3P] SH-evaluated colours
	size_t temp;         // CUB temp (sort over P and scan over P)
	size_t temp_bytes;
	size_t total;
};
struct ImageLayout {    // "imgBuffer"
	size_t final_T;      // float[N]
	size_t n_contrib;    // u32[N]
	size_t ranges;       // uint2[tiles]
	size_t total;
};
struct BinningLayout {  // "binningBuffer"; R below is the CAPACITY the buffer was sized for (>= num_rendered)
	size_t header;       // u32[64]: the buffer describes itself (kHdr*), written by the instance emission
	size_t key[2];       // u32[R] x2 tile ids: [0] emitted, [1] sorted
	size_t val[2];       // u32[R] x2 Gaussian ids: [0] emitted, [1] sorted = point_list
	size_t temp;
	size_t temp_bytes;
	size_t total;
};

// header words of the binning buffer
constexpr int kHdrPacked = 0;    // 1: point_list entries carry the per-warp overlap mask above a 24-bit id, 0: plain ids
constexpr int kHdrCount = 1;     // num_rendered as the device computed it
constexpr int kHdrOverflow = 2;  // 1: num_rendered exceeded the buffer's capacity, the lists are empty
constexpr int kHdrBigCount = 3;  // Gaussians whose rectangle was queued for the CTA-wide emission (gsr_binning.cu)

GeomLayout geom_layout(int P);
ImageLayout image_layout(int W, int H);
BinningLayout binning_layout(int P, int64_t R, int W, int H);
int64_t binning_capacity(int P, int W, int H, size_t bytes); // largest R whose layout fits `bytes`

// ---- launch accounting ------------------------------------------------------------------
void count_launch(int n = 1);

// ---- device helpers ---------------------------------------------------------------------
// Tile rectangle of a projected Gaussian — same arithmetic as CR/auxiliary.h:46-56 (getRect).
__device__ __forceinline__ void get_rect(float px, float py, int max_radius, int gx, int gy,
                                         int &x0, int &y0, int &x1, int &y1)
{
	x0 = min(gx, max(0, (int)((px - max_radius) / GSR_BLOCK_X)));
	y0 = min(gy, max(0, (int)((py - max_radius) / GSR_BLOCK_Y)));
	x1 = min(gx, max(0, (int)((px + max_radius + GSR_BLOCK_X - 1) / GSR_BLOCK_X)));
	y1 = min(gy, max(0, (int)((py + max_radius + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y)));
}

// One bit per warp of a tile's CTA (warp w owns the 8x4 pixel block at x = 8 (w & 1), y = 4 (w >> 1)): does the bounding
// box {|x - cx| <= hx, |y - cy| <= hy} of the Gaussian's alpha >= 1/255 region touch that block?  (Pixel centres are
// integer coordinates, CR/forward.cu:466.)  hx < 0 encodes "never contributes"; +inf encodes "never cull".
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

// point_list entries carry that mask in their top byte when every Gaussian id fits 24 bits (P <= 2^24); above that the
// entries are plain ids and the blend kernels treat every instance as a candidate for every warp (still exact, slower).
// (Host-side decision; gsr_debug_plain_point_list(1) forces the plain form so that tests can cover it at small P.)
bool point_list_packed(int P);

// mbarrier + bulk-async (TMA, non-tensor form: SASS UBLKCP) wrappers, sm_90+/sm_100a.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-B aligned; completes on `bar`.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

// 16-byte vector reduction to global memory (sm_90+): one RED for four consecutive floats.
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b)
{
	asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add(float *addr, float a)
{
	asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

} // namespace gsr
