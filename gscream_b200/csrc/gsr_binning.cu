// gsr_binning.cu — tile binning: depth ordering, instance emission, per-tile stable bucketing, ranges.
//
// Replaces CR/rasterizer_impl.cu:283-324 of the reference (cub InclusiveSum over tiles_touched,
// duplicateWithKeys :70-111, a 6-pass 64-bit cub::DeviceRadixSort over (tile<<32 | depth) keys
// :309-314, identifyTileRanges :116-138) with a two-level scheme that produces the SAME point_list
// and ranges bit for bit while moving ~4x fewer bytes:
//
//   1. sort the P Gaussians once by their 32-bit depth key (stable; culled ones carry 0xFFFFFFFF and
//      end up last).  Ties keep ascending Gaussian index — exactly the tie-break the reference gets
//      from emitting in index order and sorting stably.
//   2. scan tiles_touched in that depth order, emit (tile id, Gaussian id) instances in depth order,
//   3. stable radix sort of the R instances by tile id only (ceil(log2(tiles)) bits, 2 digit passes at
//      1080p instead of 6), which leaves every tile's slab already depth-sorted,
//   4. ranges from the sorted tile ids.
//
// Equivalence: the reference orders instances by (tile, depth bits) with ties broken by emission
// order = Gaussian index (a Gaussian emits each tile at most once).  Step 1+3 yield (tile, depth bits,
// Gaussian index) as well.  Depth keys are positive floats (z > 0.2), so their bit patterns order like
// the values (CR/rasterizer_impl.cu:102-106).
//
// The stable radix sort and the prefix sum are hand-written (gsr_sort.cu; A/B against cub::DeviceRadixSort / DeviceScan in
// round 1: profiles/r1_sort_ab.md, within +-8 %, identical results).
#include "gsr_internal.cuh"
#include "gsr_sort.cuh"
#include <algorithm>

namespace gsr {

// -------- scratch layouts ---------------------------------------------------------------------
static int tile_bits(int W, int H)
{
	const int tiles = ((W + GSR_BLOCK_X - 1) / GSR_BLOCK_X) * ((H + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y);
	int bits = 1;
	while ((1 << bits) < tiles) bits++;
	return bits;
}

// Where the sorted arrays end up ([0] or [1] of the ping-pong pairs): a pure function of the problem shape, so
// forward, backward and the debug export agree without any state.
int depth_order_index()
{
	return radix_plan(1, 32).passes & 1; // 4 passes -> back in [0]
}
int point_list_index(int W, int H)
{
	return radix_plan(1, tile_bits(W, H)).passes & 1;
}

GeomLayout geom_layout(int P)
{
	GeomLayout L;
	size_t off = 0;
	const size_t p = (size_t)std::max(P, 1);
	auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
	L.rec = take(p * GSR_REC_BYTES);
	L.tiles_touched = take(p * 4);
	L.depth_key[0] = take(p * 4);
	L.depth_key[1] = take(p * 4);
	L.depth_val[0] = take(p * 4);
	L.depth_val[1] = take(p * 4);
	L.offsets = take(p * 4);
	L.gacc = take(p * 32);
	L.clamped = take(p * 3);
	L.rgb = take(p * 12);
	L.temp_bytes = radix_plan(P, 32).bytes + scan_scratch_bytes(P);
	L.temp = take(L.temp_bytes);
	L.total = off;
	return L;
}

ImageLayout image_layout(int W, int H)
{
	ImageLayout L;
	size_t off = 0;
	const size_t n = (size_t)std::max(W, 1) * (size_t)std::max(H, 1);
	const size_t tiles = (size_t)((W + GSR_BLOCK_X - 1) / GSR_BLOCK_X) * (size_t)((H + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y);
	auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
	L.final_T = take(n * 4);
	L.n_contrib = take(n * 4);
	L.ranges = take(std::max<size_t>(tiles, 1) * 8);
	L.total = off;
	return L;
}

BinningLayout binning_layout(int P, int64_t R, int W, int H)
{
	(void)P;
	BinningLayout L;
	size_t off = 0;
	const size_t r = (size_t)std::max<int64_t>(R, 1);
	auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
	L.header = take(256);
	L.key[0] = take(r * 4);
	L.key[1] = take(r * 4);
	L.val[0] = take(r * 4);
	L.val[1] = take(r * 4);
	L.temp_bytes = radix_plan(R, tile_bits(W, H)).bytes;
	L.temp = take(L.temp_bytes);
	L.total = off;
	return L;
}

// -------- kernels -----------------------------------------------------------------------------
// Instance emission in depth order (the reference's duplicateWithKeys runs in index order, one thread per Gaussian,
// and writes 64-bit keys; here the depth is already encoded in the position, so the key is just the tile id).
// One lane per INSTANCE: a warp takes 32 consecutive Gaussians of the depth order, whose instances are one contiguous run of
// the output (offsets is the inclusive scan in that order); lane k of each 32-instance step finds its owner Gaussian by a
// 5-step binary search over the warp's prefix sums (shuffles), derives the tile from the owner's rectangle and writes
// (tile id, packed value) to position run_start + k — fully coalesced stores and full lane utilisation whatever the
// rectangle sizes.  History (profiles/r1_sort_ab.md, r1_feed_ab.md): one thread per Gaussian was L1-wavefront bound
// (616 us at R = 40 M); one warp iteration per Gaussian left 3/4 of the lanes idle at 7 instances per Gaussian and doubled
// in cost (77 -> 139 us) once the per-instance warp mask was added.
// The value carries, above the 24-bit Gaussian id, the 8-bit mask of the tile's warps whose 8x4 pixel block the Gaussian's
// alpha >= 1/255 bounding box touches (gsr_blend.cuh).
// A Gaussian covering more than kBigRect tiles (a splat close to the camera plane can cover the whole grid: thousands of
// instances) is not emitted by its warp — 32 such neighbours in the depth order would serialise 10^5 instances in one warp
// (measured: the stage tripled, 0.25 -> 0.75 ms, on views with a few dozen of them, profiles/r2_binning.md) — but queued for
// emit_big_kernel, which spreads each of them over a whole CTA.
constexpr uint32_t kBigRect = 256;

__device__ __forceinline__ void emit_one(uint32_t g, int e, int x0, int y0, int w, float cx, float cy, float hx, float hy, int gx, bool packed,
                                         uint32_t *__restrict__ key, uint32_t *__restrict__ val)
{
	// entry e of the rectangle in row-major order (y outer, x inner), as duplicateWithKeys emits it
	const int ry = e / w, rx = e - ry * w;
	const int tx = x0 + rx, ty = y0 + ry;
	uint32_t v = g;
	if (packed) v |= warp_overlap_mask(cx, cy, hx, hy, (float)(tx * GSR_BLOCK_X), (float)(ty * GSR_BLOCK_Y)) << 24;
	*key = (uint32_t)(ty * gx + tx);
	*val = v;
}

__global__ void __launch_bounds__(256) emit_instances_kernel(int P, const uint32_t *__restrict__ order,
                                                             const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ tiles_touched,
                                                             const float *__restrict__ rec, int gx, int gy, bool packed, int64_t capacity,
                                                             uint32_t *__restrict__ header, uint32_t *__restrict__ big_list,
                                                             uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
	const int lane = threadIdx.x & 31;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t R = __ldg(offsets + (P - 1)); // num_rendered: known on the device only (the host has not waited for it)
	if (i == 0) { // the buffer describes itself: the backward pass and the debug export read the format from here
		header[kHdrPacked] = packed ? 1u : 0u;
		header[kHdrCount] = R;
		header[kHdrOverflow] = (int64_t)R > capacity ? 1u : 0u;
	}
	if ((int64_t)R > capacity) return; // the caller sized the buffer from a guess that was too small: it re-runs this stage
	uint32_t g = 0, tt = 0, start = 0;
	int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
	float2 xy = {0.f, 0.f}, ext = {-1.f, -1.f};
	if (i < P) {
		g = order[i];
		tt = tiles_touched[g];
		start = offsets[i] - tt; // offsets: inclusive scan of tiles_touched in depth order
		if (tt > kBigRect) {
			big_list[atomicAdd(header + kHdrBigCount, 1u)] = (uint32_t)i;
			tt = 0;
		} else if (tt != 0) {
			xy = *reinterpret_cast<const float2 *>(rec + (size_t)g * GSR_REC_FLOATS);
			ext = *reinterpret_cast<const float2 *>(rec + (size_t)g * GSR_REC_FLOATS + 8);
			const int radius = (int)rec[(size_t)g * GSR_REC_FLOATS + 13];
			get_rect(xy.x, xy.y, radius, gx, gy, x0, y0, x1, y1); // same rect as the forward (CR/rasterizer_impl.cu:92)
		}
	}
	// the warp's instances, numbered 0 .. total-1 through an inclusive scan of the lanes' counts
	uint32_t end_rel = tt;
#pragma unroll
	for (int s = 1; s < 32; s <<= 1) {
		const uint32_t o = __shfl_up_sync(0xffffffffu, end_rel, s);
		if (lane >= s) end_rel += o;
	}
	const uint32_t total = __shfl_sync(0xffffffffu, end_rel, 31);
	const int w = x1 - x0;
	for (uint32_t kb = 0; kb < total; kb += 32) {
		const uint32_t k = kb + lane;
		// owner = first lane whose end_rel > k (lanes without instances have end_rel equal to their predecessor's)
		int lo = 0;
#pragma unroll
		for (int step = 16; step >= 1; step >>= 1) {
			const uint32_t probe = __shfl_sync(0xffffffffu, end_rel, lo + step - 1);
			if (probe <= k) lo += step;
		}
		const int src = min(lo, 31);
		const uint32_t s_g = __shfl_sync(0xffffffffu, g, src), s_tt = __shfl_sync(0xffffffffu, tt, src);
		const uint32_t s_end = __shfl_sync(0xffffffffu, end_rel, src), s_start = __shfl_sync(0xffffffffu, start, src);
		const int s_x0 = __shfl_sync(0xffffffffu, x0, src), s_y0 = __shfl_sync(0xffffffffu, y0, src), s_w = __shfl_sync(0xffffffffu, w, src);
		const float s_cx = __shfl_sync(0xffffffffu, xy.x, src), s_cy = __shfl_sync(0xffffffffu, xy.y, src);
		const float s_hx = __shfl_sync(0xffffffffu, ext.x, src), s_hy = __shfl_sync(0xffffffffu, ext.y, src);
		if (k < total) {
			const int e = (int)(k - (s_end - s_tt));
			emit_one(s_g, e, s_x0, s_y0, s_w, s_cx, s_cy, s_hx, s_hy, gx, packed, keys + s_start + e, vals + s_start + e);
		}
	}
}

// the queued large rectangles: one CTA per Gaussian (round robin over a fixed grid), one thread per instance
__global__ void __launch_bounds__(256) emit_big_kernel(int P, const uint32_t *__restrict__ order, const uint32_t *__restrict__ offsets,
                                                       const uint32_t *__restrict__ tiles_touched, const float *__restrict__ rec, int gx, int gy,
                                                       bool packed, int64_t capacity, const uint32_t *__restrict__ header,
                                                       const uint32_t *__restrict__ big_list, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
	if ((int64_t)__ldg(offsets + (P - 1)) > capacity) return;
	const uint32_t count = header[kHdrBigCount];
	for (uint32_t b = blockIdx.x; b < count; b += gridDim.x) {
		const uint32_t i = big_list[b], g = order[i], tt = tiles_touched[g], start = offsets[i] - tt;
		const float2 xy = *reinterpret_cast<const float2 *>(rec + (size_t)g * GSR_REC_FLOATS);
		const float2 ext = *reinterpret_cast<const float2 *>(rec + (size_t)g * GSR_REC_FLOATS + 8);
		const int radius = (int)rec[(size_t)g * GSR_REC_FLOATS + 13];
		int x0, y0, x1, y1;
		get_rect(xy.x, xy.y, radius, gx, gy, x0, y0, x1, y1);
		for (uint32_t e = threadIdx.x; e < tt; e += blockDim.x)
			emit_one(g, (int)e, x0, y0, x1 - x0, xy.x, xy.y, ext.x, ext.y, gx, packed, keys + start + e, vals + start + e);
	}
}

// identifyTileRanges, CR/rasterizer_impl.cu:116-138, on 32-bit tile ids; 8 sorted keys per thread (two 16-B loads).
__global__ void __launch_bounds__(256) tile_ranges_kernel(int64_t capacity, const uint32_t *__restrict__ n_dev, const uint32_t *__restrict__ keys,
                                                          uint2 *__restrict__ ranges)
{
	if ((int64_t)__ldg(n_dev) > capacity) return; // overflow: every range stays (0, 0)
	const int64_t L = (int64_t)__ldg(n_dev);
	const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
	if (base >= L) return;
	uint32_t k[8];
	if (base + 8 <= L) {
		const uint4 a = __ldg(reinterpret_cast<const uint4 *>(keys + base)), b = __ldg(reinterpret_cast<const uint4 *>(keys + base + 4));
		k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
	} else {
#pragma unroll
		for (int j = 0; j < 8; j++) k[j] = base + j < L ? __ldg(keys + base + j) : 0u;
	}
	uint32_t prev = base > 0 ? __ldg(keys + base - 1) : 0u;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		const int64_t idx = base + j;
		if (idx >= L) break;
		const uint32_t cur = k[j];
		if (idx == 0)
			ranges[cur].x = 0;
		else if (cur != prev) {
			ranges[prev].y = (uint32_t)idx;
			ranges[cur].x = (uint32_t)idx;
		}
		if (idx == L - 1) ranges[cur].y = (uint32_t)L;
		prev = cur;
	}
}

// -------- host orchestration --------------------------------------------------------------------
// Stage-1 tail: stable depth order of the Gaussians + inclusive scan of tiles_touched in that order.
// offsets[P-1] is then R (num_rendered).
cudaError_t depth_order_and_scan(int P, char *geom, const GeomLayout &L, cudaStream_t stream)
{
	if (P <= 0) return cudaSuccess;
	cudaError_t e;
	const int which = radix_sort_pairs((uint32_t *)(geom + L.depth_key[0]), (uint32_t *)(geom + L.depth_val[0]), (uint32_t *)(geom + L.depth_key[1]),
	                                   (uint32_t *)(geom + L.depth_val[1]), P, nullptr, 32, geom + L.temp, stream, &e);
	if (e != cudaSuccess) return e;
	const uint32_t *order = (const uint32_t *)(geom + L.depth_val[which]);
	return inclusive_sum_gather((const uint32_t *)(geom + L.tiles_touched), order, (uint32_t *)(geom + L.offsets), P,
	                            geom + L.temp + radix_plan(P, 32).bytes, stream);
}

// Stage-2 head: emission, stable bucketing by tile id, ranges.  `capacity` is what the binning buffer was sized for; the
// number of instances itself (num_rendered = offsets[P-1]) is read by the kernels from device memory, so nothing here waits
// for the host to learn it.  If it exceeds the capacity the stage leaves every range empty and flags the buffer's header.
cudaError_t bin_instances(int P, int64_t capacity, int W, int H, char *geom, const GeomLayout &GL,
                          char *binning, const BinningLayout &BL, char *image, const ImageLayout &IL, cudaStream_t stream)
{
	const int gx = (W + GSR_BLOCK_X - 1) / GSR_BLOCK_X, gy = (H + GSR_BLOCK_Y - 1) / GSR_BLOCK_Y;
	const int tiles = gx * gy;
	cudaError_t e = cudaMemsetAsync(image + IL.ranges, 0, (size_t)tiles * sizeof(uint2), stream); // CR/rasterizer_impl.cu:316
	if (e != cudaSuccess) return e;
	count_launch();
	if (P <= 0 || capacity <= 0) return cudaSuccess;

	const uint32_t *order = (const uint32_t *)(geom + GL.depth_val[depth_order_index()]);
	const uint32_t *n_dev = (const uint32_t *)(geom + GL.offsets) + (P - 1);
	uint32_t *header = (uint32_t *)(binning + BL.header);
	uint32_t *big_list = (uint32_t *)(geom + GL.depth_key[1 - depth_order_index()]); // the depth sort's spare key array: free by now
	if ((e = cudaMemsetAsync(header, 0, 256, stream)) != cudaSuccess) return e;
	emit_instances_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, order, (const uint32_t *)(geom + GL.offsets),
	                                                         (const uint32_t *)(geom + GL.tiles_touched), (const float *)(geom + GL.rec),
	                                                         gx, gy, point_list_packed(P), capacity, header, big_list,
	                                                         (uint32_t *)(binning + BL.key[0]), (uint32_t *)(binning + BL.val[0]));
	emit_big_kernel<<<4 * 148, 256, 0, stream>>>(P, order, (const uint32_t *)(geom + GL.offsets), (const uint32_t *)(geom + GL.tiles_touched),
	                                            (const float *)(geom + GL.rec), gx, gy, point_list_packed(P), capacity, header, big_list,
	                                            (uint32_t *)(binning + BL.key[0]), (uint32_t *)(binning + BL.val[0]));
	count_launch(3);
	e = cudaGetLastError();
	if (e != cudaSuccess) return e;

	const int bits = tile_bits(W, H);
	const int which = radix_sort_pairs((uint32_t *)(binning + BL.key[0]), (uint32_t *)(binning + BL.val[0]), (uint32_t *)(binning + BL.key[1]),
	                                   (uint32_t *)(binning + BL.val[1]), capacity, n_dev, bits, binning + BL.temp, stream, &e);
	if (e != cudaSuccess) return e;
	tile_ranges_kernel<<<(unsigned)((capacity + 2047) / 2048), 256, 0, stream>>>(capacity, n_dev, (const uint32_t *)(binning + BL.key[which]), (uint2 *)(image + IL.ranges));
	count_launch();
	return cudaGetLastError();
}

// Largest instance count whose layout fits `bytes`: forward, backward and the debug export all derive the buffer's layout
// from its size, so no capacity has to travel beside it.
int64_t binning_capacity(int P, int W, int H, size_t bytes)
{
	if (binning_layout(P, 1, W, H).total > bytes) return 0;
	int64_t lo = 1, hi = (int64_t)(bytes / 16) + 1; // layout(lo) fits, layout(hi) does not (16 B per instance alone exceed it)
	while (hi - lo > 1) {
		const int64_t mid = lo + (hi - lo) / 2;
		if (binning_layout(P, mid, W, H).total <= bytes) lo = mid; else hi = mid;
	}
	return lo;
}

} // namespace gsr
