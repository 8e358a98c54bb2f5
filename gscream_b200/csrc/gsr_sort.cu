// gsr_sort.cu — implementation of gsr_sort.cuh (stable LSD radix sort of u32 pairs + gathered inclusive scan).
#include "gsr_sort.cuh"
#include <algorithm>

namespace gsr {

// ------------------------------------------------------------------------------------------------
// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix, *total = sum
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *s_warp /*[8]*/, uint32_t *total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int s = 1; s < 32; s <<= 1) {
		const uint32_t o = __shfl_up_sync(0xffffffffu, inc, s);
		if (lane >= s) inc += o;
	}
	if (lane == 31) s_warp[warp] = inc;
	__syncthreads();
	uint32_t warp_base = 0, tot = 0;
#pragma unroll
	for (int w = 0; w < 8; w++) {
		const uint32_t c = s_warp[w];
		if (w < warp) warp_base += c;
		tot += c;
	}
	__syncthreads(); // s_warp may be reused by the caller's next scan
	if (total) *total = tot;
	return warp_base + inc - v;
}

// ------------------------------------------------------------------------------------------------
// per-tile digit histogram of one pass (tiles are those of the CURRENT element order, so this runs per pass)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRsThreads) rs_histogram_kernel(const uint32_t *__restrict__ keys, int64_t n_cap, const uint32_t *__restrict__ n_dev, int shift, uint32_t mask, int tiles,
                                                                  uint32_t *__restrict__ counts, uint32_t *__restrict__ totals)
{
	const int64_t n = device_count(n_cap, n_dev);
	__shared__ uint32_t s_hist[256];
	const int t = threadIdx.x;
	s_hist[t] = 0;
	__syncthreads();
	const int64_t base = (int64_t)blockIdx.x * kRsTile;
	uint32_t key[kRsItems];
#pragma unroll
	for (int i = 0; i < kRsItems; i++) { // all 16 loads in flight before the first shared-memory atomic
		const int64_t idx = base + i * kRsThreads + t;
		key[i] = idx < n ? __ldg(keys + idx) : 0u;
	}
#pragma unroll
	for (int i = 0; i < kRsItems; i++)
		if (base + i * kRsThreads + t < n) atomicAdd(&s_hist[(key[i] >> shift) & mask], 1u);
	__syncthreads();
	const uint32_t c = s_hist[t];
	counts[(size_t)t * tiles + blockIdx.x] = c;
	// digit totals: by atomics for small inputs; above kRsTotalsByRows tiles they serialise at L2 (thousands of adds per
	// address: 50 us at 10^4 tiles) and rs_totals_kernel sums the rows instead
	if (totals && c) atomicAdd(&totals[t], c);
}

// one CTA per digit: totals[digit] = sum over tiles of counts[digit][tile]  (large inputs only)
__global__ void __launch_bounds__(kRsThreads) rs_totals_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ totals, int tiles)
{
	__shared__ uint32_t s_warp[8];
	const uint32_t *row = counts + (size_t)blockIdx.x * tiles;
	uint32_t sum = 0;
	for (int i = threadIdx.x; i < tiles; i += kRsThreads) sum += row[i];
	uint32_t tot;
	block_exclusive_scan_256(sum, s_warp, &tot);
	if (threadIdx.x == 0) totals[blockIdx.x] = tot;
}

// one CTA per digit: counts[digit][tile] <- global base of the digit + exclusive prefix over tiles
__global__ void __launch_bounds__(kRsThreads) rs_offsets_kernel(uint32_t *__restrict__ counts, const uint32_t *__restrict__ totals, int tiles)
{
	__shared__ uint32_t s_warp[8];
	const int d = blockIdx.x, t = threadIdx.x;
	if (totals[d] == 0) return; // no tile holds this digit: its row is never read
	uint32_t base;
	block_exclusive_scan_256(t < d ? totals[t] : 0u, s_warp, &base);
	uint32_t *row = counts + (size_t)d * tiles;
	uint32_t carry = base;
	// four consecutive tiles per thread and round: a quarter of the block scans (each costs two barriers)
	for (int c0 = 0; c0 < tiles; c0 += 4 * kRsThreads) {
		const int i = c0 + 4 * t;
		uint32_t v[4];
#pragma unroll
		for (int j = 0; j < 4; j++) v[j] = i + j < tiles ? row[i + j] : 0u;
		uint32_t tot;
		const uint32_t ex = block_exclusive_scan_256(v[0] + v[1] + v[2] + v[3], s_warp, &tot);
		uint32_t run = carry + ex;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			if (i + j < tiles) row[i + j] = run;
			run += v[j];
		}
		carry += tot;
	}
}

// resident CTAs per SM asked of ptxas: two (no spill with the kept ranks); three spill 72 B and are 10 % slower
#ifndef GSR_RS_MINCTAS
#define GSR_RS_MINCTAS 2
#endif
// stable scatter of one digit pass
__global__ void __launch_bounds__(kRsThreads, GSR_RS_MINCTAS) rs_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n_cap,
                                                                const uint32_t *__restrict__ n_dev, int shift, int width,
                                                                const uint32_t *__restrict__ offsets, int tiles)
{
	const int64_t n = device_count(n_cap, n_dev);
	if ((int64_t)blockIdx.x * kRsTile >= n) return; // (uniform per block) nothing of this tile exists
	const uint32_t mask = (1u << width) - 1u;
	__shared__ uint32_t s_hist[8][256]; // per-warp digit counts, then per-warp base inside the tile
	__shared__ uint32_t s_tile_start[256];
	__shared__ uint32_t s_goff[256];
	__shared__ uint32_t s_warp[8];
	__shared__ uint32_t s_keys[kRsTile];
	__shared__ uint32_t s_vals[kRsTile];

	const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const int64_t tile_base = (int64_t)blockIdx.x * kRsTile;
#pragma unroll
	for (int w = 0; w < 8; w++) s_hist[w][t] = 0;
	__syncthreads();

	// each warp owns the contiguous run [tile_base + warp*512, +512): element (i, lane) = run[i*32 + lane]
	uint32_t key[kRsItems], val[kRsItems];
	const int64_t run_base = tile_base + warp * (kRsItems * 32);
#pragma unroll
	for (int i = 0; i < kRsItems; i++) {
		const int64_t idx = run_base + i * 32 + lane;
		const bool valid = idx < n;
		key[i] = valid ? __ldg(keys_in + idx) : 0xFFFFFFFFu;
		val[i] = valid ? __ldg(vals_in + idx) : 0u;
	}
	// lanes holding the same digit, by `width` ballots.  (The hardware MATCH.ANY iterates once per distinct value — up to
	// 32 rounds per call on these keys — and made this kernel 10x slower, profiles/r1_sort_ab.md.)
	auto peers_of = [&](uint32_t d, bool valid) -> uint32_t {
		uint32_t m = __ballot_sync(0xffffffffu, valid);
		for (int b = 0; b < width; b++) {
			const bool bit = (d >> b) & 1u;
			const uint32_t bal = __ballot_sync(0xffffffffu, bit);
			m &= bit ? bal : ~bal;
		}
		return m;
	};
	// pass A: this warp's digit counts (the lowest lane of every peer group adds the group size).  Each element's rank inside
	// its peer group and the group's size are kept (two 16-bit fields per register pair of elements) so that pass B does not
	// have to rebuild the peer masks: `width` ballots per element once instead of twice.
	uint32_t rank_cnt[kRsItems / 2];
#pragma unroll
	for (int i = 0; i < kRsItems / 2; i++) rank_cnt[i] = 0;
#pragma unroll
	for (int i = 0; i < kRsItems; i++) {
		const bool valid = run_base + i * 32 + lane < n;
		const uint32_t d = (key[i] >> shift) & mask;
		const uint32_t m = peers_of(d, valid);
		const uint32_t r = __popc(m & lt_mask), c = __popc(m);
		if (valid && r == 0) s_hist[warp][d] += c; // one writer per (warp, digit) per step
		rank_cnt[i >> 1] |= (r | (c << 8)) << (16 * (i & 1));
		__syncwarp();
	}
	__syncthreads();

	// digit t: prefix over the 8 warps (warp order == input order), then over digits inside the tile
	uint32_t run = 0;
#pragma unroll
	for (int w = 0; w < 8; w++) {
		const uint32_t c = s_hist[w][t];
		s_hist[w][t] = run;
		run += c;
	}
	const uint32_t tile_start = block_exclusive_scan_256(run, s_warp, nullptr);
	s_tile_start[t] = tile_start;
	s_goff[t] = run ? offsets[(size_t)t * tiles + blockIdx.x] : 0u;
#pragma unroll
	for (int w = 0; w < 8; w++) s_hist[w][t] += tile_start; // s_hist[w][d] = next free slot of (warp w, digit d) in the tile
	__syncthreads();

	// pass B: the same peer groups, now handing out slots in input order; reorder through shared memory so that equal
	// digits are contiguous
#pragma unroll
	for (int i = 0; i < kRsItems; i++) {
		const bool valid = run_base + i * 32 + lane < n;
		const uint32_t d = (key[i] >> shift) & mask;
		const uint32_t rc = (rank_cnt[i >> 1] >> (16 * (i & 1))) & 0xFFFFu, r = rc & 0xFFu, c = rc >> 8;
		uint32_t slot = 0;
		if (valid) slot = s_hist[warp][d];
		__syncwarp();
		if (valid) {
			if (r == 0) s_hist[warp][d] = slot + c;
			s_keys[slot + r] = key[i];
			s_vals[slot + r] = val[i];
		}
		__syncwarp();
	}
	__syncthreads();
	const int valid_count = (int)min((int64_t)kRsTile, n - tile_base);
#pragma unroll 4
	for (int i = 0; i < kRsItems; i++) {
		const int p = i * kRsThreads + t;
		if (p < valid_count) {
			const uint32_t k = s_keys[p];
			const uint32_t d = (k >> shift) & mask;
			const uint32_t dst = s_goff[d] + ((uint32_t)p - s_tile_start[d]);
			keys_out[dst] = k;
			vals_out[dst] = s_vals[p];
		}
	}
}

RadixPlan radix_plan(int64_t n, int bits)
{
	RadixPlan pl;
	pl.passes = std::max(1, (bits + 7) / 8);
	pl.tiles = (int)std::max<int64_t>(1, (n + kRsTile - 1) / kRsTile);
	pl.counts_off = 0;                                       // u32[256][tiles], reused by every pass
	pl.totals_off = align_up((size_t)256 * pl.tiles * 4);    // u32[passes][256]
	pl.bytes = pl.totals_off + align_up((size_t)pl.passes * 256 * 4);
	return pl;
}

int radix_sort_pairs(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n, const uint32_t *n_dev, int bits,
                     void *scratch, cudaStream_t stream, cudaError_t *err)
{
	*err = cudaSuccess;
	if (n <= 0) return 0;
	bits = std::min(std::max(bits, 1), 32);
	const RadixPlan pl = radix_plan(n, bits);
	uint32_t *counts = (uint32_t *)((char *)scratch + pl.counts_off);
	uint32_t *totals = (uint32_t *)((char *)scratch + pl.totals_off);
	if ((*err = cudaMemsetAsync(totals, 0, (size_t)pl.passes * 256 * 4, stream)) != cudaSuccess) return 0;
	count_launch();
	uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
	for (int p = 0; p < pl.passes; p++) {
		const int width = std::min(8, bits - 8 * p);
		const uint32_t mask = (1u << width) - 1u;
		const bool by_rows = pl.tiles > kRsTotalsByRows;
		rs_histogram_kernel<<<pl.tiles, kRsThreads, 0, stream>>>(kin, n, n_dev, 8 * p, mask, pl.tiles, counts, by_rows ? nullptr : totals + p * 256);
		if (by_rows) {
			rs_totals_kernel<<<256, kRsThreads, 0, stream>>>(counts, totals + p * 256, pl.tiles);
			count_launch();
		}
		rs_offsets_kernel<<<256, kRsThreads, 0, stream>>>(counts, totals + p * 256, pl.tiles);
		rs_scatter_kernel<<<pl.tiles, kRsThreads, 0, stream>>>(kin, vin, kout, vout, n, n_dev, 8 * p, width, counts, pl.tiles);
		count_launch(3);
		std::swap(kin, kout);
		std::swap(vin, vout);
	}
	*err = cudaGetLastError();
	return pl.passes & 1;
}

// ------------------------------------------------------------------------------------------------
// inclusive sum of in[order[i]]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRsThreads) scan_local_kernel(const uint32_t *__restrict__ in, const uint8_t *__restrict__ in8, const uint32_t *__restrict__ order,
                                                                uint32_t *__restrict__ out, int64_t n, uint32_t *__restrict__ tile_sums)
{
	__shared__ uint32_t s_warp[8];
	const int t = threadIdx.x;
	const int64_t base = (int64_t)blockIdx.x * kRsTile + (int64_t)t * kRsItems; // blocked: 16 consecutive items per thread
	uint32_t v[kRsItems];
	uint32_t sum = 0;
#pragma unroll
	for (int i = 0; i < kRsItems; i++) {
		const int64_t idx = base + i;
		uint32_t x = 0;
		if (idx < n) x = in8 ? (__ldg(in8 + idx) ? 1u : 0u) : (order ? __ldg(in + __ldg(order + idx)) : __ldg(in + idx));
		sum += x;
		v[i] = sum;
	}
	uint32_t tot;
	const uint32_t ex = block_exclusive_scan_256(sum, s_warp, &tot);
#pragma unroll
	for (int i = 0; i < kRsItems; i++) {
		const int64_t idx = base + i;
		if (idx < n) out[idx] = ex + v[i];
	}
	if (t == 0) tile_sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(kRsThreads) scan_tiles_kernel(uint32_t *tile_sums, int tiles)
{
	__shared__ uint32_t s_warp[8];
	uint32_t carry = 0;
	for (int c0 = 0; c0 < tiles; c0 += kRsThreads) {
		const int i = c0 + threadIdx.x;
		const uint32_t v = i < tiles ? tile_sums[i] : 0u;
		uint32_t tot;
		const uint32_t ex = block_exclusive_scan_256(v, s_warp, &tot);
		if (i < tiles) tile_sums[i] = carry + ex; // exclusive prefix of the tile
		carry += tot;
	}
}
// Adds the tile's exclusive prefix.  raw: tile_sums still holds the tile TOTALS (at most kRsThreads tiles) and every CTA sums
// the ones before it itself — one launch less than scanning them first.  mask8 / ids: compaction of the set entries.
__global__ void __launch_bounds__(kRsThreads) scan_finish_kernel(uint32_t *__restrict__ out, int64_t n, const uint32_t *__restrict__ tile_sums, int raw,
                                                                 const uint8_t *__restrict__ mask8, uint32_t *__restrict__ ids)
{
	__shared__ uint32_t s_warp[8];
	uint32_t add;
	if (raw)
		block_exclusive_scan_256((int)threadIdx.x < (int)blockIdx.x ? tile_sums[threadIdx.x] : 0u, s_warp, &add);
	else
		add = tile_sums[blockIdx.x];
	if (add == 0 && !ids) return;
	const int64_t base = (int64_t)blockIdx.x * kRsTile;
#pragma unroll 4
	for (int i = 0; i < kRsItems; i++) {
		const int64_t idx = base + i * kRsThreads + threadIdx.x;
		if (idx < n) {
			const uint32_t o = out[idx] + add;
			if (add) out[idx] = o;
			if (ids && __ldg(mask8 + idx)) ids[o - 1] = (uint32_t)idx;
		}
	}
}

size_t scan_scratch_bytes(int64_t n) { return align_up((size_t)std::max<int64_t>(1, (n + kRsTile - 1) / kRsTile) * 4); }

static cudaError_t inclusive_sum_any(const uint32_t *in, const uint8_t *in8, const uint32_t *order, uint32_t *out, uint32_t *ids, int64_t n, void *scratch,
                                     cudaStream_t stream)
{
	if (n <= 0) return cudaSuccess;
	const int tiles = (int)((n + kRsTile - 1) / kRsTile);
	uint32_t *tile_sums = (uint32_t *)scratch;
	scan_local_kernel<<<tiles, kRsThreads, 0, stream>>>(in, in8, order, out, n, tile_sums);
	count_launch();
	if (tiles > 1 || ids) {
		const int raw = tiles <= kRsThreads ? 1 : 0;
		if (!raw) {
			scan_tiles_kernel<<<1, kRsThreads, 0, stream>>>(tile_sums, tiles);
			count_launch();
		}
		scan_finish_kernel<<<tiles, kRsThreads, 0, stream>>>(out, n, tile_sums, raw, in8, ids);
		count_launch();
	}
	return cudaGetLastError();
}
cudaError_t inclusive_sum_gather(const uint32_t *in, const uint32_t *order, uint32_t *out, int64_t n, void *scratch, cudaStream_t stream)
{
	return inclusive_sum_any(in, nullptr, order, out, nullptr, n, scratch, stream);
}
cudaError_t inclusive_sum_mask(const uint8_t *mask, uint32_t *out, uint32_t *ids, int64_t n, void *scratch, cudaStream_t stream)
{
	return inclusive_sum_any(nullptr, mask, nullptr, out, ids, n, scratch, stream);
}

} // namespace gsr
