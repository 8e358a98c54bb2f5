// gsr_sort.cuh — hand-written device primitives for the binning stage: a stable LSD radix sort of
// (u32 key, u32 value) pairs and an inclusive prefix sum, sized for this path's two sorts
// (P depth keys x 32 bits; R tile ids x ceil(log2 tiles) bits).
//
// Replaces cub::DeviceRadixSort::SortPairs / cub::DeviceScan::InclusiveSum as used by the reference
// (CR/rasterizer_impl.cu:283,309-314).  Stability is what makes the two-level binning reproduce the
// reference's order (gsr_binning.cu), so every step below preserves input order among equal digits.
//
// Structure (8-bit digits, tiles of 4096 elements = 256 threads x 16):
//   per digit pass:
//   rs_histogram_kernel : per-tile digit histogram of the current element order + global digit totals
//                         (counts[digit][tile], totals[digit]).
//   rs_offsets_kernel   : one CTA per digit: exclusive scan of that digit's row over the tiles, offset by the
//                         digit's global base -> counts becomes the first output index of (digit, tile).
//   rs_scatter_kernel   : each warp owns a contiguous 512-element run of its tile; ranks by __match_any_sync in
//                         input order, reorders the tile through shared memory so that equal digits are
//                         contiguous, and writes coalesced runs to their final positions.
#pragma once
#include "gsr_common.cuh"

namespace gsr {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems; // 4096
constexpr int kRsMaxPasses = 4;
constexpr int kRsTotalsByRows = 1024; // above this many tiles the digit totals are row sums, not atomics

struct RadixPlan {
	int passes;       // ceil(bits / 8)
	int tiles;        // ceil(n / 4096)
	size_t counts_off;  // u32[256][tiles]
	size_t totals_off;  // u32[passes][256]
	size_t bytes;     // scratch bytes (aligned)
};
RadixPlan radix_plan(int64_t n, int bits);

// Sorts n pairs by the low `bits` bits of the key (stable).  keys_a/vals_a hold the input; the result lands in
// (keys_a, vals_a) when the pass count is even and in (keys_b, vals_b) when it is odd — the return value tells
// which (0 = a, 1 = b).  `scratch` must hold radix_plan(n, bits).bytes.  Launches on `stream`, never syncs.
// n_dev != nullptr: the element count is only known on the device (*n_dev, e.g. num_rendered); `n` is then the capacity the
// arrays and the grids are sized for, and min(*n_dev, n) elements are sorted — no host round trip between the kernel that
// produces the count and the sort.
int radix_sort_pairs(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n, const uint32_t *n_dev, int bits,
                     void *scratch, cudaStream_t stream, cudaError_t *err);
__device__ __forceinline__ int64_t device_count(int64_t cap, const uint32_t *n_dev)
{
	if (!n_dev) return cap;
	const int64_t n = (int64_t)__ldg(n_dev);
	return n < cap ? n : cap;
}

// Inclusive prefix sum of in[order[i]] (order may be null: in[i]) into out[0..n).  scratch: scan_scratch_bytes(n).
size_t scan_scratch_bytes(int64_t n);
cudaError_t inclusive_sum_gather(const uint32_t *in, const uint32_t *order, uint32_t *out, int64_t n, void *scratch, cudaStream_t stream);
// Inclusive prefix sum of the 0 / 1 flags of a byte mask (mask[i] != 0) into out[0..n); with ids != null also the compaction
// ids[out[i] - 1] = i of the set entries (ascending).  Same scratch.
cudaError_t inclusive_sum_mask(const uint8_t *mask, uint32_t *out, uint32_t *ids, int64_t n, void *scratch, cudaStream_t stream);

} // namespace gsr
