// gsr_decode.cu — fused anchor -> neural-Gaussian decode, forward and backward.
//
// Replaces `generate_neural_gaussians` of W-Ted/GScream (gaussian_renderer/__init__.py:18-102, use_feat_bank = False),
// the step that runs immediately before the rasterizer every iteration: gather the visible anchors, build the
// view-dependent input [feat(32) | ob_view(3) | ob_dist(1)], run the four 36->32->{k, k, 7k, 3k} MLPs
// (scene/gaussian_model.py:118-144: opacity/Tanh, uncertainty/Sigmoid, cov/linear, colour/Sigmoid), keep the offsets
// whose neural opacity is > 0 and post-process them into the rasterizer's inputs (xyz, colour, opacity,
// uncertainty, scaling, rotation).  The reference does this with ~40 elementwise / index / cat / split launches and
// several (A*k)-sized temporaries plus the MLPs' eight GEMM launches; here it is
//   stage 1: visible-anchor list (scan) -> opacity MLP -> neural_opacity, mask, per-anchor counts -> scan
//   stage 2: the other three MLPs + post-processing, written straight into the compacted SoA outputs
//   backward: one kernel that recomputes the activations, back-propagates to anchor / feature / offset / scaling
//             and accumulates the sixteen weight / bias gradients on chip (registers across a persistent CTA).
//
// The MLPs run on the tensor pipe: mma.sync.m16n8k8 TF32 with the 3xTF32 split (x = hi + lo; lo*hi + hi*lo + hi*hi, FP32
// accumulate — FP32-grade products, as in the blend kernels).  Every operand map below is modelled lane by lane in
// tests/_decode_fragments.py (same names, same expressions) and checked against plain matrix products on the CPU
// (tests/test_decode_fragments.py).  With lane = 4 g + t:
//     A  a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)      B  b0 (k = t, n = g)  b1 (k = t+4, n = g)
//     C  c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
// "Part A" (all three kernels): one warp owns a tile of 16 visible anchors = the M dimension.  The gather builds the A
// fragments of x straight from global memory (lane t reads feat[8t .. 8t+7] of rows g and g+8; the contraction slots are
// permuted to match, in1()), layer 1 leaves H as C fragments, and a C fragment IS the A fragment of the next product when
// contraction slot t is column 2t and slot t+4 is column 2t+1 (c_as_a) — so layer 2 forward (H -> OUT) and layer 1 backward
// (dH -> dx) chain in registers.  The 12k pre-activations go to a per-warp shared-memory tile OUT[16][S] where the
// per-(anchor, offset) activations / post-processing pass picks them up (and, in the backward, replaces them by their
// gradients in place).  The weights sit in shared memory in fragment order (one conflict-free LDS.128 per lane brings the B
// fragments of two tiles).
// "Part B" (backward only): the weight gradients contract over anchors, which a C fragment of part A holds along its rows —
// the one dimension that cannot become a contraction index without a transpose.  So after a CTA barrier warp (m, hh) —
// MLP m, hidden units 16hh .. 16hh+15 — walks over the iteration's anchors in n-tiles of 8 with the hidden units as the M
// dimension: it recomputes H^T = relu(W1 x^T + b1) and dH^T = W2^T dOUT^T from the X / OUT tiles part A left in shared memory
// (+1/3 MMAs), and now both C fragments have anchors along their columns: dH^T is the A operand of dW1 += dH^T x, H^T is the B
// operand of dW2 += dOUT^T H.  The accumulators (<= 92 registers per lane) stay in registers across the persistent CTA's
// anchors; one atomic per entry at the end.
#include "gsr_internal.cuh"
#include "gsr_sort.cuh"
#include "gsr_decode.cuh"

namespace gsr {

namespace {

constexpr int kHid = 32;          // hidden width == feat_dim == warp width
constexpr int kIn = 36;           // feat_dim + 3 + 1
constexpr int kTile = kDecNA;     // anchors per warp tile (the M dimension of part A)
constexpr int kWarps = 8;
constexpr int kSX = 41;           // X tile row stride: [0..31] feat, [32..34] view dir, [35] dist, [36..39] their gradients, [40] anchor id
constexpr int kPP = 9;            // per (anchor, offset) partials: d anchor (3), d get_scaling (6)
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline int out_count(int m, int k) { return m == 0 ? k : m == 1 ? k : m == 2 ? 7 * k : 3 * k; }
// 1 / (1 + e^-x) with the SFU exponential (ex2.approx on x log2 e: relative error ~1e-6 at |x| = 20, i.e. <= 3e-7 absolute on
// the sigmoid — the parity bound of the decode is 2e-5) and a correctly rounded reciprocal
__device__ __forceinline__ float sigmoidf_(float x) { return __frcp_rn(1.0f + __expf(-x)); }

// Column plan of the OUT tile for the MLPs [m0, m1): each MLP padded to a multiple of 16 columns (pad columns hold zeros),
// row stride S = cols + 4 (S = 4 mod 8 with S/4 odd: the fragment loads of parts A and B are bank-conflict free).
struct ColPlan {
	int cb[4], cols, S;
};
__host__ __device__ inline ColPlan col_plan(int k, int m0, int m1)
{
	ColPlan p{};
	int c = 0;
	for (int m = 0; m < 4; m++) {
		p.cb[m] = c;
		if (m >= m0 && m < m1) c += (out_count(m, k) + 15) / 16 * 16;
	}
	p.cols = c;
	p.S = c + 4;
	return p;
}

// shared-memory carve-up (in floats)
struct SmemPlan {
	int w1f, w2f, w1b, w2b, b1, b2, warp0, per_warp, out, tab, x, pp, total;
};
constexpr int kTab = 12;          // per-anchor table row: id, mask bits, first output row, anchor xyz (3), get_scaling (6)
__host__ __device__ inline SmemPlan smem_plan(int k, int m0, int m1, bool backward, int tiles)
{
	const ColPlan cp = col_plan(k, m0, m1);
	const int nm = m1 - m0;
	SmemPlan p{};
	int off = 0;
	p.w1f = off;  off += nm * 5 * 2 * 32 * 4;
	p.w2f = off;  off += (cp.cols / 8) * 2 * 32 * 4;
	p.w1b = off;  if (backward) off += nm * 2 * 5 * 32 * 4;
	p.w2b = off;  if (backward) off += (cp.cols / 8) * 2 * 32 * 4;
	p.b1 = off;   off += nm * kHid;
	p.b2 = off;   off += cp.cols + 8;
	off = (off + 3) & ~3;
	p.warp0 = off;
	int w = 0;
	p.out = w;  w += kTile * cp.S;
	p.tab = w;  if (m1 > 1) w += kTile * kTab;
	p.x = w;    if (backward) w += kTile * kSX;
	p.pp = w;   if (backward) w += kTile * k * kPP;
	w = (w + 3) & ~3;
	p.per_warp = w;
	p.total = off + tiles * w;
	return p;
}

// ---- tensor-pipe helpers (the same instruction and split as gsr_blend.cuh) ---------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
	hi = __float_as_uint(x) & 0xffffe000u;
	lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
	asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
	             : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
	             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
struct AFrag {
	uint32_t hi[4], lo[4];
};
__device__ __forceinline__ AFrag make_a(float a0, float a1, float a2, float a3)
{
	AFrag f;
	split_tf32(a0, f.hi[0], f.lo[0]);
	split_tf32(a1, f.hi[1], f.lo[1]);
	split_tf32(a2, f.hi[2], f.lo[2]);
	split_tf32(a3, f.hi[3], f.lo[3]);
	return f;
}
// a C fragment as the A fragment of the next product: contraction slot t <-> column 2t, slot t+4 <-> column 2t+1
__device__ __forceinline__ AFrag c_as_a(const float (&c)[4]) { return make_a(c[0], c[2], c[1], c[3]); }
// c += a b with FP32-grade products
__device__ __forceinline__ void mma3(float (&c)[4], const AFrag &a, float b0, float b1)
{
	uint32_t b0h, b0l, b1h, b1l;
	split_tf32(b0, b0h, b0l);
	split_tf32(b1, b1h, b1l);
	mma_tf32(c, a.lo, b0h, b1h);
	mma_tf32(c, a.hi, b0l, b1l);
	mma_tf32(c, a.hi, b0h, b1h);
}
// The same with the two correction products in an accumulator of their own (c + d is the result): dependent HMMAs are
// issued back to back otherwise, and a chain of them — 3 per k-step — is what a warp waits on in the short products.
__device__ __forceinline__ void mma3(float (&c)[4], float (&d)[4], const AFrag &a, float b0, float b1)
{
	uint32_t b0h, b0l, b1h, b1l;
	split_tf32(b0, b0h, b0l);
	split_tf32(b1, b1h, b1l);
	mma_tf32(c, a.hi, b0h, b1h);
	mma_tf32(d, a.lo, b0h, b1h);
	mma_tf32(d, a.hi, b0l, b1l);
}
__device__ __forceinline__ void zero4(float (&c)[4]) { c[0] = c[1] = c[2] = c[3] = 0.f; }
__device__ __forceinline__ void add4(float (&c)[4], const float (&d)[4])
{
	c[0] += d[0]; c[1] += d[1]; c[2] += d[2]; c[3] += d[3];
}
__device__ __forceinline__ float f4c(const float4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// ---- input-slot maps ---------------------------------------------------------------------------------------------------------
// layer-1 contraction slot s (0..7) of k-step ks (0..4) -> input index, -1 = zero pad.  Lane t holds feat[8t .. 8t+7] as two
// float4 F0, F1: slot t of k-step ks is F0[ks], slot t+4 is F1[ks]; k-step 4 carries (ux, uy, uz, dist) in slots 0..3.
__host__ __device__ inline int in1(int ks, int s) { return ks < 4 ? 8 * (s & 3) + 4 * (s >> 2) + ks : (s < 4 ? 32 + s : -1); }
// layer-1 backward output column c (0..7) of n-tile nt (0..4) -> input index: lane t ends up with d feat[8t .. 8t+7] in tiles
// 0..3 (two float4 stores); tile 4 carries the view / distance gradients in columns 0..3.
__host__ __device__ inline int in1b(int nt, int c) { return nt < 4 ? 8 * (c >> 1) + 2 * nt + (c & 1) : (c < 4 ? 32 + c : -1); }

// ---- shared-memory weight copies in fragment order ---------------------------------------------------------------------------
//   W1F[m][ks][p][lane][4]   (b0, b1) of n-tiles 2p, 2p+1:    b0 = W1[m][8nt+g][in1(ks,t)]      b1 = W1[m][8nt+g][in1(ks,t+4)]
//   W2F[tile8][ksp][lane][4] (b0, b1) of k-steps 2ksp, +1:    b0 = W2[m][8nt+g][8ks+2t]         b1 = W2[m][8nt+g][8ks+2t+1]
//   W1B[m][ksp][nt][lane][4] (b0, b1) of k-steps 2ksp, +1:    b0 = W1[m][8ks+2t][in1b(nt,g)]    b1 = W1[m][8ks+2t+1][in1b(nt,g)]
//   W2B[tile8][p][lane][4]   (b0, b1) of n-tiles 2p, 2p+1:    b0 = W2[m][8ks+t][8nt+g]          b1 = W2[m][8ks+t+4][8nt+g]
// (tile8 = (cb[m] / 8) + tile index inside MLP m; rows o >= out_count(m) are zeros).  W1F and W2B double as the A operands of
// part B: read as (a0, a2, a1, a3) of the 16 hidden units 16p .. 16p+15.
// Loads go out in batches of eight per thread before the first store (a load -> store loop would pay the L2 latency once per
// element: 15 to 76 elements per thread).
template <typename F>
__device__ __forceinline__ void fill_smem(float *dst, int n, int tid, F value_of)
{
	constexpr int kBatch = 8;     // (16 measured no faster)
	for (int base = tid; base < n; base += 256 * kBatch) {
		float v[kBatch];
#pragma unroll
		for (int u = 0; u < kBatch; u++) {
			const int idx = base + 256 * u;
			v[u] = idx < n ? value_of(idx) : 0.f;
		}
#pragma unroll
		for (int u = 0; u < kBatch; u++) {
			const int idx = base + 256 * u;
			if (idx < n) dst[idx] = v[u];
		}
	}
}
__device__ __forceinline__ void load_weights(float *sm, const SmemPlan &pl, const ColPlan &cp, const DecodeWeights &wt, int k, int m0, int m1,
                                             bool backward, int tid)
{
	const int nm = m1 - m0;
	fill_smem(sm + pl.w1f, nm * 1280, tid, [&](int idx) {
		const int e = idx & 3, lane = (idx >> 2) & 31, g = lane >> 2, t = lane & 3;
		const int p = (idx >> 7) & 1, ks = (idx >> 8) % 5, m = m0 + idx / 1280;
		const int nt = 2 * p + (e >> 1), i = in1(ks, t + 4 * (e & 1));
		return i < 0 ? 0.f : __ldg(wt.w1[m] + (8 * nt + g) * kIn + i);
	});
	if (backward)
		fill_smem(sm + pl.w1b, nm * 1280, tid, [&](int idx) {
			const int e = idx & 3, lane = (idx >> 2) & 31, g = lane >> 2, t = lane & 3;
			const int nt = (idx >> 7) % 5, ksp = ((idx >> 7) / 5) & 1, m = m0 + idx / 1280;
			const int h = 8 * (2 * ksp + (e >> 1)) + 2 * t + (e & 1), i = in1b(nt, g);
			return i < 0 ? 0.f : __ldg(wt.w1[m] + h * kIn + i);
		});
	auto mlp_of_tile = [&](int tile8) {
		int m = m0;
		while (m + 1 < m1 && cp.cb[m + 1] <= 8 * tile8) m++;
		return m;
	};
	fill_smem(sm + pl.w2f, cp.cols * 32, tid, [&](int idx) {   // q = ksp
		const int e = idx & 3, lane = (idx >> 2) & 31, g = lane >> 2, t = lane & 3;
		const int q = (idx >> 7) & 1, tile8 = idx >> 8, m = mlp_of_tile(tile8);
		const int o = 8 * (tile8 - cp.cb[m] / 8) + g, h = 8 * (2 * q + (e >> 1)) + 2 * t + (e & 1);
		return o < out_count(m, k) ? __ldg(wt.w2[m] + o * kHid + h) : 0.f;
	});
	if (backward)
		fill_smem(sm + pl.w2b, cp.cols * 32, tid, [&](int idx) {   // q = p
			const int e = idx & 3, lane = (idx >> 2) & 31, g = lane >> 2, t = lane & 3;
			const int q = (idx >> 7) & 1, tile8 = idx >> 8, m = mlp_of_tile(tile8);
			const int o = 8 * (tile8 - cp.cb[m] / 8) + t + 4 * (e & 1), h = 8 * (2 * q + (e >> 1)) + g;
			return o < out_count(m, k) ? __ldg(wt.w2[m] + o * kHid + h) : 0.f;
		});
	fill_smem(sm + pl.b1, nm * kHid, tid, [&](int idx) { return __ldg(wt.b1[m0 + (idx >> 5)] + (idx & 31)); });
	fill_smem(sm + pl.b2, cp.cols + 8, tid, [&](int c) {
		float v = 0.f;
		for (int m = m0; m < m1; m++)
			if (c >= cp.cb[m] && c < cp.cb[m] + out_count(m, k)) v = __ldg(wt.b2[m] + (c - cp.cb[m]));
		return v;
	});
}

// ---- the gather ---------------------------------------------------------------------------------------------------------------
// One tile of 16 visible anchors as it comes from global memory: lane (g, t) holds rows g and g+8.  Fetched one loop iteration
// ahead of its use, so that the two dependent loads (visible list -> anchor row / feature row) are off the critical path.
struct TileRaw {
	int id[2];
	bool valid[2];
	float4 f[2][2];                 // feat[8t .. 8t+3], feat[8t+4 .. 8t+7]
	float ax[2], ay[2], az[2];
	uint32_t bits[2], incl[2];      // TAB: the anchor's kept-offset mask and inclusive output count (lane t == 0)
	float gsa[2], gsb[2];           // TAB: get_scaling[t], get_scaling[4 + t] (t < 2)
};
// TAB: also fetch what the per-(anchor, offset) pass needs per anchor (stage 2 and the backward; stage 1 produces it)
template <bool TAB>
__device__ __forceinline__ void fetch_tile(TileRaw &r, int tile, int ntiles, int n_vis, const DecodeArgs &a, int g, int t)
{
	const uint32_t *__restrict__ vis_ids = a.vis_ids;
	const float *__restrict__ anchor = a.anchor, *__restrict__ feat = a.feat;
#pragma unroll
	for (int q = 0; q < 2; q++) {
		const int rank = tile * kTile + g + 8 * q;
		r.valid[q] = tile < ntiles && rank < n_vis;
		r.id[q] = r.valid[q] ? (vis_ids ? (int)__ldg(vis_ids + rank) : rank) : 0;
	}
#pragma unroll
	for (int q = 0; q < 2; q++) {
		r.f[q][0] = r.f[q][1] = make_float4(0.f, 0.f, 0.f, 0.f);
		r.ax[q] = r.ay[q] = r.az[q] = 0.f;
		if (r.valid[q]) {
			const float4 *row = reinterpret_cast<const float4 *>(feat + (size_t)r.id[q] * kHid + 8 * t);
			r.f[q][0] = __ldg(row);
			r.f[q][1] = __ldg(row + 1);
			r.ax[q] = __ldg(anchor + (size_t)r.id[q] * 3 + 0);
			r.ay[q] = __ldg(anchor + (size_t)r.id[q] * 3 + 1);
			r.az[q] = __ldg(anchor + (size_t)r.id[q] * 3 + 2);
		}
		if (TAB) {
			r.bits[q] = r.incl[q] = 0u;
			r.gsa[q] = r.gsb[q] = 0.f;
			if (r.valid[q]) {
				const int rank = tile * kTile + g + 8 * q;
				if (t == 0) {
					r.bits[q] = __ldg(a.maskbits + rank);
					r.incl[q] = __ldg(a.gauss_incl + rank);
				}
				r.gsa[q] = __ldg(a.scaling + (size_t)r.id[q] * 6 + t);
				if (t < 2) r.gsb[q] = __ldg(a.scaling + (size_t)r.id[q] * 6 + 4 + t);
			}
		}
	}
}
// the per-anchor table of the tile, from the registers of the gather
__device__ __forceinline__ void write_table(float *TAB, const TileRaw &r, int g, int t)
{
#pragma unroll
	for (int q = 0; q < 2; q++) {
		float *row = TAB + (g + 8 * q) * kTab;
		row[6 + t] = r.gsa[q];
		if (t < 2) row[10 + t] = r.gsb[q];
		if (t == 0) {
			row[0] = __int_as_float(r.valid[q] ? r.id[q] : -1);
			row[1] = __uint_as_float(r.bits[q]);
			row[2] = __uint_as_float(r.incl[q] - (uint32_t)__popc(r.bits[q]));
			row[3] = r.ax[q];
			row[4] = r.ay[q];
			row[5] = r.az[q];
		}
	}
}
// The tile's input x = [feat | ob_view | ob_dist] (gaussian_renderer/__init__.py:26-52) as A-fragment values: xa[ks] =
// (row g slot t, row g+8 slot t, row g slot t+4, row g+8 slot t+4).  Rows past n_vis are zeros.
struct TileX {
	float xa[5][4];
	float u[2][3], dist[2];
};
__device__ __forceinline__ void stage_tile(TileX &x, const TileRaw &r, float cx, float cy, float cz, int t)
{
#pragma unroll
	for (int q = 0; q < 2; q++) {
		x.u[q][0] = x.u[q][1] = x.u[q][2] = 0.f;
		x.dist[q] = 0.f;
		if (r.valid[q]) {
			const float vx = r.ax[q] - cx, vy = r.ay[q] - cy, vz = r.az[q] - cz;
			x.dist[q] = sqrtf(vx * vx + vy * vy + vz * vz);
			x.u[q][0] = vx / x.dist[q];
			x.u[q][1] = vy / x.dist[q];
			x.u[q][2] = vz / x.dist[q];
		}
	}
#pragma unroll
	for (int ks = 0; ks < 4; ks++) {
		x.xa[ks][0] = f4c(r.f[0][0], ks);
		x.xa[ks][1] = f4c(r.f[1][0], ks);
		x.xa[ks][2] = f4c(r.f[0][1], ks);
		x.xa[ks][3] = f4c(r.f[1][1], ks);
	}
#pragma unroll
	for (int q = 0; q < 2; q++) x.xa[4][q] = t == 0 ? x.u[q][0] : t == 1 ? x.u[q][1] : t == 2 ? x.u[q][2] : x.dist[q];
	x.xa[4][2] = x.xa[4][3] = 0.f;
}

// ---- part A products ----------------------------------------------------------------------------------------------------------
// layer 1 of MLP m (ml = m - m0): h[nt] = C fragments of relu(x W1^T + b1): rows = anchors, columns = hidden units 8nt + 2t, +1
__device__ __forceinline__ void layer1_forward(float (&h)[4][4], const TileX &x, const float *sm, const SmemPlan &pl, int ml, int lane, int t)
{
#pragma unroll
	for (int nt = 0; nt < 4; nt++) {
		const float2 b = *reinterpret_cast<const float2 *>(sm + pl.b1 + ml * kHid + 8 * nt + 2 * t);
		h[nt][0] = h[nt][2] = b.x;
		h[nt][1] = h[nt][3] = b.y;
	}
	float hc[4][4];
#pragma unroll
	for (int nt = 0; nt < 4; nt++) zero4(hc[nt]);
#pragma unroll
	for (int ks = 0; ks < 5; ks++) {
		const AFrag a = make_a(x.xa[ks][0], x.xa[ks][1], x.xa[ks][2], x.xa[ks][3]);
#pragma unroll
		for (int p = 0; p < 2; p++) {
			const float4 e = *reinterpret_cast<const float4 *>(sm + pl.w1f + (((ml * 5 + ks) * 2 + p) * 32 + lane) * 4);
			mma3(h[2 * p], hc[2 * p], a, e.x, e.y);
			mma3(h[2 * p + 1], hc[2 * p + 1], a, e.z, e.w);
		}
	}
#pragma unroll
	for (int nt = 0; nt < 4; nt++)
#pragma unroll
		for (int e = 0; e < 4; e++) h[nt][e] = fmaxf(h[nt][e] + hc[nt][e], 0.f);
}
// layer 2 of MLP m: OUT[a][cb + o] = b2[o] + sum_h W2[o][h] H[a][h], H from the C fragments of layer 1
__device__ __forceinline__ void layer2_forward(const float (&h)[4][4], const float *sm, const SmemPlan &pl, float *OUT, int S, int cb, int n_m,
                                               int lane, int g, int t)
{
	AFrag ha[4];
#pragma unroll
	for (int ks = 0; ks < 4; ks++) ha[ks] = c_as_a(h[ks]);
	// one 16-column block = two n-tiles per iteration (the column plan pads every MLP to whole blocks; pad tiles have zero
	// weights and biases and produce the zeros the pad columns must hold)
	const int nblk = (n_m + 15) >> 4;
	for (int blk = 0; blk < nblk; blk++) {
		const int col = cb + 16 * blk + 2 * t;
		const float2 bA = *reinterpret_cast<const float2 *>(sm + pl.b2 + col), bB = *reinterpret_cast<const float2 *>(sm + pl.b2 + col + 8);
		float accA[4] = {bA.x, bA.y, bA.x, bA.y}, accB[4] = {bB.x, bB.y, bB.x, bB.y}, corA[4], corB[4];
		zero4(corA);
		zero4(corB);
		const float *w = sm + pl.w2f + ((cb >> 3) + 2 * blk) * 256 + lane * 4;
#pragma unroll
		for (int ksp = 0; ksp < 2; ksp++) {
			const float4 eA = *reinterpret_cast<const float4 *>(w + ksp * 128), eB = *reinterpret_cast<const float4 *>(w + 256 + ksp * 128);
			mma3(accA, corA, ha[2 * ksp], eA.x, eA.y);
			mma3(accB, corB, ha[2 * ksp], eB.x, eB.y);
			mma3(accA, corA, ha[2 * ksp + 1], eA.z, eA.w);
			mma3(accB, corB, ha[2 * ksp + 1], eB.z, eB.w);
		}
		add4(accA, corA);
		add4(accB, corB);
		*reinterpret_cast<float2 *>(OUT + g * S + col) = make_float2(accA[0], accA[1]);
		*reinterpret_cast<float2 *>(OUT + (g + 8) * S + col) = make_float2(accA[2], accA[3]);
		*reinterpret_cast<float2 *>(OUT + g * S + col + 8) = make_float2(accB[0], accB[1]);
		*reinterpret_cast<float2 *>(OUT + (g + 8) * S + col + 8) = make_float2(accB[2], accB[3]);
	}
}
// layer 2 backward of MLP m: dh[nt] = C fragments of dOUT[:, MLP m] W2 (rows = anchors, columns = hidden 8nt + 2t, +1), not gated
__device__ __forceinline__ void layer2_backward(float (&dh)[4][4], const float *sm, const SmemPlan &pl, const float *OUT, int S, int cb, int n_m,
                                                int lane, int g, int t)
{
#pragma unroll
	for (int nt = 0; nt < 4; nt++) zero4(dh[nt]);
	float dc[4][4];
#pragma unroll
	for (int nt = 0; nt < 4; nt++) zero4(dc[nt]);
	const int nks = (n_m + 7) >> 3;
	for (int ks = 0; ks < nks; ks++) {
		const float *r0 = OUT + g * S + cb + 8 * ks + t, *r1 = r0 + 8 * S;
		const AFrag a = make_a(r0[0], r1[0], r0[4], r1[4]);
#pragma unroll
		for (int p = 0; p < 2; p++) {
			const float4 e = *reinterpret_cast<const float4 *>(sm + pl.w2b + (((cb >> 3) + ks) * 2 + p) * 128 + lane * 4);
			mma3(dh[2 * p], dc[2 * p], a, e.x, e.y);
			mma3(dh[2 * p + 1], dc[2 * p + 1], a, e.z, e.w);
		}
	}
#pragma unroll
	for (int nt = 0; nt < 4; nt++) add4(dh[nt], dc[nt]);
}
// layer 1 backward of MLP m: dx[nt] += dh W1; afterwards lane t holds d feat[8t + 2nt + e] of rows g (c0, c1) and g+8 (c2, c3)
__device__ __forceinline__ void layer1_backward(float (&dx)[5][4], float (&dxc)[5][4], const float (&dh)[4][4], const float *sm, const SmemPlan &pl,
                                                int ml, int lane)
{
#pragma unroll
	for (int ksp = 0; ksp < 2; ksp++) {
		const AFrag a0 = c_as_a(dh[2 * ksp]), a1 = c_as_a(dh[2 * ksp + 1]);
#pragma unroll
		for (int nt = 0; nt < 5; nt++) {
			const float4 e = *reinterpret_cast<const float4 *>(sm + pl.w1b + (((ml * 2 + ksp) * 5 + nt) * 32 + lane) * 4);
			mma3(dx[nt], dxc[nt], a0, e.x, e.y);
			mma3(dx[nt], dxc[nt], a1, e.z, e.w);
		}
	}
}

// ---- stage 1: opacity MLP, mask, per-anchor counts ---------------------------------------------------------------
#ifndef GSR_DEC_S1_CTAS
#define GSR_DEC_S1_CTAS 2     // (3 CTAs of 80 registers measured the same: 0.133 ms either way)
#endif
__global__ void __launch_bounds__(256, GSR_DEC_S1_CTAS) decode_opacity_kernel(DecodeArgs a)
{
	extern __shared__ __align__(16) float sm[];
	const int k = a.k;
	const ColPlan cp = col_plan(k, 0, 1);
	const SmemPlan pl = smem_plan(k, 0, 1, false, kWarps);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
	for (int i = tid; i < kWarps * pl.per_warp; i += 256) sm[pl.warp0 + i] = 0.f;
	load_weights(sm, pl, cp, a.wt, k, 0, 1, false, tid);
	__syncthreads();
	float *OUT = sm + pl.warp0 + warp * pl.per_warp + pl.out;
	const int S = cp.S;
	const int n_vis = a.n_vis_dev ? (int)*a.n_vis_dev : a.n_vis;
	const int ntiles = (n_vis + kTile - 1) / kTile;
	// the scan behind this kernel runs over all A counts: ranks past the visible anchors count nothing
	for (int r = n_vis + blockIdx.x * 256 + tid; r < a.n_vis; r += gridDim.x * 256) a.count[r] = 0u;
	const float cx = __ldg(a.campos), cy = __ldg(a.campos + 1), cz = __ldg(a.campos + 2);
	TileRaw raw;
	fetch_tile<false>(raw, blockIdx.x * kWarps + warp, ntiles, n_vis, a, g, t);
	for (int tile = blockIdx.x * kWarps + warp; tile < ntiles; tile += gridDim.x * kWarps) {
		TileX x;
		stage_tile(x, raw, cx, cy, cz, t);
		fetch_tile<false>(raw, tile + gridDim.x * kWarps, ntiles, n_vis, a, g, t);
		float h[4][4];
		layer1_forward(h, x, sm, pl, 0, lane, t);
		layer2_forward(h, sm, pl, OUT, S, 0, k, lane, g, t);
		__syncwarp();
		// lanes = (anchor, offset) pairs
		for (int idx = lane; idx < kTile * k; idx += 32) {
			const int aa = idx / k, j = idx - aa * k;
			const int r = tile * kTile + aa;
			if (r < n_vis) {
				const float nop = tanhf(OUT[aa * S + j]);
				a.neural_opacity[(size_t)r * k + j] = nop;
				a.mask[(size_t)r * k + j] = nop > 0.0f ? 1 : 0;                  // gaussian_renderer/__init__.py:59
			}
		}
		// per-anchor count and bit mask: lanes 0..15 re-read their anchor's k pre-activations (cheap, shared memory)
		if (lane < kTile) {
			const int r = tile * kTile + lane;
			if (r < n_vis) {
				uint32_t bits = 0;
				for (int j = 0; j < k; j++)
					if (OUT[lane * S + j] > 0.0f) bits |= 1u << j;   // == tanhf(.) > 0: tanh keeps the sign and does not underflow to 0
				a.count[r] = (uint32_t)__popc(bits);
				a.maskbits[r] = bits;
			}
		}
		__syncwarp();
	}
}

// ---- stage 2: the other three MLPs + post-processing into the compacted outputs ------------------------------------
struct OutPair {          // one (anchor, offset) pair of the post-processing pass and what it reads from global memory
	bool keep;
	int aa, j;
	size_t p;
	float off[3], nop;
};
__global__ void __launch_bounds__(256, 2) decode_outputs_kernel(DecodeArgs a)
{
	extern __shared__ __align__(16) float sm[];
	const int k = a.k;
	const ColPlan cp = col_plan(k, 1, 4);
	const SmemPlan pl = smem_plan(k, 1, 4, false, kWarps);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
	for (int i = tid; i < kWarps * pl.per_warp; i += 256) sm[pl.warp0 + i] = 0.f;
	load_weights(sm, pl, cp, a.wt, k, 1, 4, false, tid);
	__syncthreads();
	float *OUT = sm + pl.warp0 + warp * pl.per_warp + pl.out;
	float *TAB = sm + pl.warp0 + warp * pl.per_warp + pl.tab;
	const int S = cp.S, cU = cp.cb[1], cC = cp.cb[2], cR = cp.cb[3];
	const int n_vis = a.n_vis_dev ? (int)*a.n_vis_dev : a.n_vis; // (device copy when the host has not read the counts yet)
	const int ntiles = (n_vis + kTile - 1) / kTile;
	const float cx = __ldg(a.campos), cy = __ldg(a.campos + 1), cz = __ldg(a.campos + 2);
	TileRaw raw;
	fetch_tile<true>(raw, blockIdx.x * kWarps + warp, ntiles, n_vis, a, g, t);
	for (int tile = blockIdx.x * kWarps + warp; tile < ntiles; tile += gridDim.x * kWarps) {
		TileX x;
		stage_tile(x, raw, cx, cy, cz, t);
		write_table(TAB, raw, g, t);
		fetch_tile<true>(raw, tile + gridDim.x * kWarps, ntiles, n_vis, a, g, t);
		__syncwarp();
		// lanes = (anchor, offset) pairs; the first round's loads are issued here, ahead of the MLPs
		auto load_out_pair = [&](int idx) {
			OutPair q;
			q.keep = false;
			q.aa = q.j = 0;
			q.p = 0;
			q.off[0] = q.off[1] = q.off[2] = q.nop = 0.f;
			if (idx < kTile * k) {
				q.aa = idx / k;
				q.j = idx - q.aa * k;
				const float *tb = TAB + q.aa * kTab;
				const uint32_t bits = __float_as_uint(tb[1]);        // 0 for rows past n_vis
				if ((bits >> q.j) & 1u) {
					q.keep = true;
					const int id = __float_as_int(tb[0]);
					q.p = (size_t)__float_as_uint(tb[2]) + __popc(bits & ((1u << q.j) - 1u));
					const float *off = a.offset + ((size_t)id * k + q.j) * 3;
					q.off[0] = __ldg(off);
					q.off[1] = __ldg(off + 1);
					q.off[2] = __ldg(off + 2);
					q.nop = __ldg(a.neural_opacity + (size_t)(tile * kTile + q.aa) * k + q.j);
				}
			}
			return q;
		};
		OutPair cur = load_out_pair(lane);
#pragma unroll
		for (int m = 1; m < 4; m++) {
			float h[4][4];
			layer1_forward(h, x, sm, pl, m - 1, lane, t);
			layer2_forward(h, sm, pl, OUT, S, cp.cb[m], out_count(m, k), lane, g, t);
		}
		__syncwarp();
		for (int idx = lane; idx < kTile * k; idx += 32) {
			const OutPair nx = load_out_pair(idx + 32);      // the next round's loads fly during this round's arithmetic
			if (cur.keep) {
				const float *tb = TAB + cur.aa * kTab, *gs = tb + 6, *an = tb + 3;
				const float *row = OUT + cur.aa * S;
				const size_t p = cur.p;
				const int j = cur.j;
				a.out_opacity[p] = cur.nop;
				a.out_uncertainty[p] = sigmoidf_(row[cU + j]);
				float sr[7];
#pragma unroll
				for (int c = 0; c < 7; c++) sr[c] = row[cC + 7 * j + c];
#pragma unroll
				for (int c = 0; c < 3; c++) {
					a.out_color[p * 3 + c] = sigmoidf_(row[cR + 3 * j + c]);
					a.out_scaling[p * 3 + c] = gs[3 + c] * sigmoidf_(sr[c]);                         // :89
					a.out_xyz[p * 3 + c] = an[c] + cur.off[c] * gs[c];                              // :93-94
				}
				// F.normalize (scene/gaussian_model.py:52): v / max(||v||, 1e-12)
				const float nrm = fmaxf(sqrtf(sr[3] * sr[3] + sr[4] * sr[4] + sr[5] * sr[5] + sr[6] * sr[6]), 1e-12f);
#pragma unroll
				for (int c = 0; c < 4; c++) a.out_rot[p * 4 + c] = sr[3 + c] / nrm;
			}
			cur = nx;
		}
		__syncwarp();
	}
}

// ---- backward -------------------------------------------------------------------------------------------------------
// Per CTA iteration `tiles` warps run part A on one tile each: recompute the activations, turn the upstream gradients of the
// kept Gaussians into gradients of the 12k output pre-activations (in place over OUT), back-propagate through layer 2 and
// layer 1 in registers, write d_feat / d_anchor / d_offset / d_scaling.  After a CTA barrier all eight warps run part B over
// the iteration's X / OUT tiles (see the file header).  KMT = 16-row tiles of the widest MLP (cov: 7k rows).
struct BwdPair {          // one (anchor, offset) pair of the activation-backward pass and what it reads from global memory
	bool on, valid, keep;
	int aa, j, aid;
	float g_nop, g_opa, g_unc, g_col[3], g_scl[3], g_xyz[3], g_rot[4], off[3];
};
template <int KMT>
__global__ void __launch_bounds__(256, 1) decode_backward_kernel(DecodeBwdArgs a, int tiles)
{
	extern __shared__ __align__(16) float sm[];
	const int k = a.f.k;
	const ColPlan cp = col_plan(k, 0, 4);
	const SmemPlan pl = smem_plan(k, 0, 4, true, tiles);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
	for (int i = tid; i < tiles * pl.per_warp; i += 256) sm[pl.warp0 + i] = 0.f;
	load_weights(sm, pl, cp, a.f.wt, k, 0, 4, true, tid);
	__syncthreads();
	const int S = cp.S, cU = cp.cb[1], cC = cp.cb[2], cR = cp.cb[3];
	const int my = warp < tiles ? warp : 0;
	float *OUT = sm + pl.warp0 + my * pl.per_warp + pl.out;
	float *X = sm + pl.warp0 + my * pl.per_warp + pl.x;
	float *PP = sm + pl.warp0 + my * pl.per_warp + pl.pp;
	float *TAB = sm + pl.warp0 + my * pl.per_warp + pl.tab;
	const int n_vis = a.f.n_vis;
	const int ntiles = (n_vis + kTile - 1) / kTile;
	const int cta_iters = (ntiles + gridDim.x * tiles - 1) / (gridDim.x * tiles);
	const float cx = __ldg(a.f.campos), cy = __ldg(a.f.campos + 1), cz = __ldg(a.f.campos + 2);

	// part B: this warp's slice of the weight gradients — MLP mB, hidden units 16hh .. 16hh+15
	const int mB = warp >> 1, hh = warp & 1;
	const int nB = out_count(mB, k), cbB = cp.cb[mB], nksB = (nB + 7) >> 3, nmtB = (nB + 15) >> 4;
	float acc1[5][4], accb1[2] = {0.f, 0.f};      // dW1: rows h = 16hh + g (+8), columns i = 8nt + 2t (+1)
	float acc2[KMT][2][4], accb2[KMT][2];         // dW2: rows o = 16mt + g (+8), columns h = 16hh + 8nt + 2t (+1)
#pragma unroll
	for (int nt = 0; nt < 5; nt++) acc1[nt][0] = acc1[nt][1] = acc1[nt][2] = acc1[nt][3] = 0.f;
#pragma unroll
	for (int mt = 0; mt < KMT; mt++) {
		accb2[mt][0] = accb2[mt][1] = 0.f;
#pragma unroll
		for (int nt = 0; nt < 2; nt++) acc2[mt][nt][0] = acc2[mt][nt][1] = acc2[mt][nt][2] = acc2[mt][nt][3] = 0.f;
	}
	const float b1lo = sm[pl.b1 + mB * kHid + 16 * hh + g], b1hi = sm[pl.b1 + mB * kHid + 16 * hh + g + 8];

	TileRaw raw;
	fetch_tile<true>(raw, warp < tiles ? blockIdx.x * tiles + warp : ntiles, ntiles, n_vis, a.f, g, t);
	for (int it = 0; it < cta_iters; it++) {
		const int first = (it * gridDim.x + blockIdx.x) * tiles;     // this iteration's tiles: first .. first + tiles - 1
		const int tile = first + warp;
		if (warp < tiles && tile < ntiles) {
			// ================================================ part A ================================================
			TileX x;
			stage_tile(x, raw, cx, cy, cz, t);
			const int id0 = raw.id[0], id1 = raw.id[1];
			const bool v0 = raw.valid[0], v1 = raw.valid[1];
			// the X tile for part B and for the anchor-gradient pass below
#pragma unroll
			for (int q = 0; q < 2; q++) {
				float *xr = X + (g + 8 * q) * kSX;
#pragma unroll
				for (int c = 0; c < 4; c++) {
					xr[8 * t + c] = f4c(raw.f[q][0], c);
					xr[8 * t + 4 + c] = f4c(raw.f[q][1], c);
				}
				xr[32 + t] = x.xa[4][q];
				if (t == 0) xr[40] = __int_as_float(raw.valid[q] ? raw.id[q] : -1);
			}
			write_table(TAB, raw, g, t);
			fetch_tile<true>(raw, tile + gridDim.x * tiles, ntiles, n_vis, a.f, g, t);
			__syncwarp();
			// lanes = (anchor, offset) pairs of the activation-backward pass; the loader only ISSUES loads (no arithmetic on the
			// loaded values), and the first round's go out here, ahead of the recomputation
			auto load_bwd_pair = [&](int idx) {
				BwdPair q;
				q.on = idx < kTile * k;
				q.valid = q.keep = false;
				q.aa = q.j = q.aid = 0;
				q.g_nop = q.g_opa = q.g_unc = 0.f;
#pragma unroll
				for (int c = 0; c < 3; c++) q.g_col[c] = q.g_scl[c] = q.g_xyz[c] = q.off[c] = 0.f;
#pragma unroll
				for (int c = 0; c < 4; c++) q.g_rot[c] = 0.f;
				if (q.on) {
					q.aa = idx / k;
					q.j = idx - q.aa * k;
					const int r = tile * kTile + q.aa;
					q.valid = r < n_vis;
					if (q.valid) {
						const float *tb = TAB + q.aa * kTab;
						const uint32_t bits = __float_as_uint(tb[1]);
						q.keep = (bits >> q.j) & 1u;
						q.aid = __float_as_int(tb[0]);
						if (a.d_neural_opacity) q.g_nop = __ldg(a.d_neural_opacity + (size_t)r * k + q.j);
						if (q.keep) {
							const size_t p = (size_t)__float_as_uint(tb[2]) + __popc(bits & ((1u << q.j) - 1u));
							const float *off = a.f.offset + ((size_t)q.aid * k + q.j) * 3;
							if (a.d_opacity) q.g_opa = __ldg(a.d_opacity + p);
							if (a.d_uncertainty) q.g_unc = __ldg(a.d_uncertainty + p);
#pragma unroll
							for (int c = 0; c < 3; c++) {
								if (a.d_color) q.g_col[c] = __ldg(a.d_color + p * 3 + c);
								if (a.d_scaling) q.g_scl[c] = __ldg(a.d_scaling + p * 3 + c);
								if (a.d_xyz) {
									q.g_xyz[c] = __ldg(a.d_xyz + p * 3 + c);
									q.off[c] = __ldg(off + c);
								}
							}
							if (a.d_rot) {
#pragma unroll
								for (int c = 0; c < 4; c++) q.g_rot[c] = __ldg(a.d_rot + p * 4 + c);
							}
						}
					}
				}
				return q;
			};
			BwdPair cur = load_bwd_pair(lane);
			// (the loops over the four MLPs stay rolled: unrolled, the kernel is 9.6 k instructions and its eight warps, spread
			// over part A and part B code, miss the instruction cache)
			unsigned long long relu = 0ull;     // bit (16 m + 4 nt + e): H[m] fragment element > 0
#pragma unroll 1
			for (int m = 0; m < 4; m++) {
				const int cbm = m == 0 ? 0 : m == 1 ? cU : m == 2 ? cC : cR;
				float h[4][4];
				layer1_forward(h, x, sm, pl, m, lane, t);
				uint32_t bits = 0u;
#pragma unroll
				for (int nt = 0; nt < 4; nt++)
#pragma unroll
					for (int e = 0; e < 4; e++)
						if (h[nt][e] > 0.f) bits |= 1u << (4 * nt + e);
				relu |= (unsigned long long)bits << (16 * m);
				layer2_forward(h, sm, pl, OUT, S, cbm, out_count(m, k), lane, g, t);
			}
			__syncwarp();

			// ---- output activations backward; lanes = (anchor, offset) pairs ----------------------------------------
			for (int idx0 = 0; idx0 < kTile * k; idx0 += 32) {
				const BwdPair nx = load_bwd_pair(idx0 + 32 + lane);   // the next round's loads fly during this round's arithmetic
				const int j = cur.j;
				float *row = OUT + cur.aa * S;
				float d_op = 0.f, d_unc = 0.f, d_col[3] = {0.f, 0.f, 0.f}, d_sr[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
				float pp[kPP] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
				if (cur.valid) {
					const float *gs = TAB + cur.aa * kTab + 6;
					const float nop = tanhf(row[j]);
					float g_op = cur.g_nop;
					if (cur.keep) {
						g_op += cur.g_opa;
						if (a.d_uncertainty) {
							const float s = sigmoidf_(row[cU + j]);
							d_unc = cur.g_unc * s * (1.f - s);
						}
						float sr[7];
#pragma unroll
						for (int c = 0; c < 7; c++) sr[c] = row[cC + 7 * j + c];
#pragma unroll
						for (int c = 0; c < 3; c++) {
							if (a.d_color) {
								const float s = sigmoidf_(row[cR + 3 * j + c]);
								d_col[c] = cur.g_col[c] * s * (1.f - s);
							}
							if (a.d_scaling) {
								const float s = sigmoidf_(sr[c]);
								const float gsc = cur.g_scl[c];
								d_sr[c] = gsc * gs[3 + c] * s * (1.f - s);
								pp[6 + c] = gsc * s;                      // d get_scaling[:, 3 + c]
							}
							if (a.d_xyz) {
								pp[c] = cur.g_xyz[c];                     // d anchor
								pp[3 + c] = cur.g_xyz[c] * cur.off[c];    // d get_scaling[:, c]
							}
						}
						if (a.d_rot) {
							// r = v / max(|v|, eps):  dv = (g - r (r . g)) / max(|v|, eps)   (|v| > eps branch; below it dv = g / eps)
							const float n2 = sr[3] * sr[3] + sr[4] * sr[4] + sr[5] * sr[5] + sr[6] * sr[6];
							const float nrm = sqrtf(n2);
							float dot = 0.f;
#pragma unroll
							for (int c = 0; c < 4; c++) dot += cur.g_rot[c] * sr[3 + c];
							if (nrm > 1e-12f) {
#pragma unroll
								for (int c = 0; c < 4; c++) d_sr[3 + c] = (cur.g_rot[c] - sr[3 + c] * dot / n2) / nrm;
							} else {
#pragma unroll
								for (int c = 0; c < 4; c++) d_sr[3 + c] = cur.g_rot[c] / 1e-12f;
							}
						}
						if (a.d_xyz) {
#pragma unroll
							for (int c = 0; c < 3; c++) a.g_offset[((size_t)cur.aid * k + j) * 3 + c] = cur.g_xyz[c] * gs[c];
						}
					}
					d_op = g_op * (1.f - nop * nop);
				}
				__syncwarp();
				if (cur.on) {
					// gradients of the pre-activations replace the pre-activations (all reads of this pair's slots are done)
					row[j] = d_op;
					row[cU + j] = d_unc;
#pragma unroll
					for (int c = 0; c < 7; c++) row[cC + 7 * j + c] = d_sr[c];
#pragma unroll
					for (int c = 0; c < 3; c++) row[cR + 3 * j + c] = d_col[c];
#pragma unroll
					for (int c = 0; c < kPP; c++) PP[(idx0 + lane) * kPP + c] = pp[c];
				}
				cur = nx;
			}
			__syncwarp();

			// ---- layer 2 backward (gated by the relu of layer 1), layer 1 backward: registers ----------------------------
			float dx[5][4], dxc[5][4];
#pragma unroll
			for (int nt = 0; nt < 5; nt++) {
				zero4(dx[nt]);
				zero4(dxc[nt]);
			}
#pragma unroll 1
			for (int m = 0; m < 4; m++) {
				const int cbm = m == 0 ? 0 : m == 1 ? cU : m == 2 ? cC : cR;
				float dh[4][4];
				layer2_backward(dh, sm, pl, OUT, S, cbm, out_count(m, k), lane, g, t);
				const uint32_t bits = (uint32_t)(relu >> (16 * m));
#pragma unroll
				for (int nt = 0; nt < 4; nt++)
#pragma unroll
					for (int e = 0; e < 4; e++)
						if (!((bits >> (4 * nt + e)) & 1u)) dh[nt][e] = 0.f;
				layer1_backward(dx, dxc, dh, sm, pl, m, lane);
			}
#pragma unroll
			for (int nt = 0; nt < 5; nt++) add4(dx[nt], dxc[nt]);
			// d feat: lane t holds features 8t .. 8t+7 of rows g (c0, c1) and g+8 (c2, c3)
			if (v0) {
				float4 *o = reinterpret_cast<float4 *>(a.g_feat + (size_t)id0 * kHid + 8 * t);
				o[0] = make_float4(dx[0][0], dx[0][1], dx[1][0], dx[1][1]);
				o[1] = make_float4(dx[2][0], dx[2][1], dx[3][0], dx[3][1]);
			}
			if (v1) {
				float4 *o = reinterpret_cast<float4 *>(a.g_feat + (size_t)id1 * kHid + 8 * t);
				o[0] = make_float4(dx[0][2], dx[0][3], dx[1][2], dx[1][3]);
				o[1] = make_float4(dx[2][2], dx[2][3], dx[3][2], dx[3][3]);
			}
			// gradients of (ux, uy, uz, dist): tile 4, columns 0..3 = lanes t < 2
			if (t < 2) {
				X[g * kSX + 36 + 2 * t] = dx[4][0];
				X[g * kSX + 37 + 2 * t] = dx[4][1];
				X[(g + 8) * kSX + 36 + 2 * t] = dx[4][2];
				X[(g + 8) * kSX + 37 + 2 * t] = dx[4][3];
			}
			__syncwarp();
			// d anchor (3) and d get_scaling (6) per anchor: lanes = (anchor, component)
			for (int idx = lane; idx < kTile * kPP; idx += 32) {
				const int aa = idx / kPP, c = idx - aa * kPP;
				const float *xr = X + aa * kSX;
				const int id = __float_as_int(xr[40]);
				if (id < 0) continue;
				// sum this anchor's per-offset partials: [0..2] d anchor (from xyz), [3..8] d get_scaling
				float s = 0.f;
				for (int j = 0; j < k; j++) s += PP[(aa * k + j) * kPP + c];
				if (c < 3) {
					// view / distance path (gaussian_renderer/__init__.py:31-35): v = anchor - cam, dist = |v|, u = v / dist
					const float ux = xr[32], uy = xr[33], uz = xr[34], dist = xr[35];
					const float gux = xr[36], guy = xr[37], guz = xr[38], gdist = xr[39];
					const float u = c == 0 ? ux : c == 1 ? uy : uz;
					const float gu = c == 0 ? gux : c == 1 ? guy : guz;
					const float udot = ux * gux + uy * guy + uz * guz;
					s += (gu - u * udot) / dist + u * gdist;
					a.g_anchor[(size_t)id * 3 + c] = s;
				} else {
					a.g_scaling[(size_t)id * 6 + (c - 3)] = s;
				}
			}
		}
		__syncthreads(); // every tile's X / dOUT of this CTA iteration is complete

		// ==================================================== part B ====================================================
		const int live = min(tiles, ntiles - first);              // tiles of this iteration that exist (<= 0: none)
		for (int w = 0; w < live; w++) {
			// both 8-anchor halves of tile w at once: independent chains for the tensor pipe
			const float *Xt = sm + pl.warp0 + w * pl.per_warp + pl.x;
			const float *Dt = sm + pl.warp0 + w * pl.per_warp + pl.out;
			// H^T = relu(W1 x^T + b1): rows h = 16hh + g (+8), columns = anchors 2t (+1) of the n-tile
			float hT[2][4] = {{b1lo, b1lo, b1hi, b1hi}, {b1lo, b1lo, b1hi, b1hi}}, hC[2][4];
			zero4(hC[0]);
			zero4(hC[1]);
#pragma unroll
			for (int ks = 0; ks < 5; ks++) {
				const float4 e = *reinterpret_cast<const float4 *>(sm + pl.w1f + (((mB * 5 + ks) * 2 + hh) * 32 + lane) * 4);
				const AFrag wa = make_a(e.x, e.z, e.y, e.w);
#pragma unroll
				for (int sub = 0; sub < 2; sub++) {
					const float *xr = Xt + (8 * sub + g) * kSX;
					const float xb0 = ks < 4 ? xr[8 * t + ks] : xr[32 + t];
					const float xb1 = ks < 4 ? xr[8 * t + 4 + ks] : 0.f;
					mma3(hT[sub], hC[sub], wa, xb0, xb1);
				}
			}
			// dH^T = W2^T dOUT^T, gated
			float dT[2][4], dC[2][4];
			zero4(dT[0]); zero4(dT[1]); zero4(dC[0]); zero4(dC[1]);
			for (int ks = 0; ks < nksB; ks++) {
				const float4 e = *reinterpret_cast<const float4 *>(sm + pl.w2b + (((cbB >> 3) + ks) * 2 + hh) * 128 + lane * 4);
				const AFrag wa = make_a(e.x, e.z, e.y, e.w);
#pragma unroll
				for (int sub = 0; sub < 2; sub++) {
					const float *dr = Dt + (8 * sub + g) * S + cbB + 8 * ks + t;
					mma3(dT[sub], dC[sub], wa, dr[0], dr[4]);
				}
			}
#pragma unroll
			for (int sub = 0; sub < 2; sub++) {
#pragma unroll
				for (int e = 0; e < 4; e++) {
					const float hv = hT[sub][e] + hC[sub][e];
					dT[sub][e] = hv > 0.f ? dT[sub][e] + dC[sub][e] : 0.f;
					hT[sub][e] = fmaxf(hv, 0.f);
				}
				accb1[0] += dT[sub][0] + dT[sub][1];
				accb1[1] += dT[sub][2] + dT[sub][3];
			}
			// dW1 += dH^T x   (A = the dT fragment, contraction = the 8 anchors; B from the X tile)
#pragma unroll
			for (int sub = 0; sub < 2; sub++) {
				const AFrag da = c_as_a(dT[sub]);
				const float *x0 = Xt + (8 * sub + 2 * t) * kSX + g, *x1 = x0 + kSX;
#pragma unroll
				for (int nt = 0; nt < 5; nt++) mma3(acc1[nt], da, x0[8 * nt], x1[8 * nt]);
			}
			// dW2 += dOUT^T H   (A from the dOUT tile: rows o = 16mt + g (+8), contraction = anchors; B = the hT fragment)
#pragma unroll
			for (int sub = 0; sub < 2; sub++) {
				const float *d0 = Dt + (8 * sub + 2 * t) * S + cbB + g, *d1 = d0 + S;
#pragma unroll
				for (int mt = 0; mt < KMT; mt++) {
					if (mt < nmtB) {
						const float a0 = d0[16 * mt], a1 = d0[16 * mt + 8], a2 = d1[16 * mt], a3 = d1[16 * mt + 8];
						const AFrag oa = make_a(a0, a1, a2, a3);
						mma3(acc2[mt][0], oa, hT[sub][0], hT[sub][1]);
						mma3(acc2[mt][1], oa, hT[sub][2], hT[sub][3]);
						accb2[mt][0] += a0 + a2;
						accb2[mt][1] += a1 + a3;
					}
				}
			}
		}
		__syncthreads(); // tiles consumed before the next iteration overwrites them
	}

	// ---- flush: one atomic per owned entry (torch layouts: w1[h][i], b1[h], w2[o][h], b2[o]) ---------------------
#pragma unroll
	for (int nt = 0; nt < 5; nt++)
#pragma unroll
		for (int e = 0; e < 4; e++) {
			const int h = 16 * hh + g + 8 * (e >> 1), i = 8 * nt + 2 * t + (e & 1);
			if (i < kIn) atomicAdd(a.g_w1[mB] + h * kIn + i, acc1[nt][e]);
		}
#pragma unroll
	for (int q = 0; q < 2; q++) {
		float v = accb1[q];
		v += __shfl_xor_sync(kFull, v, 1);
		v += __shfl_xor_sync(kFull, v, 2);
		if (t == 0) atomicAdd(a.g_b1[mB] + 16 * hh + g + 8 * q, v);
	}
#pragma unroll
	for (int mt = 0; mt < KMT; mt++) {
		if (mt < nmtB) {
#pragma unroll
			for (int nt = 0; nt < 2; nt++)
#pragma unroll
				for (int e = 0; e < 4; e++) {
					const int o = 16 * mt + g + 8 * (e >> 1), h = 16 * hh + 8 * nt + 2 * t + (e & 1);
					if (o < nB) atomicAdd(a.g_w2[mB] + o * kHid + h, acc2[mt][nt][e]);
				}
#pragma unroll
			for (int q = 0; q < 2; q++) {
				float v = accb2[mt][q];
				v += __shfl_xor_sync(kFull, v, 1);
				v += __shfl_xor_sync(kFull, v, 2);
				const int o = 16 * mt + g + 8 * q;
				if (hh == 0 && t == 0 && o < nB) atomicAdd(a.g_b2[mB] + o, v);
			}
		}
	}
}

// ---- densification statistics (scene/gaussian_model.py:729-757, GaussianModel.training_statis) --------------------------------
// One thread per anchor.  vis_incl / sel_incl are inclusive scans of the anchor-visibility mask and of the offset selection
// mask (neural_opacity > 0) — the same bookkeeping the reference does with boolean-mask assignments:
//   opacity_accum[a]            += sum_j max(neural_opacity[r, j], 0)           for visible anchors (r = visible rank)
//   anchor_demon[a]             += 1
//   offset_gradient_accum[a, j] += |viewspace_grad[p, :2]|,  offset_denom[a, j] += 1   for selected offsets whose Gaussian p has radii > 0
// n_vis / P: row counts of neural_opacity (n_vis * k) and of update_filter / viewspace_grad.  A mask whose popcount disagrees with
// them (the reference raises a shape mismatch from its boolean-mask assignment there) must not index out of bounds: such rows
// are skipped.
__global__ void __launch_bounds__(256) training_statis_kernel(int A, int k, int64_t n_vis, int64_t P, const uint8_t *__restrict__ anchor_visible, const uint32_t *__restrict__ vis_incl,
                                                              const uint8_t *__restrict__ offset_selected, const uint32_t *__restrict__ sel_incl,
                                                              const uint8_t *__restrict__ update_filter, const float *__restrict__ neural_opacity,
                                                              const float *__restrict__ viewspace_grad, float *__restrict__ opacity_accum,
                                                              float *__restrict__ anchor_demon, float *__restrict__ offset_gradient_accum,
                                                              float *__restrict__ offset_denom)
{
	const int a = blockIdx.x * blockDim.x + threadIdx.x;
	if (a >= A || !anchor_visible[a]) return;
	const size_t r = (size_t)vis_incl[a] - 1;
	if ((int64_t)r >= n_vis) return;
	float osum = 0.f;
	for (int j = 0; j < k; j++) {
		const size_t t = r * k + j;
		const float o = neural_opacity[t];
		osum += o < 0.f ? 0.f : o;
		if (offset_selected[t]) {
			const size_t p = (size_t)sel_incl[t] - 1;
			if ((int64_t)p >= P) continue;
			if (update_filter[p]) {
				const float gx = viewspace_grad[p * 3 + 0], gy = viewspace_grad[p * 3 + 1];
				offset_gradient_accum[(size_t)a * k + j] += sqrtf(gx * gx + gy * gy);
				offset_denom[(size_t)a * k + j] += 1.f;
			}
		}
	}
	opacity_accum[a] += osum;
	anchor_demon[a] += 1.f;
}

// persistent CTAs per SM of the two forward kernels (<= 100 KB of shared memory and <= 128 registers each)
#ifndef GSR_DEC_FWD_CTAS
#define GSR_DEC_FWD_CTAS 2
#endif
constexpr size_t kMaxSmemBytes = 232448;   // 227 KB opt-in limit per CTA on sm_100

int sm_count()   // of the CURRENT device (asked per call: a process may drive several devices)
{
	int dev = 0, sms = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
	return sms;
}
// persistent grid: enough CTAs to cover the tiles (`tiles` per CTA iteration), at most `per_sm` per SM
int grid_for(int anchors, int per_sm, int tiles)
{
	const int ntiles = (anchors + kTile - 1) / kTile;
	const int want = (ntiles + tiles - 1) / tiles;
	return max(1, min(want, per_sm * sm_count()));
}

template <typename K>
cudaError_t set_smem(K kernel, size_t bytes)
{
	return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
// resident CTAs per SM of a persistent forward kernel with this much shared memory (2 at k <= 10, 1 above)
template <typename K>
int resident_ctas(K kernel, size_t smem, int cap)
{
	int n = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 256, smem) != cudaSuccess || n < 1) n = 1;
	return min(n, cap);
}

} // namespace

DecodeLayout decode_layout(int A)
{
	DecodeLayout L{};
	size_t off = 0;
	const size_t n = (size_t)(A > 0 ? A : 1) * 4;
	L.vis_incl = off;   off += align_up(n);
	L.vis_ids = off;    off += align_up(n);
	L.count = off;      off += align_up(n);
	L.maskbits = off;   off += align_up(n);
	L.gauss_incl = off; off += align_up(n);
	L.scan_tmp = off;   off += scan_scratch_bytes(A > 0 ? A : 1);
	L.total = off;
	return L;
}

cudaError_t decode_stage1(int A, int k, const float *anchor, const float *feat, const uint8_t *visible_mask, const float *campos,
                          const DecodeWeights &wt, char *scratch, const DecodeLayout &L, float *neural_opacity, uint8_t *mask,
                          int64_t *counts_host, cudaStream_t stream)
{
	cudaError_t e;
	uint32_t *incl = (uint32_t *)(scratch + L.vis_incl), *ids = (uint32_t *)(scratch + L.vis_ids);
	uint32_t *count = (uint32_t *)(scratch + L.count), *bits = (uint32_t *)(scratch + L.maskbits), *gincl = (uint32_t *)(scratch + L.gauss_incl);
	if (visible_mask) {   // visible ranks (inclusive scan of the mask) and the ascending list of visible anchors
		if ((e = inclusive_sum_mask(visible_mask, incl, ids, A, scratch + L.scan_tmp, stream)) != cudaSuccess) return e;
	}
	DecodeArgs a{};
	a.k = k; a.n_vis = A; a.n_vis_dev = visible_mask ? incl + (A - 1) : nullptr;
	a.vis_ids = visible_mask ? ids : nullptr;
	a.anchor = anchor; a.feat = feat; a.campos = campos; a.wt = wt;
	a.neural_opacity = neural_opacity; a.mask = mask; a.count = count; a.maskbits = bits;
	const size_t smem = (size_t)smem_plan(k, 0, 1, false, kWarps).total * 4;
	if ((e = set_smem(decode_opacity_kernel, smem)) != cudaSuccess) return e;
	decode_opacity_kernel<<<grid_for(A, resident_ctas(decode_opacity_kernel, smem, GSR_DEC_S1_CTAS), kWarps), 256, smem, stream>>>(a);
	count_launch();
	if ((e = cudaGetLastError()) != cudaSuccess) return e;
	if ((e = inclusive_sum_gather(count, nullptr, gincl, A, scratch + L.scan_tmp, stream)) != cudaSuccess) return e;
	// counts_host[0] = n_vis, [1] = P: 4 bytes each into the low halves of pre-zeroed little-endian int64s
	if (visible_mask) {
		if ((e = cudaMemcpyAsync(&counts_host[0], incl + (A - 1), 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
	}
	return cudaMemcpyAsync(&counts_host[1], gincl + (A - 1), 4, cudaMemcpyDeviceToHost, stream);
}

cudaError_t decode_stage2(const DecodeArgs &a, cudaStream_t stream)
{
	cudaError_t e;
	const size_t smem = (size_t)smem_plan(a.k, 1, 4, false, kWarps).total * 4;
	if ((e = set_smem(decode_outputs_kernel, smem)) != cudaSuccess) return e;
	decode_outputs_kernel<<<grid_for(a.n_vis, resident_ctas(decode_outputs_kernel, smem, GSR_DEC_FWD_CTAS), kWarps), 256, smem, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}

cudaError_t decode_backward(const DecodeBwdArgs &a, cudaStream_t stream)
{
	cudaError_t e;
	// persistent, one CTA per SM.  Eight tiles (128 anchors) per CTA iteration while their X / OUT / partial tiles fit next to
	// the four weight copies (k <= 10: 217 KB), four above that.
	const int tiles = (size_t)smem_plan(a.f.k, 0, 4, true, kWarps).total * 4 <= kMaxSmemBytes ? kWarps : kWarps / 2;
	const size_t smem = (size_t)smem_plan(a.f.k, 0, 4, true, tiles).total * 4;
	if (smem > kMaxSmemBytes) return cudaErrorInvalidConfiguration;
	const int grid = grid_for(a.f.n_vis, 1, tiles);
	if ((7 * a.f.k + 15) / 16 <= 5) {
		if ((e = set_smem(decode_backward_kernel<5>, smem)) != cudaSuccess) return e;
		decode_backward_kernel<5><<<grid, 256, smem, stream>>>(a, tiles);
	} else {
		if ((e = set_smem(decode_backward_kernel<(7 * kDecMaxK + 15) / 16>, smem)) != cudaSuccess) return e;
		decode_backward_kernel<(7 * kDecMaxK + 15) / 16><<<grid, 256, smem, stream>>>(a, tiles);
	}
	count_launch();
	return cudaGetLastError();
}

size_t statis_scratch_bytes(int A, int k)
{
	const size_t a = (size_t)(A > 0 ? A : 1), n = a * (size_t)(k > 0 ? k : 1);
	return align_up(a * 4) + align_up(n * 4) + scan_scratch_bytes((int64_t)n);   // the two inclusive scans + the scan's own scratch
}

cudaError_t training_statis(int A, int k, int64_t n_vis, int64_t P, const uint8_t *anchor_visible, const uint8_t *offset_selected, const uint8_t *update_filter,
                            const float *neural_opacity, const float *viewspace_grad, float *opacity_accum, float *anchor_demon,
                            float *offset_gradient_accum, float *offset_denom, char *scratch, cudaStream_t stream)
{
	const size_t a = (size_t)A, n = (size_t)n_vis * k;
	uint32_t *vincl = (uint32_t *)scratch;
	uint32_t *sincl = (uint32_t *)(scratch + align_up(a * 4));
	char *tmp = scratch + align_up(a * 4) + align_up((size_t)A * k * 4);
	cudaError_t e;
	if ((e = inclusive_sum_mask(anchor_visible, vincl, nullptr, A, tmp, stream)) != cudaSuccess) return e;
	if (n > 0) {
		if ((e = inclusive_sum_mask(offset_selected, sincl, nullptr, (int64_t)n, tmp, stream)) != cudaSuccess) return e;
	}
	training_statis_kernel<<<(A + 255) / 256, 256, 0, stream>>>(A, k, n_vis, P, anchor_visible, vincl, offset_selected, sincl, update_filter, neural_opacity,
	                                                          viewspace_grad, opacity_accum, anchor_demon, offset_gradient_accum, offset_denom);
	count_launch();
	return cudaGetLastError();
}

} // namespace gsr
