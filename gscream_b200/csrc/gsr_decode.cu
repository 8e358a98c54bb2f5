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
// Work decomposition: one warp decodes kDecNA = 4 anchors at a time; lane = hidden unit in layer 1 (the hidden
// width, the feature width and the warp width are all 32), lane = output unit in layer 2; weights live in shared
// memory in the layouts that make both conflict-free.  The arithmetic is plain fp32 FMA; the whole decode is
// ~8.4 kMAC per anchor, far below the rasterizer's cost, so no tensor cores are used here.
#include "gsr_internal.cuh"
#include "gsr_sort.cuh"
#include "gsr_decode.cuh"

namespace gsr {

namespace {

constexpr int kHid = 32;          // hidden width == feat_dim == warp width
constexpr int kIn = 36;           // feat_dim + 3 + 1
constexpr int kNA = kDecNA;       // anchors per warp iteration
constexpr int kWarps = 8;
constexpr int kHStride = 33;      // H4[m][h] rows padded so that lanes of different MLPs hit different banks

__host__ __device__ inline int out_base(int m, int k) { return m == 0 ? 0 : m == 1 ? k : m == 2 ? 2 * k : 9 * k; }
__host__ __device__ inline int out_count(int m, int k) { return m == 0 ? k : m == 1 ? k : m == 2 ? 7 * k : 3 * k; }
__device__ __forceinline__ int mlp_of(int o, int k) { return o < k ? 0 : (o < 2 * k ? 1 : (o < 9 * k ? 2 : 3)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// shared-memory carve-up (in floats).  One padded copy of each weight matrix serves both access patterns:
//   W1 as [m][i][33]  : forward layer 1 reads a row with lane = h (consecutive), backward layer 1 a column with lane = i (stride 33)
//   W2 as [h][Os], Os odd >= 12k : forward layer 2 reads with lane = o (consecutive), backward layer 2 with lane = h (stride Os)
constexpr int kW1Stride = kHid + 1;
struct SmemPlan {
	int w1, b1, w2, os, b2, warp0, per_warp, x4, h4, dh4, out4, pp, total;
};
__host__ __device__ inline SmemPlan smem_plan(int k, bool backward)
{
	const int O = 12 * k;
	const int Opad = (O + 3) & ~3;
	SmemPlan p{};
	int off = 0;
	p.w1 = off;  off += 4 * kIn * kW1Stride;
	p.b1 = off;  off += 4 * kHid;
	p.os = O | 1;
	p.w2 = off;  off += kHid * p.os;
	p.b2 = off;  off += Opad;
	off = (off + 3) & ~3;
	p.warp0 = off;
	int w = 0;
	p.x4 = w;   w += kIn * kNA;                         // [i][a]
	p.h4 = w;   w += 4 * kHStride * kNA;                // [m][h(+1)][a]
	p.dh4 = w;  if (backward) w += 4 * kHStride * kNA;
	p.out4 = w; w += Opad * kNA;                        // [o][a]  (pre-activations, then their gradients in place)
	p.pp = w;   if (backward) w += ((kNA * k * 10 + 3) & ~3); // per (anchor, offset) partials: dxyz(3) dgs(6)
	p.per_warp = w;
	p.total = off + kWarps * w;
	return p;
}

struct GroupIO {
	int id[kNA];
	bool valid[kNA];
	float ax[kNA], ay[kNA], az[kNA];   // anchor position
	float ux[kNA], uy[kNA], uz[kNA];   // normalised view direction
	float dist[kNA];
};

__device__ __forceinline__ void load_weights(float *sm, const SmemPlan &pl, const DecodeWeights &wt, int k, int tid)
{
	const int O = 12 * k;
	for (int idx = tid; idx < 4 * kHid * kIn; idx += 256) {
		const int m = idx / (kHid * kIn), rem = idx % (kHid * kIn), h = rem / kIn, i = rem % kIn;
		sm[pl.w1 + (m * kIn + i) * kW1Stride + h] = __ldg(wt.w1[m] + rem);
	}
	for (int idx = tid; idx < 4 * kHid; idx += 256) sm[pl.b1 + idx] = __ldg(wt.b1[idx >> 5] + (idx & 31));
	for (int idx = tid; idx < O * kHid; idx += 256) {
		const int o = idx / kHid, h = idx % kHid;
		const int m = o < k ? 0 : (o < 2 * k ? 1 : (o < 9 * k ? 2 : 3));
		sm[pl.w2 + h * pl.os + o] = __ldg(wt.w2[m] + (o - out_base(m, k)) * kHid + h);
	}
	for (int o = tid; o < O; o += 256) {
		const int m = o < k ? 0 : (o < 2 * k ? 1 : (o < 9 * k ? 2 : 3));
		sm[pl.b2 + o] = __ldg(wt.b2[m] + (o - out_base(m, k)));
	}
}

// One group of kNA visible anchors as it comes from global memory.  Fetched one loop iteration ahead of its use, so that the
// two dependent loads (visible list -> anchor row / feature row) are off the critical path of the MLP arithmetic.
struct GroupRaw {
	int id[kNA];
	bool valid[kNA];
	float f[kNA];                      // feature `lane` of each anchor
	float ax[kNA], ay[kNA], az[kNA];
};
__device__ __forceinline__ void fetch_group(GroupRaw &r, int g, int groups, int n_vis, const uint32_t *__restrict__ vis_ids,
                                            const float *__restrict__ anchor, const float *__restrict__ feat, int lane)
{
#pragma unroll
	for (int a = 0; a < kNA; a++) {
		const int rank = g * kNA + a;
		r.valid[a] = g < groups && rank < n_vis;
		r.id[a] = r.valid[a] ? (vis_ids ? (int)__ldg(vis_ids + rank) : rank) : 0;
	}
#pragma unroll
	for (int a = 0; a < kNA; a++) {
		r.f[a] = r.ax[a] = r.ay[a] = r.az[a] = 0.f;
		if (r.valid[a]) {
			r.f[a] = __ldg(feat + (size_t)r.id[a] * kHid + lane);
			r.ax[a] = __ldg(anchor + (size_t)r.id[a] * 3 + 0);
			r.ay[a] = __ldg(anchor + (size_t)r.id[a] * 3 + 1);
			r.az[a] = __ldg(anchor + (size_t)r.id[a] * 3 + 2);
		}
	}
}
// Build X4[i][a] = [feat | ob_view | ob_dist] (gaussian_renderer/__init__.py:26-52).  Anchors past n_vis contribute zeros.
__device__ __forceinline__ void stage_group(GroupIO &io, float *X4, const GroupRaw &r, float cx, float cy, float cz, int lane)
{
#pragma unroll
	for (int a = 0; a < kNA; a++) {
		io.valid[a] = r.valid[a];
		io.id[a] = r.id[a];
		io.ax[a] = r.ax[a]; io.ay[a] = r.ay[a]; io.az[a] = r.az[a];
		io.ux[a] = io.uy[a] = io.uz[a] = 0.f;
		io.dist[a] = 0.f;
		if (io.valid[a]) {
			const float vx = io.ax[a] - cx, vy = io.ay[a] - cy, vz = io.az[a] - cz;
			io.dist[a] = sqrtf(vx * vx + vy * vy + vz * vz);
			io.ux[a] = vx / io.dist[a];
			io.uy[a] = vy / io.dist[a];
			io.uz[a] = vz / io.dist[a];
		}
		X4[lane * kNA + a] = r.f[a];
		if (lane < 4) X4[(kHid + lane) * kNA + a] = lane == 0 ? io.ux[a] : lane == 1 ? io.uy[a] : lane == 2 ? io.uz[a] : io.dist[a];
	}
	__syncwarp();
}

// Layer 1 of MLPs m0..m1-1: lane = hidden unit.  H4[m][lane][a] = relu(b1 + sum_i W1[lane][i] x[i][a]).
template <int M0, int M1>
__device__ __forceinline__ void layer1(const float *sm, const SmemPlan &pl, const float *X4, float *H4, int lane)
{
	float acc[M1 - M0][kNA];
#pragma unroll
	for (int m = M0; m < M1; m++) {
		const float b = sm[pl.b1 + m * kHid + lane];
#pragma unroll
		for (int a = 0; a < kNA; a++) acc[m - M0][a] = b;
	}
#pragma unroll 4
	for (int i = 0; i < kIn; i++) {
		const float4 x = *reinterpret_cast<const float4 *>(X4 + i * kNA);
#pragma unroll
		for (int m = M0; m < M1; m++) {
			const float w = sm[pl.w1 + (m * kIn + i) * kW1Stride + lane];
			acc[m - M0][0] = fmaf(w, x.x, acc[m - M0][0]);
			acc[m - M0][1] = fmaf(w, x.y, acc[m - M0][1]);
			acc[m - M0][2] = fmaf(w, x.z, acc[m - M0][2]);
			acc[m - M0][3] = fmaf(w, x.w, acc[m - M0][3]);
		}
	}
#pragma unroll
	for (int m = M0; m < M1; m++) {
		float4 h = {fmaxf(acc[m - M0][0], 0.f), fmaxf(acc[m - M0][1], 0.f), fmaxf(acc[m - M0][2], 0.f), fmaxf(acc[m - M0][3], 0.f)};
		*reinterpret_cast<float4 *>(H4 + (m * kHStride + lane) * kNA) = h;
	}
	__syncwarp();
}

// Layer 2 for outputs [o_begin, o_end): lane = output unit.  OUT4[o][a] = b2[o] + sum_h W2[o][h] H[m(o)][h][a].
__device__ __forceinline__ void layer2(const float *sm, const SmemPlan &pl, const float *H4, float *OUT4, int k, int o_begin, int o_end, int lane)
{
	for (int o0 = o_begin; o0 < o_end; o0 += 32) {
		const int o = o0 + lane;
		const bool on = o < o_end;
		const int oc = on ? o : o_begin;
		const int m = mlp_of(oc, k);
		const float b = sm[pl.b2 + oc];
		float acc0 = b, acc1 = b, acc2 = b, acc3 = b;
		const float *hrow = H4 + m * kHStride * kNA;
#pragma unroll 8
		for (int h = 0; h < kHid; h++) {
			const float w = sm[pl.w2 + h * pl.os + oc];
			const float4 hv = *reinterpret_cast<const float4 *>(hrow + h * kNA);
			acc0 = fmaf(w, hv.x, acc0);
			acc1 = fmaf(w, hv.y, acc1);
			acc2 = fmaf(w, hv.z, acc2);
			acc3 = fmaf(w, hv.w, acc3);
		}
		if (on) *reinterpret_cast<float4 *>(OUT4 + o * kNA) = make_float4(acc0, acc1, acc2, acc3);
	}
	__syncwarp();
}

// ---- stage 1: opacity MLP, mask, per-anchor counts ---------------------------------------------------------------
__global__ void __launch_bounds__(256) decode_opacity_kernel(DecodeArgs a)
{
	extern __shared__ __align__(16) float sm[];
	const int k = a.k;
	const SmemPlan pl = smem_plan(k, false);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	load_weights(sm, pl, a.wt, k, tid);
	__syncthreads();
	float *X4 = sm + pl.warp0 + warp * pl.per_warp + pl.x4;
	float *H4 = sm + pl.warp0 + warp * pl.per_warp + pl.h4;
	float *OUT4 = sm + pl.warp0 + warp * pl.per_warp + pl.out4;
	const int n_vis = a.n_vis_dev ? (int)*a.n_vis_dev : a.n_vis;
	const int groups = (n_vis + kNA - 1) / kNA;
	const float cx = __ldg(a.campos), cy = __ldg(a.campos + 1), cz = __ldg(a.campos + 2);
	GroupRaw raw;
	fetch_group(raw, blockIdx.x * kWarps + warp, groups, n_vis, a.vis_ids, a.anchor, a.feat, lane);
	for (int g = blockIdx.x * kWarps + warp; g < groups; g += gridDim.x * kWarps) {
		GroupIO io;
		stage_group(io, X4, raw, cx, cy, cz, lane);
		fetch_group(raw, g + gridDim.x * kWarps, groups, n_vis, a.vis_ids, a.anchor, a.feat, lane);
		layer1<0, 1>(sm, pl, X4, H4, lane);
		layer2(sm, pl, H4, OUT4, k, 0, k, lane);
		// lanes = (anchor, offset) pairs; k <= 16 so 4 anchors need at most two rounds
		for (int idx = lane; idx < kNA * k; idx += 32) {
			const int aa = idx / k, j = idx % k;
			const int r = g * kNA + aa;
			const bool valid = r < n_vis;
			const float nop = tanhf(OUT4[j * kNA + aa]);
			const bool keep = valid && nop > 0.0f;                       // gaussian_renderer/__init__.py:59
			if (valid) {
				a.neural_opacity[(size_t)r * k + j] = nop;
				a.mask[(size_t)r * k + j] = keep ? 1 : 0;
			}
		}
		__syncwarp();
		// per-anchor count and bit mask: lanes 0..3 re-read their anchor's k pre-activations (cheap, shared memory)
		if (lane < kNA) {
			const int r = g * kNA + lane;
			if (r < n_vis) {
				uint32_t bits = 0;
				for (int j = 0; j < k; j++)
					if (tanhf(OUT4[j * kNA + lane]) > 0.0f) bits |= 1u << j;
				a.count[r] = (uint32_t)__popc(bits);
				a.maskbits[r] = bits;
			}
		}
		__syncwarp();
	}
}

// ---- stage 2: the other three MLPs + post-processing into the compacted outputs ------------------------------------
__global__ void __launch_bounds__(256) decode_outputs_kernel(DecodeArgs a)
{
	extern __shared__ __align__(16) float sm[];
	const int k = a.k;
	const SmemPlan pl = smem_plan(k, false);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	load_weights(sm, pl, a.wt, k, tid);
	__syncthreads();
	float *X4 = sm + pl.warp0 + warp * pl.per_warp + pl.x4;
	float *H4 = sm + pl.warp0 + warp * pl.per_warp + pl.h4;
	float *OUT4 = sm + pl.warp0 + warp * pl.per_warp + pl.out4;
	const int n_vis = a.n_vis_dev ? (int)*a.n_vis_dev : a.n_vis; // (device copy when the host has not read the counts yet)
	const int groups = (n_vis + kNA - 1) / kNA;
	const float cx = __ldg(a.campos), cy = __ldg(a.campos + 1), cz = __ldg(a.campos + 2);
	GroupRaw raw;
	fetch_group(raw, blockIdx.x * kWarps + warp, groups, n_vis, a.vis_ids, a.anchor, a.feat, lane);
	for (int g = blockIdx.x * kWarps + warp; g < groups; g += gridDim.x * kWarps) {
		GroupIO io;
		stage_group(io, X4, raw, cx, cy, cz, lane);
		fetch_group(raw, g + gridDim.x * kWarps, groups, n_vis, a.vis_ids, a.anchor, a.feat, lane);
		layer1<1, 4>(sm, pl, X4, H4, lane);
		layer2(sm, pl, H4, OUT4, k, k, 12 * k, lane);
		for (int idx = lane; idx < kNA * k; idx += 32) {
			const int aa = idx / k, j = idx % k;
			const int r = g * kNA + aa;
			if (r >= n_vis) continue;
			const uint32_t bits = __ldg(a.maskbits + r);
			if (!((bits >> j) & 1u)) continue;
			const int id = a.vis_ids ? (int)__ldg(a.vis_ids + r) : r;
			const size_t p = (size_t)(__ldg(a.gauss_incl + r) - (uint32_t)__popc(bits)) + __popc(bits & ((1u << j) - 1u));
			const float *gs = a.scaling + (size_t)id * 6;
			const float *off = a.offset + ((size_t)id * k + j) * 3;
			const float *an = a.anchor + (size_t)id * 3;
			a.out_opacity[p] = __ldg(a.neural_opacity + (size_t)r * k + j);
			a.out_uncertainty[p] = sigmoidf_(OUT4[(k + j) * kNA + aa]);
			float sr[7];
#pragma unroll
			for (int c = 0; c < 7; c++) sr[c] = OUT4[(2 * k + 7 * j + c) * kNA + aa];
#pragma unroll
			for (int c = 0; c < 3; c++) {
				a.out_color[p * 3 + c] = sigmoidf_(OUT4[(9 * k + 3 * j + c) * kNA + aa]);
				a.out_scaling[p * 3 + c] = __ldg(gs + 3 + c) * sigmoidf_(sr[c]);                 // :89
				a.out_xyz[p * 3 + c] = __ldg(an + c) + __ldg(off + c) * __ldg(gs + c);          // :93-94
			}
			// F.normalize (scene/gaussian_model.py:52): v / max(||v||, 1e-12)
			const float nrm = fmaxf(sqrtf(sr[3] * sr[3] + sr[4] * sr[4] + sr[5] * sr[5] + sr[6] * sr[6]), 1e-12f);
#pragma unroll
			for (int c = 0; c < 4; c++) a.out_rot[p * 4 + c] = sr[3 + c] / nrm;
		}
		__syncwarp();
	}
}

// ---- backward -------------------------------------------------------------------------------------------------------
// Per (warp, group of 4 anchors): recompute activations, turn the upstream gradients of the kept Gaussians into
// gradients of the 12k output pre-activations (DOUT4, in place over OUT4), back-propagate through layer 2 (lane = h) and
// layer 1 (lane = input i), write d_feat / d_anchor / d_offset / d_scaling.  Per CTA iteration (8 groups = 32 anchors):
// every thread adds its slice of the weight-gradient outer products into registers; one atomic per entry at the end.
constexpr int kAcc2 = (7 * kDecMaxK + 3) / 4;   // output rows of W2 owned by one thread (cov MLP over 4 warps)

__global__ void __launch_bounds__(256, 2) decode_backward_kernel(DecodeBwdArgs a)
{
	extern __shared__ __align__(16) float sm[];
	const int k = a.f.k, O = 12 * k;
	const SmemPlan pl = smem_plan(k, true);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	load_weights(sm, pl, a.f.wt, k, tid);
	__syncthreads();
	float *X4 = sm + pl.warp0 + warp * pl.per_warp + pl.x4;
	float *H4 = sm + pl.warp0 + warp * pl.per_warp + pl.h4;
	float *DH4 = sm + pl.warp0 + warp * pl.per_warp + pl.dh4;
	float *OUT4 = sm + pl.warp0 + warp * pl.per_warp + pl.out4;
	float *PP = sm + pl.warp0 + warp * pl.per_warp + pl.pp;
	const int n_vis = a.f.n_vis;
	const int groups = (n_vis + kNA - 1) / kNA;
	const int cta_iters = (groups + gridDim.x * kWarps - 1) / (gridDim.x * kWarps);
	const float cx = __ldg(a.f.campos), cy = __ldg(a.f.campos + 1), cz = __ldg(a.f.campos + 2);

	// this thread's slice of the weight gradients
	//   W1: (m1, h = lane) x inputs [18 * half, 18 * half + 18)
	const int m1 = (tid >> 5) & 3, half = tid >> 7;
	float acc1[18];
#pragma unroll
	for (int q = 0; q < 18; q++) acc1[q] = 0.f;
	float accb1 = 0.f;
	//   W2: warps 0 / 1 own the opacity / uncertainty rows, warps 2-5 a quarter of the cov rows, warps 6-7 half of the colour rows
	const int m2 = warp == 0 ? 0 : warp == 1 ? 1 : warp < 6 ? 2 : 3;
	const int parts = m2 == 2 ? 4 : m2 == 3 ? 2 : 1, part = m2 == 2 ? warp - 2 : m2 == 3 ? warp - 6 : 0;
	const int rows_m = out_count(m2, k), rows_per = (rows_m + parts - 1) / parts;
	const int o_first = out_base(m2, k) + part * rows_per;
	const int o_cnt = max(0, min(rows_per, rows_m - part * rows_per));
	float acc2[kAcc2];
#pragma unroll
	for (int q = 0; q < kAcc2; q++) acc2[q] = 0.f;
	float accb2 = 0.f;

	GroupRaw raw;
	fetch_group(raw, blockIdx.x * kWarps + warp, groups, n_vis, a.f.vis_ids, a.f.anchor, a.f.feat, lane);
	for (int it = 0; it < cta_iters; it++) {
		const int g = (it * gridDim.x + blockIdx.x) * kWarps + warp;
		GroupIO io;
		stage_group(io, X4, raw, cx, cy, cz, lane); // g >= groups: all invalid
		fetch_group(raw, g + gridDim.x * kWarps, groups, n_vis, a.f.vis_ids, a.f.anchor, a.f.feat, lane);
		layer1<0, 4>(sm, pl, X4, H4, lane);
		layer2(sm, pl, H4, OUT4, k, 0, O, lane);

		// ---- output activations backward; lanes = (anchor, offset) pairs --------------------------------------------
		for (int idx0 = 0; idx0 < kNA * k; idx0 += 32) {
			const int idx = idx0 + lane;
			const bool on = idx < kNA * k;
			const int aa = on ? idx / k : 0, j = on ? idx % k : 0;
			const int r = g * kNA + aa;
			const bool valid = on && g < groups && r < n_vis;
			float d_op = 0.f, d_unc = 0.f, d_col[3] = {0.f, 0.f, 0.f}, d_sr[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
			float pp[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
			if (valid) {
				const uint32_t bits = __ldg(a.f.maskbits + r);
				const bool keep = (bits >> j) & 1u;
				const int aid = a.f.vis_ids ? (int)__ldg(a.f.vis_ids + r) : r;
				const float nop = tanhf(OUT4[j * kNA + aa]);
				float g_op = a.d_neural_opacity ? __ldg(a.d_neural_opacity + (size_t)r * k + j) : 0.f;
				float dxyz[3] = {0.f, 0.f, 0.f};
				if (keep) {
					const size_t p = (size_t)(__ldg(a.f.gauss_incl + r) - (uint32_t)__popc(bits)) + __popc(bits & ((1u << j) - 1u));
					const float *gs = a.f.scaling + (size_t)aid * 6;
					const float *off = a.f.offset + ((size_t)aid * k + j) * 3;
					if (a.d_opacity) g_op += __ldg(a.d_opacity + p);
					if (a.d_uncertainty) {
						const float s = sigmoidf_(OUT4[(k + j) * kNA + aa]);
						d_unc = __ldg(a.d_uncertainty + p) * s * (1.f - s);
					}
					float sr[7];
#pragma unroll
					for (int c = 0; c < 7; c++) sr[c] = OUT4[(2 * k + 7 * j + c) * kNA + aa];
#pragma unroll
					for (int c = 0; c < 3; c++) {
						if (a.d_color) {
							const float s = sigmoidf_(OUT4[(9 * k + 3 * j + c) * kNA + aa]);
							d_col[c] = __ldg(a.d_color + p * 3 + c) * s * (1.f - s);
						}
						if (a.d_scaling) {
							const float s = sigmoidf_(sr[c]);
							const float gsc = __ldg(a.d_scaling + p * 3 + c);
							d_sr[c] = gsc * __ldg(gs + 3 + c) * s * (1.f - s);
							pp[6 + c] = gsc * s;                      // d get_scaling[:, 3 + c]
						}
						if (a.d_xyz) {
							dxyz[c] = __ldg(a.d_xyz + p * 3 + c);
							pp[c] = dxyz[c];                          // d anchor
							pp[3 + c] = dxyz[c] * __ldg(off + c);     // d get_scaling[:, c]
						}
					}
					if (a.d_rot) {
						// r = v / max(|v|, eps):  dv = (g - r (r . g)) / max(|v|, eps)   (|v| > eps branch; below it dv = g / eps)
						const float n2 = sr[3] * sr[3] + sr[4] * sr[4] + sr[5] * sr[5] + sr[6] * sr[6];
						const float nrm = sqrtf(n2);
						float gr[4], dot = 0.f;
#pragma unroll
						for (int c = 0; c < 4; c++) {
							gr[c] = __ldg(a.d_rot + p * 4 + c);
							dot += gr[c] * sr[3 + c];
						}
						if (nrm > 1e-12f) {
#pragma unroll
							for (int c = 0; c < 4; c++) d_sr[3 + c] = (gr[c] - sr[3 + c] * dot / n2) / nrm;
						} else {
#pragma unroll
							for (int c = 0; c < 4; c++) d_sr[3 + c] = gr[c] / 1e-12f;
						}
					}
					if (a.d_xyz) {
#pragma unroll
						for (int c = 0; c < 3; c++) a.g_offset[((size_t)aid * k + j) * 3 + c] = dxyz[c] * __ldg(gs + c);
					}
				}
				d_op = g_op * (1.f - nop * nop);
			}
			__syncwarp();
			if (on) {
				// gradients of the pre-activations replace the pre-activations (all reads of this pair's slots are done)
				OUT4[j * kNA + aa] = d_op;
				OUT4[(k + j) * kNA + aa] = d_unc;
#pragma unroll
				for (int c = 0; c < 7; c++) OUT4[(2 * k + 7 * j + c) * kNA + aa] = d_sr[c];
#pragma unroll
				for (int c = 0; c < 3; c++) OUT4[(9 * k + 3 * j + c) * kNA + aa] = d_col[c];
#pragma unroll
				for (int c = 0; c < 9; c++) PP[(aa * k + j) * 10 + c] = pp[c];
			}
		}
		__syncwarp();

		// ---- layer 2 backward: lane = h.  dh[m][a] = sum_{o in m} W2[o][h] dout[o][a], gated by relu -----------------
#pragma unroll
		for (int m = 0; m < 4; m++) {
			float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
			const int ob = out_base(m, k), oe = ob + out_count(m, k);
			for (int o = ob; o < oe; o++) {
				const float w = sm[pl.w2 + lane * pl.os + o];
				const float4 dv = *reinterpret_cast<const float4 *>(OUT4 + o * kNA);
				d0 = fmaf(w, dv.x, d0);
				d1 = fmaf(w, dv.y, d1);
				d2 = fmaf(w, dv.z, d2);
				d3 = fmaf(w, dv.w, d3);
			}
			const float4 hv = *reinterpret_cast<const float4 *>(H4 + (m * kHStride + lane) * kNA);
			*reinterpret_cast<float4 *>(DH4 + (m * kHStride + lane) * kNA) =
			    make_float4(hv.x > 0.f ? d0 : 0.f, hv.y > 0.f ? d1 : 0.f, hv.z > 0.f ? d2 : 0.f, hv.w > 0.f ? d3 : 0.f);
		}
		__syncwarp();

		// ---- layer 1 backward: lane = input i (and input 32 + (lane & 3) in a second set) ----------------------------
		{
			float dx[kNA] = {0.f, 0.f, 0.f, 0.f}, de[kNA] = {0.f, 0.f, 0.f, 0.f};
			const int ie = kHid + (lane & 3);
			for (int mh = 0; mh < 4 * kHid; mh++) {
				const int m = mh >> 5, h = mh & 31;
				const float w = sm[pl.w1 + (m * kIn + lane) * kW1Stride + h];
				const float we = sm[pl.w1 + (m * kIn + ie) * kW1Stride + h];
				const float4 dv = *reinterpret_cast<const float4 *>(DH4 + (m * kHStride + h) * kNA);
				dx[0] = fmaf(w, dv.x, dx[0]); dx[1] = fmaf(w, dv.y, dx[1]); dx[2] = fmaf(w, dv.z, dx[2]); dx[3] = fmaf(w, dv.w, dx[3]);
				de[0] = fmaf(we, dv.x, de[0]); de[1] = fmaf(we, dv.y, de[1]); de[2] = fmaf(we, dv.z, de[2]); de[3] = fmaf(we, dv.w, de[3]);
			}
#pragma unroll
			for (int aa = 0; aa < kNA; aa++) {
				// view / distance path (gaussian_renderer/__init__.py:31-35): v = anchor - cam, dist = |v|, u = v / dist
				const float gux = __shfl_sync(0xffffffffu, de[aa], 0), guy = __shfl_sync(0xffffffffu, de[aa], 1);
				const float guz = __shfl_sync(0xffffffffu, de[aa], 2), gdist = __shfl_sync(0xffffffffu, de[aa], 3);
				if (!io.valid[aa] || g >= groups) continue;
				a.g_feat[(size_t)io.id[aa] * kHid + lane] = dx[aa];
				if (lane < 9) {
					// sum this anchor's per-offset partials: [0..2] d anchor (from xyz), [3..8] d get_scaling
					float s = 0.f;
					for (int j = 0; j < k; j++) s += PP[(aa * k + j) * 10 + lane];
					if (lane < 3) {
						const float u = lane == 0 ? io.ux[aa] : lane == 1 ? io.uy[aa] : io.uz[aa];
						const float gu = lane == 0 ? gux : lane == 1 ? guy : guz;
						const float udot = io.ux[aa] * gux + io.uy[aa] * guy + io.uz[aa] * guz;
						s += (gu - u * udot) / io.dist[aa] + u * gdist;
						a.g_anchor[(size_t)io.id[aa] * 3 + lane] = s;
					} else {
						a.g_scaling[(size_t)io.id[aa] * 6 + (lane - 3)] = s;
					}
				}
			}
		}
		__syncthreads(); // every warp's X4 / H4 / DH4 / DOUT4 of this CTA iteration are complete

		// ---- weight gradients: each thread owns a slice, loops over the CTA's 8 groups ------------------------------
		for (int w = 0; w < kWarps; w++) {
			const float *wb = sm + pl.warp0 + w * pl.per_warp;
			const float4 dh = *reinterpret_cast<const float4 *>(wb + pl.dh4 + (m1 * kHStride + lane) * kNA);
			if (half == 0) accb1 += (dh.x + dh.y) + (dh.z + dh.w);
#pragma unroll
			for (int q = 0; q < 18; q++) {
				const float4 x = *reinterpret_cast<const float4 *>(wb + pl.x4 + (18 * half + q) * kNA);
				acc1[q] = fmaf(dh.x, x.x, fmaf(dh.y, x.y, fmaf(dh.z, x.z, fmaf(dh.w, x.w, acc1[q]))));
			}
			const float4 hv = *reinterpret_cast<const float4 *>(wb + pl.h4 + (m2 * kHStride + lane) * kNA);
#pragma unroll
			for (int q = 0; q < kAcc2; q++) {
				if (q < o_cnt) {
					const float4 dv = *reinterpret_cast<const float4 *>(wb + pl.out4 + (o_first + q) * kNA);
					acc2[q] = fmaf(hv.x, dv.x, fmaf(hv.y, dv.y, fmaf(hv.z, dv.z, fmaf(hv.w, dv.w, acc2[q]))));
				}
			}
			if (lane < o_cnt) {
				const float4 dv = *reinterpret_cast<const float4 *>(wb + pl.out4 + (o_first + lane) * kNA);
				accb2 += (dv.x + dv.y) + (dv.z + dv.w);
			}
		}
		__syncthreads(); // slices consumed before the next iteration overwrites the activations
	}

	// ---- flush: one atomic per owned entry (torch layouts: w1[h][i], b1[h], w2[o][h], b2[o]) ---------------------
#pragma unroll
	for (int q = 0; q < 18; q++) atomicAdd(a.g_w1[m1] + lane * kIn + 18 * half + q, acc1[q]);
	if (half == 0) atomicAdd(a.g_b1[m1] + lane, accb1);
#pragma unroll
	for (int q = 0; q < kAcc2; q++)
		if (q < o_cnt) atomicAdd(a.g_w2[m2] + (o_first + q - out_base(m2, k)) * kHid + lane, acc2[q]);
	if (lane < o_cnt) atomicAdd(a.g_b2[m2] + (o_first + lane - out_base(m2, k)), accb2);
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

// ---- small helpers for the visible-anchor list ------------------------------------------------------------------------
__global__ void mask_to_flags_kernel(int A, const uint8_t *__restrict__ mask, uint32_t *__restrict__ flags)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < A) flags[i] = mask[i] ? 1u : 0u;
}
__global__ void compact_visible_kernel(int A, const uint32_t *__restrict__ flags, const uint32_t *__restrict__ incl, uint32_t *__restrict__ ids)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < A && flags[i]) ids[incl[i] - 1] = (uint32_t)i;
}

// persistent CTAs per SM of the two forward kernels (72 KB of shared memory and 64 registers each: three fit)
#ifndef GSR_DEC_FWD_CTAS
#define GSR_DEC_FWD_CTAS 3
#endif

int sm_count()
{
	static int sms = 0;
	if (!sms) {
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (sms <= 0) sms = 148;
	}
	return sms;
}
// persistent grid: enough CTAs to cover the groups, at most `per_sm` per SM
int grid_for(int anchors, int per_sm)
{
	const int groups = (anchors + kNA - 1) / kNA;
	const int want = (groups + kWarps - 1) / kWarps;
	return max(1, min(want, per_sm * sm_count()));
}

template <typename K>
cudaError_t set_smem(K kernel, size_t bytes)
{
	return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

} // namespace

DecodeLayout decode_layout(int A)
{
	DecodeLayout L{};
	size_t off = 0;
	const size_t n = (size_t)(A > 0 ? A : 1) * 4;
	L.vis_flag = off;   off += align_up(n);
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
	uint32_t *flags = (uint32_t *)(scratch + L.vis_flag), *incl = (uint32_t *)(scratch + L.vis_incl), *ids = (uint32_t *)(scratch + L.vis_ids);
	uint32_t *count = (uint32_t *)(scratch + L.count), *bits = (uint32_t *)(scratch + L.maskbits), *gincl = (uint32_t *)(scratch + L.gauss_incl);
	if (visible_mask) {
		mask_to_flags_kernel<<<(A + 255) / 256, 256, 0, stream>>>(A, visible_mask, flags);
		if ((e = inclusive_sum_gather(flags, nullptr, incl, A, scratch + L.scan_tmp, stream)) != cudaSuccess) return e;
		compact_visible_kernel<<<(A + 255) / 256, 256, 0, stream>>>(A, flags, incl, ids);
		count_launch(2);
	}
	if ((e = cudaMemsetAsync(count, 0, (size_t)A * 4, stream)) != cudaSuccess) return e;
	DecodeArgs a{};
	a.k = k; a.n_vis = A; a.n_vis_dev = visible_mask ? incl + (A - 1) : nullptr;
	a.vis_ids = visible_mask ? ids : nullptr;
	a.anchor = anchor; a.feat = feat; a.campos = campos; a.wt = wt;
	a.neural_opacity = neural_opacity; a.mask = mask; a.count = count; a.maskbits = bits;
	const size_t smem = (size_t)smem_plan(k, false).total * 4;
	if ((e = set_smem(decode_opacity_kernel, smem)) != cudaSuccess) return e;
	decode_opacity_kernel<<<grid_for(A, GSR_DEC_FWD_CTAS), 256, smem, stream>>>(a);
	count_launch(2);
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
	const size_t smem = (size_t)smem_plan(a.k, false).total * 4;
	if ((e = set_smem(decode_outputs_kernel, smem)) != cudaSuccess) return e;
	decode_outputs_kernel<<<grid_for(a.n_vis, GSR_DEC_FWD_CTAS), 256, smem, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}

cudaError_t decode_backward(const DecodeBwdArgs &a, cudaStream_t stream)
{
	cudaError_t e;
	const size_t smem = (size_t)smem_plan(a.f.k, true).total * 4;
	if ((e = set_smem(decode_backward_kernel, smem)) != cudaSuccess) return e;
	// persistent; 2 CTAs/SM (102 KB of shared memory each at k = 10, 128 registers)
	decode_backward_kernel<<<grid_for(a.f.n_vis, 2), 256, smem, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}

size_t statis_scratch_bytes(int A, int k)
{
	const size_t a = (size_t)(A > 0 ? A : 1), n = a * (size_t)(k > 0 ? k : 1);
	return 2 * align_up(a * 4) + 2 * align_up(n * 4) + scan_scratch_bytes((int64_t)n);
}

cudaError_t training_statis(int A, int k, int64_t n_vis, int64_t P, const uint8_t *anchor_visible, const uint8_t *offset_selected, const uint8_t *update_filter,
                            const float *neural_opacity, const float *viewspace_grad, float *opacity_accum, float *anchor_demon,
                            float *offset_gradient_accum, float *offset_denom, char *scratch, cudaStream_t stream)
{
	const size_t a = (size_t)A, n = (size_t)n_vis * k;
	uint32_t *vflag = (uint32_t *)scratch, *vincl = (uint32_t *)(scratch + align_up(a * 4));
	uint32_t *sflag = (uint32_t *)(scratch + 2 * align_up(a * 4)), *sincl = (uint32_t *)(scratch + 2 * align_up(a * 4) + align_up((size_t)A * k * 4));
	char *tmp = scratch + 2 * align_up(a * 4) + 2 * align_up((size_t)A * k * 4);
	cudaError_t e;
	mask_to_flags_kernel<<<(A + 255) / 256, 256, 0, stream>>>(A, anchor_visible, vflag);
	if ((e = inclusive_sum_gather(vflag, nullptr, vincl, A, tmp, stream)) != cudaSuccess) return e;
	count_launch();
	if (n > 0) {
		mask_to_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((int)n, offset_selected, sflag);
		if ((e = inclusive_sum_gather(sflag, nullptr, sincl, (int64_t)n, tmp, stream)) != cudaSuccess) return e;
		count_launch();
	}
	training_statis_kernel<<<(A + 255) / 256, 256, 0, stream>>>(A, k, n_vis, P, anchor_visible, vincl, offset_selected, sincl, update_filter, neural_opacity,
	                                                          viewspace_grad, opacity_accum, anchor_demon, offset_gradient_accum, offset_denom);
	count_launch();
	return cudaGetLastError();
}

} // namespace gsr
