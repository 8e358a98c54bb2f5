// gsr_loss.cu — fused image loss L1 + SSIM (forward and backward), SURVEY.md section 8f rank 3.
//
// Replaces, for the photometric term of train.py:535-545 of W-Ted/GScream,
//     l1_loss / l1_loss_masked            utils/loss_utils.py:27-31
//     ssim / _ssim, ssim_masked / _ssim_masked   utils/loss_utils.py:131-207
// i.e. five depthwise 11x11 Gaussian convolutions (of x, y, x^2, y^2, xy), ~15 elementwise kernels and two reductions in the
// forward and about twice that in the backward, by one kernel each way:
//   forward : per 16x16 output tile of one image plane, stage the 26x26 halo of x (rendered) and y (target) in shared memory,
//             run the Gaussian window separably (11 taps horizontally on 26 rows, then 11 taps vertically), evaluate the SSIM
//             map and |x - y|, weight both by the optional mask, block-reduce and add to two fp64 accumulators; keep the three
//             partial-derivative planes dm/dmu1, dm/d(conv x^2), dm/d(conv xy) (times the mask) for the backward.
//   backward: convolve the three planes with the same (symmetric, zero-padded => self-adjoint) window and combine:
//             dL/dx = g_ssim/N * (conv(p1) + 2 x conv(p2) + y conv(p3)) + g_l1/N * sign(x - y) * mask.
// The window is separable by construction (create_window: outer product of the normalised 1-D Gaussian, loss_utils.py:116-121);
// the 11 taps come from the host, computed exactly as the reference computes them (fp32).
// HBM-bound by design: forward reads 2 planes and writes 3, backward reads 5 (+mask) and writes 1; ~150 FMA per pixel.
#include "gsr_internal.cuh"
#include "gsr_loss.cuh"

namespace gsr {

namespace {

constexpr int kT = 16;          // output tile edge
constexpr int kR = 5;           // window radius (window_size 11)
constexpr int kTH = kT + 2 * kR; // 26: tile + halo

struct Taps { float g[2 * kR + 1]; };

__device__ __forceinline__ double block_sum_256(float v, float *s_red /*[8]*/)
{
#pragma unroll
	for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
	const int tid = threadIdx.y * kT + threadIdx.x;
	if ((tid & 31) == 0) s_red[tid >> 5] = v;
	__syncthreads();
	double t = 0.0;
	if (tid == 0)
		for (int w = 0; w < 8; w++) t += (double)s_red[w];
	__syncthreads();
	return t; // valid on thread 0
}

__global__ void __launch_bounds__(256) l1_ssim_forward_kernel(int H, int W, Taps taps, const float *__restrict__ x, const float *__restrict__ y,
                                                              const float *__restrict__ mask, int mask_planes, double *__restrict__ sums,
                                                              float *__restrict__ p1, float *__restrict__ p2, float *__restrict__ p3)
{
	__shared__ float sx[kTH][kTH + 1], sy[kTH][kTH + 1];
	__shared__ float hc[5][kTH][kT];
	__shared__ float s_red[8];
	const int plane = blockIdx.z;
	const size_t pbase = (size_t)plane * H * W;
	const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
	const int tid = threadIdx.y * kT + threadIdx.x;
	for (int i = tid; i < kTH * kTH; i += 256) {
		const int r = i / kTH, c = i - r * kTH;
		const int gy = y0 + r - kR, gx = x0 + c - kR;
		const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W; // zero padding (padding = window_size // 2)
		sx[r][c] = in ? __ldg(x + pbase + (size_t)gy * W + gx) : 0.f;
		sy[r][c] = in ? __ldg(y + pbase + (size_t)gy * W + gx) : 0.f;
	}
	__syncthreads();
	for (int i = tid; i < kTH * kT; i += 256) {
		const int r = i / kT, c = i - r * kT;
		float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
		for (int k = 0; k <= 2 * kR; k++) {
			const float g = taps.g[k], u = sx[r][c + k], v = sy[r][c + k];
			a = fmaf(g, u, a);
			b = fmaf(g, v, b);
			aa = fmaf(g, u * u, aa);
			bb = fmaf(g, v * v, bb);
			ab = fmaf(g, u * v, ab);
		}
		hc[0][r][c] = a; hc[1][r][c] = b; hc[2][r][c] = aa; hc[3][r][c] = bb; hc[4][r][c] = ab;
	}
	__syncthreads();
	const int tx = threadIdx.x, ty = threadIdx.y;
	const int gx = x0 + tx, gy = y0 + ty;
	float m_w = 0.f, l1_w = 0.f;
	if (gx < W && gy < H) {
		float mu1 = 0.f, mu2 = 0.f, cxx = 0.f, cyy = 0.f, cxy = 0.f;
#pragma unroll
		for (int k = 0; k <= 2 * kR; k++) {
			const float g = taps.g[k];
			mu1 = fmaf(g, hc[0][ty + k][tx], mu1);
			mu2 = fmaf(g, hc[1][ty + k][tx], mu2);
			cxx = fmaf(g, hc[2][ty + k][tx], cxx);
			cyy = fmaf(g, hc[3][ty + k][tx], cyy);
			cxy = fmaf(g, hc[4][ty + k][tx], cxy);
		}
		// utils/loss_utils.py:145-157
		const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
		const float s1 = cxx - mu1_sq, s2 = cyy - mu2_sq, s12 = cxy - mu12;
		const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
		const float A = 2.f * mu12 + C1, B = 2.f * s12 + C2, D = mu1_sq + mu2_sq + C1, E = s1 + s2 + C2;
		const float inv_DE = 1.f / (D * E);
		const float m = A * B * inv_DE;
		const size_t pix = (size_t)gy * W + gx;
		const float wm = mask ? __ldg(mask + (size_t)(mask_planes == 1 ? 0 : plane) * H * W + pix) : 1.f;
		m_w = m * wm;
		l1_w = fabsf(sx[ty + kR][tx + kR] - sy[ty + kR][tx + kR]) * wm;
		if (p1) {
			// total derivatives of m w.r.t. the three convolution outputs that depend on x: mu1, conv(x^2), conv(xy)
			const float dm_dmu1 = 2.f * (mu2 * (B - A) * inv_DE + mu1 * m * (1.f / E - 1.f / D));
			p1[pbase + pix] = dm_dmu1 * wm;
			p2[pbase + pix] = -m / E * wm;
			p3[pbase + pix] = 2.f * A * inv_DE * wm;
		}
	}
	const double ts = block_sum_256(m_w, s_red);
	if (tid == 0) atomicAdd(sums + 0, ts);
	const double tl = block_sum_256(l1_w, s_red);
	if (tid == 0) atomicAdd(sums + 1, tl);
}

__global__ void __launch_bounds__(256) l1_ssim_backward_kernel(int H, int W, Taps taps, const float *__restrict__ x, const float *__restrict__ y,
                                                               const float *__restrict__ mask, int mask_planes, const float *__restrict__ p1,
                                                               const float *__restrict__ p2, const float *__restrict__ p3,
                                                               const float *__restrict__ upstream /*[2]: dL/d(ssim mean), dL/d(l1 mean)*/,
                                                               float inv_n, float *__restrict__ dx)
{
	__shared__ float sp[3][kTH][kTH + 1];
	__shared__ float hc[3][kTH][kT];
	const int plane = blockIdx.z;
	const size_t pbase = (size_t)plane * H * W;
	const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
	const int tid = threadIdx.y * kT + threadIdx.x;
	for (int i = tid; i < kTH * kTH; i += 256) {
		const int r = i / kTH, c = i - r * kTH;
		const int gy = y0 + r - kR, gx = x0 + c - kR;
		const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
		const size_t o = pbase + (size_t)gy * W + gx;
		sp[0][r][c] = in ? __ldg(p1 + o) : 0.f;
		sp[1][r][c] = in ? __ldg(p2 + o) : 0.f;
		sp[2][r][c] = in ? __ldg(p3 + o) : 0.f;
	}
	__syncthreads();
	for (int i = tid; i < kTH * kT; i += 256) {
		const int r = i / kT, c = i - r * kT;
		float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
		for (int k = 0; k <= 2 * kR; k++) {
			const float g = taps.g[k];
			a = fmaf(g, sp[0][r][c + k], a);
			b = fmaf(g, sp[1][r][c + k], b);
			d = fmaf(g, sp[2][r][c + k], d);
		}
		hc[0][r][c] = a; hc[1][r][c] = b; hc[2][r][c] = d;
	}
	__syncthreads();
	const int tx = threadIdx.x, ty = threadIdx.y;
	const int gx = x0 + tx, gy = y0 + ty;
	if (gx >= W || gy >= H) return;
	float c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
	for (int k = 0; k <= 2 * kR; k++) {
		const float g = taps.g[k];
		c1 = fmaf(g, hc[0][ty + k][tx], c1);
		c2 = fmaf(g, hc[1][ty + k][tx], c2);
		c3 = fmaf(g, hc[2][ty + k][tx], c3);
	}
	const size_t pix = (size_t)gy * W + gx;
	const float xv = __ldg(x + pbase + pix), yv = __ldg(y + pbase + pix);
	const float wm = mask ? __ldg(mask + (size_t)(mask_planes == 1 ? 0 : plane) * H * W + pix) : 1.f;
	const float gs = __ldg(upstream) * inv_n, gl = __ldg(upstream + 1) * inv_n;
	const float diff = xv - yv;
	const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f); // torch.abs backward: sign(0) = 0
	dx[pbase + pix] = gs * (c1 + 2.f * xv * c2 + yv * c3) + gl * sgn * wm;
}


// ---- scale / shift aligned depth L1 (train.py:548-569 with utils/loss_utils.py:80-102, 27-31) -----------------------------
//   (s, t) = argmin sum_i m_i (s d_i + t - y_i)^2      closed form from five masked sums (compute_scale_and_shift)
//   L      = mean_i( |abs(s) d_i + t - y_i| * lm_i )   l1_loss (lm = 1) or l1_loss_masked
// The reference builds this from ~25 eager kernels and back-propagates through the closed form; here: one reduction for the
// five sums, one for the loss (which also leaves the two sums the backward needs), one elementwise backward kernel.
// The five sums are accumulated in fp64 and the 2x2 system is solved in fp64: det = a00 a11 - a01^2 cancels badly at image
// sizes (both products ~1e13 for a 567x1008 depth map), where the reference's fp32 evaluation (utils/loss_utils.py:96-100) keeps
// only ~3 digits of it; the result here is the exactly-rounded one, the reference's own fp32 deviation is what the tests allow for.
__device__ __forceinline__ void solve_scale_shift(const double *sums, double &x0, double &x1, double &a00, double &a01, double &a11, double &det)
{
	a00 = sums[0]; a01 = sums[1]; a11 = sums[2];
	const double b0 = sums[3], b1 = sums[4];
	det = a00 * a11 - a01 * a01;
	x0 = det != 0.0 ? (a11 * b0 - a01 * b1) / det : 0.0;
	x1 = det != 0.0 ? (-a01 * b0 + a00 * b1) / det : 0.0;
}
__device__ __forceinline__ double block_sum_1d(float v, float *s_red)
{
#pragma unroll
	for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
	if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
	__syncthreads();
	double t = 0.0;
	if (threadIdx.x == 0)
		for (int w = 0; w < 8; w++) t += (double)s_red[w];
	__syncthreads();
	return t;
}
constexpr int kDepthItems = 8; // pixels per thread

__global__ void __launch_bounds__(256) depth_fit_sums_kernel(int n, const float *__restrict__ d, const float *__restrict__ y,
                                                             const float *__restrict__ m, double *__restrict__ sums)
{
	__shared__ float s_red[8];
	const size_t base = (size_t)blockIdx.y * n;
	float a00 = 0.f, a01 = 0.f, a11 = 0.f, b0 = 0.f, b1 = 0.f;
	for (int k = 0; k < kDepthItems; k++) {
		const int i = (blockIdx.x * kDepthItems + k) * 256 + threadIdx.x;
		if (i < n) {
			const float dv = d[base + i], yv = y[base + i], mv = m ? m[base + i] : 1.f;
			a00 += mv * dv * dv; a01 += mv * dv; a11 += mv; b0 += mv * dv * yv; b1 += mv * yv;
		}
	}
	double t;
	t = block_sum_1d(a00, s_red); if (threadIdx.x == 0) atomicAdd(sums + blockIdx.y * 5 + 0, t);
	t = block_sum_1d(a01, s_red); if (threadIdx.x == 0) atomicAdd(sums + blockIdx.y * 5 + 1, t);
	t = block_sum_1d(a11, s_red); if (threadIdx.x == 0) atomicAdd(sums + blockIdx.y * 5 + 2, t);
	t = block_sum_1d(b0, s_red);  if (threadIdx.x == 0) atomicAdd(sums + blockIdx.y * 5 + 3, t);
	t = block_sum_1d(b1, s_red);  if (threadIdx.x == 0) atomicAdd(sums + blockIdx.y * 5 + 4, t);
}

__global__ void __launch_bounds__(256) depth_l1_kernel(int n, const float *__restrict__ d, const float *__restrict__ y,
                                                       const float *__restrict__ lm, const double *__restrict__ sums,
                                                       double *__restrict__ aux /*[B][2]: S1, S0*/, double *__restrict__ loss_sum)
{
	__shared__ float s_red[8];
	const size_t base = (size_t)blockIdx.y * n;
	double x0d, x1d, a00, a01, a11, det;
	solve_scale_shift(sums + blockIdx.y * 5, x0d, x1d, a00, a01, a11, det);
	const float sc = fabsf((float)x0d), x1 = (float)x1d;   // train.py:552 scale = torch.abs(scale)
	float l = 0.f, s1 = 0.f, s0 = 0.f;
	for (int k = 0; k < kDepthItems; k++) {
		const int i = (blockIdx.x * kDepthItems + k) * 256 + threadIdx.x;
		if (i < n) {
			const float dv = d[base + i], w = lm ? lm[base + i] : 1.f;
			const float r = sc * dv + x1 - y[base + i];
			const float sg = (r > 0.f ? 1.f : (r < 0.f ? -1.f : 0.f)) * w;
			l += fabsf(r) * w; s1 += sg * dv; s0 += sg;
		}
	}
	double t;
	t = block_sum_1d(l, s_red);  if (threadIdx.x == 0) atomicAdd(loss_sum, t);
	t = block_sum_1d(s1, s_red); if (threadIdx.x == 0) atomicAdd(aux + blockIdx.y * 2 + 0, t);
	t = block_sum_1d(s0, s_red); if (threadIdx.x == 0) atomicAdd(aux + blockIdx.y * 2 + 1, t);
}

__global__ void __launch_bounds__(256) depth_l1_backward_kernel(int n, float inv_total, const float *__restrict__ d, const float *__restrict__ y,
                                                                const float *__restrict__ m, const float *__restrict__ lm,
                                                                const double *__restrict__ sums, const double *__restrict__ aux,
                                                                const float *__restrict__ upstream, float *__restrict__ dd)
{
	const size_t base = (size_t)blockIdx.y * n;
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	double x0d, x1d, a00, a01, a11, det;
	solve_scale_shift(sums + blockIdx.y * 5, x0d, x1d, a00, a01, a11, det);
	const float x0 = (float)x0d, x1 = (float)x1d, sc = fabsf(x0);
	// dL/dx0 = sign(x0) * S1 / N, dL/dx1 = S0 / N; lambda = A^-1 (dL/dx0, dL/dx1), in fp64 for the same cancellation reason
	const double g0 = (x0d > 0.0 ? 1.0 : (x0d < 0.0 ? -1.0 : 0.0)) * aux[blockIdx.y * 2 + 0] * (double)inv_total;
	const double g1 = aux[blockIdx.y * 2 + 1] * (double)inv_total;
	const float l0 = det != 0.0 ? (float)((a11 * g0 - a01 * g1) / det) : 0.f;
	const float l1 = det != 0.0 ? (float)((-a01 * g0 + a00 * g1) / det) : 0.f;
	const float dv = d[base + i], yv = y[base + i];
	const float mv = m ? m[base + i] : 1.f, w = lm ? lm[base + i] : 1.f;
	const float r = sc * dv + x1 - yv;
	const float sg = (r > 0.f ? 1.f : (r < 0.f ? -1.f : 0.f)) * w;
	dd[base + i] = __ldg(upstream) * (sg * sc * inv_total + mv * (l0 * (yv - 2.f * dv * x0 - x1) - l1 * x0));
}


// ---- multi-scale gradient-matching loss on the aligned depth (train.py:232-251 `gradient_loss`, called at train.py:556-560 and
// :571-574 on `aligned_depth[:, ::step, ::step]` for step = 1, 2, 4, 8) --------------------------------------------------------
//   e_p      = m_p (A_p - y_p),  A_p = abs(s) d_p + t  (or A = the prediction itself when no fit is given)
//   loss_s   = mean_b( [ sum_{p,q horizontal / vertical neighbours on the stride-2^s grid} |e_q - e_p| m_p m_q ] / M_{b,s} ),
//   M_{b,s}  = sum of the mask on that grid (not divided when M = 0: reduction_image_based, train.py:221-230)
// The reference slices, multiplies and reduces ~25 eager kernels per scale (x4 scales, twice that in the backward).  Here one
// kernel visits every full-resolution pixel once and serves all scales whose grid contains it; per (image, scale) it leaves
// {sum, M, S1 = d(sum)/d(abs s), S0 = d(sum)/dt} in fp64 so the backward can carry the gradient through the closed-form fit.
constexpr int kGradMaxScales = 4;

struct GradPix { float e, m, d; };

__device__ __forceinline__ GradPix grad_pix(const float *__restrict__ d, const float *__restrict__ y, const float *__restrict__ m,
                                            size_t i, float sc, float x1)
{
	GradPix p;
	p.d = d[i];
	p.m = m ? m[i] : 1.f;
	p.e = p.m * ((sc * p.d + x1) - y[i]);
	return p;
}
__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); } // torch.abs backward: sign(0) = 0

__global__ void __launch_bounds__(256) depth_grad_forward_kernel(int H, int W, int n_scales, const float *__restrict__ d,
                                                                 const float *__restrict__ y, const float *__restrict__ m,
                                                                 const double *__restrict__ fit_sums, double *__restrict__ gstate /*[B][4][4]*/)
{
	__shared__ float s_red[8];
	const int n = H * W;
	const size_t base = (size_t)blockIdx.y * n;
	float sc = 1.f, x1 = 0.f;
	if (fit_sums) {
		double x0d, x1d, a00, a01, a11, det;
		solve_scale_shift(fit_sums + blockIdx.y * 5, x0d, x1d, a00, a01, a11, det);
		sc = fabsf((float)x0d); x1 = (float)x1d;
	}
	float acc[kGradMaxScales][4];
#pragma unroll
	for (int s = 0; s < kGradMaxScales; s++) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.f;
	for (int k = 0; k < kDepthItems; k++) {
		const int i = (blockIdx.x * kDepthItems + k) * 256 + threadIdx.x;
		if (i >= n) continue;
		const int r = i / W, c = i - r * W;
		const GradPix p = grad_pix(d, y, m, base + i, sc, x1);
#pragma unroll
		for (int s = 0; s < kGradMaxScales; s++) {
			const int step = 1 << s;
			if (s >= n_scales || ((r | c) & (step - 1))) continue;
			acc[s][1] += p.m;
			if (c + step < W) {
				const GradPix q = grad_pix(d, y, m, base + i + step, sc, x1);
				const float dl = q.e - p.e, w = p.m * q.m, sg = sgnf(dl) * w;
				acc[s][0] += fabsf(dl) * w; acc[s][2] += sg * (q.m * q.d - p.m * p.d); acc[s][3] += sg * (q.m - p.m);
			}
			if (r + step < H) {
				const GradPix q = grad_pix(d, y, m, base + i + (size_t)step * W, sc, x1);
				const float dl = q.e - p.e, w = p.m * q.m, sg = sgnf(dl) * w;
				acc[s][0] += fabsf(dl) * w; acc[s][2] += sg * (q.m * q.d - p.m * p.d); acc[s][3] += sg * (q.m - p.m);
			}
		}
	}
#pragma unroll
	for (int s = 0; s < kGradMaxScales; s++) {
		if (s >= n_scales) break;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const double t = block_sum_1d(acc[s][j], s_red);
			if (threadIdx.x == 0 && t != 0.0) atomicAdd(gstate + ((size_t)blockIdx.y * kGradMaxScales + s) * 4 + j, t);
		}
	}
}

// total[0] = sum_s mean_b( sum_{b,s} / M_{b,s} )
__global__ void depth_grad_finalize_kernel(int B, int n_scales, const double *__restrict__ gstate, double *__restrict__ total)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	double t = 0.0;
	for (int b = 0; b < B; b++)
		for (int s = 0; s < n_scales; s++) {
			const double *g = gstate + ((size_t)b * kGradMaxScales + s) * 4;
			t += g[1] != 0.0 ? g[0] / g[1] : g[0];
		}
	total[0] = t / (double)B;
}

__global__ void __launch_bounds__(256) depth_grad_backward_kernel(int H, int W, int n_scales, float inv_B, const float *__restrict__ d,
                                                                  const float *__restrict__ y, const float *__restrict__ m,
                                                                  const float *__restrict__ fm, const double *__restrict__ fit_sums,
                                                                  const double *__restrict__ gstate, const float *__restrict__ upstream,
                                                                  float *__restrict__ dd, int accumulate)
{
	const int n = H * W;
	const size_t base = (size_t)blockIdx.y * n;
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	const float up = __ldg(upstream) * inv_B;
	float sc = 1.f, x0 = 1.f, x1 = 0.f, l0 = 0.f, l1 = 0.f;
	float coef[kGradMaxScales];
	double g0 = 0.0, g1 = 0.0;
#pragma unroll
	for (int s = 0; s < kGradMaxScales; s++) {
		coef[s] = 0.f;
		if (s >= n_scales) continue;
		const double *g = gstate + ((size_t)blockIdx.y * kGradMaxScales + s) * 4;
		const double cf = (double)up / (g[1] != 0.0 ? g[1] : 1.0);
		coef[s] = (float)cf;
		g0 += cf * g[2]; g1 += cf * g[3];
	}
	if (fit_sums) {
		double x0d, x1d, a00, a01, a11, det;
		solve_scale_shift(fit_sums + blockIdx.y * 5, x0d, x1d, a00, a01, a11, det);
		x0 = (float)x0d; x1 = (float)x1d; sc = fabsf(x0);
		g0 *= (x0d > 0.0 ? 1.0 : (x0d < 0.0 ? -1.0 : 0.0));   // d abs(s) / ds
		l0 = det != 0.0 ? (float)((a11 * g0 - a01 * g1) / det) : 0.f;
		l1 = det != 0.0 ? (float)((-a01 * g0 + a00 * g1) / det) : 0.f;
	}
	const int r = i / W, c = i - r * W;
	const GradPix p = grad_pix(d, y, m, base + i, sc, x1);
	float g = 0.f;
#pragma unroll
	for (int s = 0; s < kGradMaxScales; s++) {
		const int step = 1 << s;
		if (s >= n_scales || ((r | c) & (step - 1))) continue;
		float t = 0.f;
		if (c + step < W) { const GradPix q = grad_pix(d, y, m, base + i + step, sc, x1); t -= sgnf(q.e - p.e) * q.m; }
		if (c - step >= 0) { const GradPix q = grad_pix(d, y, m, base + i - step, sc, x1); t += sgnf(p.e - q.e) * q.m; }
		if (r + step < H) { const GradPix q = grad_pix(d, y, m, base + i + (size_t)step * W, sc, x1); t -= sgnf(q.e - p.e) * q.m; }
		if (r - step >= 0) { const GradPix q = grad_pix(d, y, m, base + i - (size_t)step * W, sc, x1); t += sgnf(p.e - q.e) * q.m; }
		g += coef[s] * t;
	}
	g *= p.m * p.m;                                        // dL/dA_p: pair weight m_p m_q times de_p/dA_p = m_p
	float out = g * sc;
	if (fit_sums) {
		const float mv = fm ? fm[base + i] : 1.f;
		out += mv * (l0 * (y[base + i] - 2.f * p.d * x0 - x1) - l1 * x0);
	}
	dd[base + i] = accumulate ? dd[base + i] + out : out;
}

} // namespace

cudaError_t launch_l1_ssim_forward(int planes, int H, int W, const float *taps11, const float *x, const float *y, const float *mask,
                                   int mask_planes, double *sums, float *p1, float *p2, float *p3, cudaStream_t stream)
{
	Taps t;
	for (int k = 0; k < 11; k++) t.g[k] = taps11[k];
	cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double), stream);
	if (e != cudaSuccess) return e;
	dim3 grid((W + kT - 1) / kT, (H + kT - 1) / kT, planes), block(kT, kT);
	l1_ssim_forward_kernel<<<grid, block, 0, stream>>>(H, W, t, x, y, mask, mask_planes, sums, p1, p2, p3);
	count_launch(2);
	return cudaGetLastError();
}

cudaError_t launch_l1_ssim_backward(int planes, int H, int W, const float *taps11, const float *x, const float *y, const float *mask,
                                    int mask_planes, const float *p1, const float *p2, const float *p3, const float *upstream, float *dx,
                                    cudaStream_t stream)
{
	Taps t;
	for (int k = 0; k < 11; k++) t.g[k] = taps11[k];
	dim3 grid((W + kT - 1) / kT, (H + kT - 1) / kT, planes), block(kT, kT);
	l1_ssim_backward_kernel<<<grid, block, 0, stream>>>(H, W, t, x, y, mask, mask_planes, p1, p2, p3, upstream,
	                                                   1.0f / ((float)planes * (float)H * (float)W), dx);
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_depth_align_l1_forward(int B, int n, const float *d, const float *y, const float *fit_mask, const float *loss_mask,
                                          double *sums, double *aux, double *loss_sum, cudaStream_t stream)
{
	cudaError_t e;
	if ((e = cudaMemsetAsync(sums, 0, (size_t)B * 5 * sizeof(double), stream)) != cudaSuccess) return e;
	if ((e = cudaMemsetAsync(aux, 0, (size_t)B * 2 * sizeof(double), stream)) != cudaSuccess) return e;
	if ((e = cudaMemsetAsync(loss_sum, 0, sizeof(double), stream)) != cudaSuccess) return e;
	dim3 grid((n + 256 * kDepthItems - 1) / (256 * kDepthItems), B);
	depth_fit_sums_kernel<<<grid, 256, 0, stream>>>(n, d, y, fit_mask, sums);
	depth_l1_kernel<<<grid, 256, 0, stream>>>(n, d, y, loss_mask, sums, aux, loss_sum);
	count_launch(5);
	return cudaGetLastError();
}

cudaError_t launch_depth_align_l1_backward(int B, int n, const float *d, const float *y, const float *fit_mask, const float *loss_mask,
                                           const double *sums, const double *aux, const float *upstream, float *dd, cudaStream_t stream)
{
	dim3 grid((n + 255) / 256, B);
	depth_l1_backward_kernel<<<grid, 256, 0, stream>>>(n, 1.0f / ((float)B * (float)n), d, y, fit_mask, loss_mask, sums, aux, upstream, dd);
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_depth_grad_forward(int B, int H, int W, int n_scales, const float *d, const float *y, const float *mask,
                                      const double *fit_sums, double *gstate, cudaStream_t stream)
{
	cudaError_t e;
	if ((e = cudaMemsetAsync(gstate, 0, (1 + (size_t)B * kGradMaxScales * 4) * sizeof(double), stream)) != cudaSuccess) return e;
	const int n = H * W;
	dim3 grid((n + 256 * kDepthItems - 1) / (256 * kDepthItems), B);
	depth_grad_forward_kernel<<<grid, 256, 0, stream>>>(H, W, n_scales, d, y, mask, fit_sums, gstate + 1);
	depth_grad_finalize_kernel<<<1, 32, 0, stream>>>(B, n_scales, gstate + 1, gstate);
	count_launch(3);
	return cudaGetLastError();
}

cudaError_t launch_depth_grad_backward(int B, int H, int W, int n_scales, const float *d, const float *y, const float *mask,
                                       const float *fit_mask, const double *fit_sums, const double *gstate, const float *upstream,
                                       float *dd, int accumulate, cudaStream_t stream)
{
	dim3 grid((H * W + 255) / 256, B);
	depth_grad_backward_kernel<<<grid, 256, 0, stream>>>(H, W, n_scales, 1.0f / (float)B, d, y, mask, fit_mask, fit_sums, gstate + 1,
	                                                     upstream, dd, accumulate);
	count_launch();
	return cudaGetLastError();
}

} // namespace gsr
