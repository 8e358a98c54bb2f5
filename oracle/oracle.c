/*
 * oracle/oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the differentiable Gaussian
 * rasterizer that W-Ted/GScream ships as submodules/diff-gaussian-rasterization
 * (CUDA only; the reference has no CPU rasterizer and no tests).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may load
 * it; gscream_b200/ never does.
 *
 * Parity status: PINNED against the reference's own CUDA build (oracle/_ref,
 * compiled from /root/reference by oracle/build_ref.py) through the golden
 * vectors in tests/golden/ generated on a B200 by tests/golden/make_golden.py.
 * The reference itself ships no golden vectors.
 *
 * Compiled twice by oracle/Makefile: REAL=float (liboracle_f32.so, IEEE fp32,
 * no FMA contraction) and REAL=double (liboracle_f64.so).  All citations are
 * relative to /root/reference/submodules/diff-gaussian-rasterization/ ; CR/ is
 * its cuda_rasterizer/ directory.
 *
 * Conventions restated from the reference:
 *  - matrices (viewmatrix, projmatrix) arrive as 16 floats, used column-major
 *    exactly as CR/auxiliary.h:58-97 indexes them;
 *  - GLM mat3 is column-major, m[col][row]; mat3*mat3 evaluates
 *    R[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]
 *    (third_party/glm/glm/detail/type_mat3x3.inl:509-517);
 *  - tiles are 16x16 (CR/config.h:16-17).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif
typedef REAL real;

#define BLOCK_X 16
#define BLOCK_Y 16

#if defined(ORACLE_F64)
#define R_SQRT sqrt
#define R_EXP exp
#define R_CEIL ceil
#else
#define R_SQRT sqrtf
#define R_EXP expf
#define R_CEIL ceilf
#endif

static real rmin(real a, real b) { return a < b ? a : b; }
static real rmax(real a, real b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

typedef struct { real m[3][3]; } mat3; /* m[col][row], as GLM */

/* GLM operator*(mat3, mat3): type_mat3x3.inl:486-519 */
static mat3 mat3_mul(const mat3 *a, const mat3 *b)
{
	mat3 r;
	for (int c = 0; c < 3; c++)
		for (int w = 0; w < 3; w++)
			r.m[c][w] = a->m[0][w] * b->m[c][0] + a->m[1][w] * b->m[c][1] + a->m[2][w] * b->m[c][2];
	return r;
}

static mat3 mat3_transpose(const mat3 *a)
{
	mat3 r;
	for (int c = 0; c < 3; c++)
		for (int w = 0; w < 3; w++)
			r.m[c][w] = a->m[w][c];
	return r;
}

/* CR/auxiliary.h:41-44 — evaluated in double because of the 1.0 / 0.5 literals, then
 * rounded to float on return. */
static real ndc2pix(real v, int S)
{
	return (real)((((double)v + 1.0) * S - 1.0) * 0.5);
}

/* CR/auxiliary.h:46-56 */
static void get_rect(real px, real py, int max_radius, int gx, int gy, int *rmin_x, int *rmin_y,
                     int *rmax_x, int *rmax_y)
{
	*rmin_x = imin(gx, imax(0, (int)((px - max_radius) / BLOCK_X)));
	*rmin_y = imin(gy, imax(0, (int)((py - max_radius) / BLOCK_Y)));
	*rmax_x = imin(gx, imax(0, (int)((px + max_radius + BLOCK_X - 1) / BLOCK_X)));
	*rmax_y = imin(gy, imax(0, (int)((py + max_radius + BLOCK_Y - 1) / BLOCK_Y)));
}

/* CR/auxiliary.h:58-80 */
static void transform_point_4x3(const real p[3], const float *m, real out[3])
{
	out[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
	out[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
	out[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static void transform_point_4x4(const real p[3], const float *m, real out[4])
{
	out[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
	out[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
	out[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
	out[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* CR/forward.cu:120-154.  The quaternion is used as given (normalisation is commented out,
 * forward.cu:129). */
static void build_rotation(const real q[4], mat3 *R)
{
	real r = q[0], x = q[1], y = q[2], z = q[3];
	R->m[0][0] = (real)1 - (real)2 * (y * y + z * z);
	R->m[0][1] = (real)2 * (x * y - r * z);
	R->m[0][2] = (real)2 * (x * z + r * y);
	R->m[1][0] = (real)2 * (x * y + r * z);
	R->m[1][1] = (real)1 - (real)2 * (x * x + z * z);
	R->m[1][2] = (real)2 * (y * z - r * x);
	R->m[2][0] = (real)2 * (x * z - r * y);
	R->m[2][1] = (real)2 * (y * z + r * x);
	R->m[2][2] = (real)1 - (real)2 * (x * x + y * y);
}

static void compute_cov3d(const real scale[3], real mod, const real q[4], real cov3D[6])
{
	mat3 S, R;
	memset(&S, 0, sizeof(S));
	S.m[0][0] = mod * scale[0];
	S.m[1][1] = mod * scale[1];
	S.m[2][2] = mod * scale[2];
	build_rotation(q, &R);
	mat3 M = mat3_mul(&S, &R);
	mat3 Mt = mat3_transpose(&M);
	mat3 Sigma = mat3_mul(&Mt, &M);
	cov3D[0] = Sigma.m[0][0];
	cov3D[1] = Sigma.m[0][1];
	cov3D[2] = Sigma.m[0][2];
	cov3D[3] = Sigma.m[1][1];
	cov3D[4] = Sigma.m[1][2];
	cov3D[5] = Sigma.m[2][2];
}

/* Shared front half of CR/forward.cu:76-115 and CR/backward.cu:155-199: clamped view-space
 * mean, Jacobian-like J, W, T = W*J, Vrk, cov2D (low-pass +0.3 added by the caller). */
typedef struct {
	real t[3];
	real txtz, tytz, limx, limy;
	mat3 J, W, T, Vrk, cov;
} cov2d_ctx;

static void cov2d_common(const real mean[3], real focal_x, real focal_y, real tan_fovx, real tan_fovy,
                         const real cov3D[6], const float *view, cov2d_ctx *c)
{
	transform_point_4x3(mean, view, c->t);
	c->limx = (real)1.3f * tan_fovx;
	c->limy = (real)1.3f * tan_fovy;
	c->txtz = c->t[0] / c->t[2];
	c->tytz = c->t[1] / c->t[2];
	c->t[0] = rmin(c->limx, rmax(-c->limx, c->txtz)) * c->t[2];
	c->t[1] = rmin(c->limy, rmax(-c->limy, c->tytz)) * c->t[2];
	real tz = c->t[2];
	memset(&c->J, 0, sizeof(mat3));
	c->J.m[0][0] = focal_x / tz;
	c->J.m[0][2] = -(focal_x * c->t[0]) / (tz * tz);
	c->J.m[1][1] = focal_y / tz;
	c->J.m[1][2] = -(focal_y * c->t[1]) / (tz * tz);
	c->W.m[0][0] = view[0]; c->W.m[0][1] = view[4]; c->W.m[0][2] = view[8];
	c->W.m[1][0] = view[1]; c->W.m[1][1] = view[5]; c->W.m[1][2] = view[9];
	c->W.m[2][0] = view[2]; c->W.m[2][1] = view[6]; c->W.m[2][2] = view[10];
	c->T = mat3_mul(&c->W, &c->J);
	c->Vrk.m[0][0] = cov3D[0]; c->Vrk.m[0][1] = cov3D[1]; c->Vrk.m[0][2] = cov3D[2];
	c->Vrk.m[1][0] = cov3D[1]; c->Vrk.m[1][1] = cov3D[3]; c->Vrk.m[1][2] = cov3D[4];
	c->Vrk.m[2][0] = cov3D[2]; c->Vrk.m[2][1] = cov3D[4]; c->Vrk.m[2][2] = cov3D[5];
	mat3 Tt = mat3_transpose(&c->T);
	mat3 Vt = mat3_transpose(&c->Vrk);
	mat3 A = mat3_mul(&Tt, &Vt);
	c->cov = mat3_mul(&A, &c->T);
}

/* ------------------------------------------------------------------------------------------
 * K1 preprocessCUDA forward (CR/forward.cu:157-267) and its two anchor-filter variants
 * filter_preprocessCUDA (:271-346) / position2D_preprocessCUDA (:352-433), plus in_frustum
 * (CR/auxiliary.h:139-164).  mode: 0 = full preprocess, 1 = visible_filter (radii only),
 * 2 = position2D_filter (radii + pixel x,y).
 * Inputs are float32 arrays as they cross the reference's C++ boundary; arithmetic is `real`.
 * Outputs not relevant to a mode may be NULL.
 * ------------------------------------------------------------------------------------------ */
void orc_preprocess(int mode, int P, const float *means3D, const float *scales, float scale_modifier,
                    const float *rotations, const float *opacities, const float *uncertainties,
                    const float *cov3D_precomp, const float *view, const float *proj, int W, int H,
                    float tan_fovx, float tan_fovy,
                    int *radii, real *xy, real *depths, real *cov3Ds, real *conic_opacity,
                    uint32_t *tiles_touched, real *unc_out, real *pos2d_x, real *pos2d_y)
{
	/* CR/rasterizer_impl.cu:226-227 */
	const real focal_y = (real)((float)H / (2.0f * tan_fovy));
	const real focal_x = (real)((float)W / (2.0f * tan_fovx));
	const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;

	for (int idx = 0; idx < P; idx++) {
		radii[idx] = 0;
		if (tiles_touched) tiles_touched[idx] = 0;
		if (pos2d_x) pos2d_x[idx] = 0;
		if (pos2d_y) pos2d_y[idx] = 0;

		real p[3] = { means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2] };
		real p_hom[4], p_view[3];
		transform_point_4x4(p, proj, p_hom);
		real p_w = (real)1 / (p_hom[3] + (real)0.0000001f);
		real p_proj[3] = { p_hom[0] * p_w, p_hom[1] * p_w, p_hom[2] * p_w };
		transform_point_4x3(p, view, p_view);
		if (p_view[2] <= (real)0.2f) /* CR/auxiliary.h:154 — the only frustum test */
			continue;

		real cov3D[6];
		if (cov3D_precomp) {
			for (int k = 0; k < 6; k++) cov3D[k] = cov3D_precomp[6 * idx + k];
		} else {
			real s[3] = { scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2] };
			real q[4] = { rotations[4 * idx], rotations[4 * idx + 1], rotations[4 * idx + 2], rotations[4 * idx + 3] };
			compute_cov3d(s, (real)scale_modifier, q, cov3D);
			if (cov3Ds) for (int k = 0; k < 6; k++) cov3Ds[6 * idx + k] = cov3D[k];
		}

		cov2d_ctx c;
		cov2d_common(p, focal_x, focal_y, (real)tan_fovx, (real)tan_fovy, cov3D, view, &c);
		real cx = c.cov.m[0][0] + (real)0.3f; /* CR/forward.cu:112-113 */
		real cy = c.cov.m[0][1];
		real cz = c.cov.m[1][1] + (real)0.3f;

		real det = cx * cz - cy * cy;
		if (det == (real)0) continue;
		real det_inv = (real)1 / det;
		real conic[3] = { cz * det_inv, -cy * det_inv, cx * det_inv };

		real mid = (real)0.5f * (cx + cz);
		real lambda1 = mid + R_SQRT(rmax((real)0.1f, mid * mid - det));
		real lambda2 = mid - R_SQRT(rmax((real)0.1f, mid * mid - det));
		real my_radius = R_CEIL((real)3 * R_SQRT(rmax(lambda1, lambda2)));
		real pix_x = ndc2pix(p_proj[0], W), pix_y = ndc2pix(p_proj[1], H);
		int x0, y0, x1, y1;
		/* getRect receives the radius as int (CR/auxiliary.h:46) and the float2 pixel centre */
		get_rect(pix_x, pix_y, (int)my_radius, gx, gy, &x0, &y0, &x1, &y1);
		if ((x1 - x0) * (y1 - y0) == 0) continue;

		radii[idx] = (int)my_radius;
		if (mode == 2) { pos2d_x[idx] = pix_x; pos2d_y[idx] = pix_y; }
		if (mode != 0) continue;

		depths[idx] = p_view[2];
		xy[2 * idx] = pix_x;
		xy[2 * idx + 1] = pix_y;
		conic_opacity[4 * idx + 0] = conic[0];
		conic_opacity[4 * idx + 1] = conic[1];
		conic_opacity[4 * idx + 2] = conic[2];
		conic_opacity[4 * idx + 3] = opacities[idx];
		tiles_touched[idx] = (uint32_t)((y1 - y0) * (x1 - x0));
		unc_out[idx] = uncertainties[idx];
	}
}

/* K12 checkFrustum (CR/rasterizer_impl.cu:54-66) */
void orc_mark_visible(int P, const float *means3D, const float *view, const float *proj, uint8_t *present)
{
	(void)proj;
	for (int idx = 0; idx < P; idx++) {
		real p[3] = { means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2] };
		real p_view[3];
		transform_point_4x3(p, view, p_view);
		present[idx] = !(p_view[2] <= (real)0.2f);
	}
}

/* ------------------------------------------------------------------------------------------
 * Binning: K2 inclusive scan, K3 duplicateWithKeys (CR/rasterizer_impl.cu:70-111), K4 stable
 * sort of (tile<<32 | depth-bits) keys (:309-314; bit range from getHigherMsb :35-50 covers
 * every set bit, so a full-key stable sort is equivalent), K5 identifyTileRanges (:116-138).
 * depth_bits: the fp32 bit pattern of each Gaussian's view-space depth.
 * Returns R.  keys_unsorted/vals_unsorted/keys_sorted/point_list need capacity >= R
 * (call once with them NULL to obtain R).
 * ------------------------------------------------------------------------------------------ */
uint32_t orc_get_higher_msb(uint32_t n)
{
	uint32_t msb = sizeof(n) * 4, step = msb;
	while (step > 1) {
		step /= 2;
		if (n >> msb) msb += step; else msb -= step;
	}
	if (n >> msb) msb++;
	return msb;
}

typedef struct { uint64_t key; uint32_t val; } kv_t;

static void stable_sort_kv(kv_t *a, kv_t *tmp, size_t n)
{
	/* bottom-up merge sort: stable, like cub::DeviceRadixSort */
	for (size_t w = 1; w < n; w *= 2) {
		for (size_t lo = 0; lo < n; lo += 2 * w) {
			size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
			size_t i = lo, j = mid, k = lo;
			while (i < mid && j < hi) tmp[k++] = (a[j].key < a[i].key) ? a[j++] : a[i++];
			while (i < mid) tmp[k++] = a[i++];
			while (j < hi) tmp[k++] = a[j++];
		}
		memcpy(a, tmp, n * sizeof(kv_t));
	}
}

int64_t orc_binning(int P, const float *xy, const uint32_t *depth_bits, const int *radii, int W, int H,
                    uint32_t *point_offsets, uint64_t *keys_unsorted, uint32_t *vals_unsorted,
                    uint64_t *keys_sorted, uint32_t *point_list, uint32_t *ranges /* [tiles][2] */)
{
	const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
	uint32_t run = 0;
	for (int idx = 0; idx < P; idx++) {
		if (radii[idx] > 0) {
			int x0, y0, x1, y1;
			get_rect((real)xy[2 * idx], (real)xy[2 * idx + 1], radii[idx], gx, gy, &x0, &y0, &x1, &y1);
			run += (uint32_t)((y1 - y0) * (x1 - x0));
		}
		if (point_offsets) point_offsets[idx] = run; /* inclusive scan of tiles_touched */
	}
	const uint32_t R = run;
	if (!keys_unsorted) return R;

	uint32_t off = 0;
	for (int idx = 0; idx < P; idx++) {
		if (!(radii[idx] > 0)) continue;
		int x0, y0, x1, y1;
		get_rect((real)xy[2 * idx], (real)xy[2 * idx + 1], radii[idx], gx, gy, &x0, &y0, &x1, &y1);
		for (int y = y0; y < y1; y++)
			for (int x = x0; x < x1; x++) {
				uint64_t key = (uint64_t)(y * gx + x);
				key <<= 32;
				key |= depth_bits[idx];
				keys_unsorted[off] = key;
				vals_unsorted[off] = (uint32_t)idx;
				off++;
			}
	}
	kv_t *a = (kv_t *)malloc(sizeof(kv_t) * (R ? R : 1)), *tmp = (kv_t *)malloc(sizeof(kv_t) * (R ? R : 1));
	for (uint32_t i = 0; i < R; i++) { a[i].key = keys_unsorted[i]; a[i].val = vals_unsorted[i]; }
	stable_sort_kv(a, tmp, R);
	for (uint32_t i = 0; i < R; i++) { keys_sorted[i] = a[i].key; point_list[i] = a[i].val; }
	free(a); free(tmp);

	memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy); /* CR/rasterizer_impl.cu:316 */
	for (uint32_t i = 0; i < R; i++) {
		uint32_t cur = (uint32_t)(keys_sorted[i] >> 32);
		if (i == 0) ranges[2 * cur] = 0;
		else {
			uint32_t prev = (uint32_t)(keys_sorted[i - 1] >> 32);
			if (cur != prev) { ranges[2 * prev + 1] = i; ranges[2 * cur] = i; }
		}
		if (i == R - 1) ranges[2 * cur + 1] = R;
	}
	return R;
}

/* ------------------------------------------------------------------------------------------
 * K6 renderCUDA forward (CR/forward.cu:441-568).  One pixel at a time; the block-cooperative
 * batching of the CUDA kernel (rounds of 256, __syncthreads_count early exit) changes no
 * result, so a pixel simply walks its tile's list until it is `done`.
 * features: [P][C]; out_color: [C][H][W]; out_depth/out_unc/final_T: [H][W]; n_contrib u32.
 * ------------------------------------------------------------------------------------------ */
void orc_render_forward(int C, int W, int H, const uint32_t *ranges, const uint32_t *point_list,
                        const real *xy, const real *features, const real *depths, const real *unc,
                        const real *conic_opacity, const real *bg,
                        real *final_T, uint32_t *n_contrib, real *out_color, real *out_depth, real *out_unc)
{
	const int gx = (W + BLOCK_X - 1) / BLOCK_X;
	real *acc = (real *)malloc(sizeof(real) * (size_t)(C > 0 ? C : 1));
	for (int py = 0; py < H; py++)
		for (int px = 0; px < W; px++) {
			const int tile = (py / BLOCK_Y) * gx + (px / BLOCK_X);
			const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
			const size_t pix_id = (size_t)W * py + px;
			const real pixf_x = (real)px, pixf_y = (real)py; /* integer pixel coords, forward.cu:466 */
			real T = 1;
			uint32_t contributor = 0, last_contributor = 0;
			real D = 0, UNC = 0;
			for (int ch = 0; ch < C; ch++) acc[ch] = 0;
			for (uint32_t i = r0; i < r1; i++) {
				contributor++;
				const uint32_t id = point_list[i];
				real dx = xy[2 * id] - pixf_x, dy = xy[2 * id + 1] - pixf_y;
				const real *co = conic_opacity + 4 * (size_t)id;
				real power = (real)-0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
				if (power > (real)0) continue;
				real alpha = rmin((real)0.99f, co[3] * R_EXP(power));
				if (alpha < (real)(1.0f / 255.0f)) continue;
				real test_T = T * ((real)1 - alpha);
				if (test_T < (real)0.0001f) break; /* done: this Gaussian is NOT blended */
				for (int ch = 0; ch < C; ch++) acc[ch] += features[(size_t)id * C + ch] * alpha * T;
				D += depths[id] * alpha * T;
				UNC += unc[id] * alpha * T;
				T = test_T;
				last_contributor = contributor;
			}
			final_T[pix_id] = T;
			n_contrib[pix_id] = last_contributor;
			for (int ch = 0; ch < C; ch++) out_color[(size_t)ch * H * W + pix_id] = acc[ch] + T * bg[ch];
			out_depth[pix_id] = D;  /* no background term, forward.cu:564-565 */
			out_unc[pix_id] = UNC;
		}
	free(acc);
}

/* ------------------------------------------------------------------------------------------
 * K7 renderCUDA backward (CR/backward.cu:409-604).  The CUDA kernel scatters with float
 * atomics in arbitrary order; here pixels are visited row-major and sums are sequential.
 * Gradient outputs must arrive zero-filled (DGR/rasterize_points.cu:160-170).
 * dL_dmean2D: [P][3] (z never written); dL_dconic: [P][4] (.z never written).
 * ------------------------------------------------------------------------------------------ */
void orc_render_backward(int C, int W, int H, const uint32_t *ranges, const uint32_t *point_list,
                         const real *bg, const real *xy, const real *conic_opacity, const real *colors,
                         const real *depths, const real *unc, const real *final_Ts, const uint32_t *n_contrib,
                         const real *dL_dpixels, const real *dL_dpixel_depths, const real *dL_dpixel_uncs,
                         real *dL_dmean2D, real *dL_dconic, real *dL_dopacity, real *dL_dcolors,
                         real *dL_ddepths, real *dL_duncs)
{
	const int gx = (W + BLOCK_X - 1) / BLOCK_X;
	const size_t cs = (size_t)(C > 0 ? C : 1);
	real *accum_rec = (real *)malloc(sizeof(real) * cs), *last_color = (real *)malloc(sizeof(real) * cs);
	real *dL_dpixel = (real *)malloc(sizeof(real) * cs);
	const real ddelx_dx = (real)(0.5 * W), ddely_dy = (real)(0.5 * H);
	for (int py = 0; py < H; py++)
		for (int px = 0; px < W; px++) {
			const int tile = (py / BLOCK_Y) * gx + (px / BLOCK_X);
			const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
			const size_t pix_id = (size_t)W * py + px;
			const real pixf_x = (real)px, pixf_y = (real)py;
			const real T_final = final_Ts[pix_id];
			real T = T_final;
			uint32_t contributor = r1 - r0;
			const uint32_t last_contributor = n_contrib[pix_id];
			for (int ch = 0; ch < C; ch++) {
				accum_rec[ch] = 0; last_color[ch] = 0;
				dL_dpixel[ch] = dL_dpixels[(size_t)ch * H * W + pix_id];
			}
			const real dL_dpixel_depth = dL_dpixel_depths[pix_id], dL_dunc = dL_dpixel_uncs[pix_id];
			real accum_depth_rec = 0, accum_unc_rec = 0;
			real last_alpha = 0, last_depth = 0, last_unc = 0;
			for (uint32_t k = 0; k < r1 - r0; k++) {
				const uint32_t id = point_list[r1 - k - 1]; /* back to front, backward.cu:500 */
				contributor--;
				if (contributor >= last_contributor) continue;
				real dx = xy[2 * id] - pixf_x, dy = xy[2 * id + 1] - pixf_y;
				const real *co = conic_opacity + 4 * (size_t)id;
				real power = (real)-0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
				if (power > (real)0) continue;
				const real G = R_EXP(power);
				const real alpha = rmin((real)0.99f, co[3] * G);
				if (alpha < (real)(1.0f / 255.0f)) continue;

				T = T / ((real)1 - alpha);
				const real dchannel_dcolor = alpha * T;
				real dL_dalpha = 0;
				for (int ch = 0; ch < C; ch++) {
					const real c = colors[(size_t)id * C + ch];
					accum_rec[ch] = last_alpha * last_color[ch] + ((real)1 - last_alpha) * accum_rec[ch];
					last_color[ch] = c;
					dL_dalpha += (c - accum_rec[ch]) * dL_dpixel[ch];
					dL_dcolors[(size_t)id * C + ch] += dchannel_dcolor * dL_dpixel[ch];
				}
				const real c_d = depths[id];
				accum_depth_rec = last_alpha * last_depth + ((real)1 - last_alpha) * accum_depth_rec;
				last_depth = c_d;
				dL_dalpha += (c_d - accum_depth_rec) * dL_dpixel_depth;
				dL_ddepths[id] += dchannel_dcolor * dL_dpixel_depth;

				const real c_unc = unc[id];
				accum_unc_rec = last_alpha * last_unc + ((real)1 - last_alpha) * accum_unc_rec;
				last_unc = c_unc;
				dL_dalpha += (c_unc - accum_unc_rec) * dL_dunc;
				dL_duncs[id] += dchannel_dcolor * dL_dunc;

				dL_dalpha *= T;
				last_alpha = alpha;

				/* background term: colour channels only (backward.cu:578-581) */
				real bg_dot_dpixel = 0;
				for (int ch = 0; ch < C; ch++) bg_dot_dpixel += bg[ch] * dL_dpixel[ch];
				dL_dalpha += (-T_final / ((real)1 - alpha)) * bg_dot_dpixel;

				/* min(0.99, .) is straight-through (backward.cu:585) */
				const real dL_dG = co[3] * dL_dalpha;
				const real gdx = G * dx, gdy = G * dy;
				const real dG_ddelx = -gdx * co[0] - gdy * co[1];
				const real dG_ddely = -gdy * co[2] - gdx * co[1];
				dL_dmean2D[3 * (size_t)id + 0] += dL_dG * dG_ddelx * ddelx_dx;
				dL_dmean2D[3 * (size_t)id + 1] += dL_dG * dG_ddely * ddely_dy;
				dL_dconic[4 * (size_t)id + 0] += (real)-0.5f * gdx * dx * dL_dG;
				dL_dconic[4 * (size_t)id + 1] += (real)-0.5f * gdx * dy * dL_dG;
				dL_dconic[4 * (size_t)id + 3] += (real)-0.5f * gdy * dy * dL_dG;
				dL_dopacity[id] += G * dL_dalpha;
			}
		}
	free(accum_rec); free(last_color); free(dL_dpixel);
}

/* ------------------------------------------------------------------------------------------
 * K8 computeCov2DCUDA (CR/backward.cu:144-274) + K9 preprocessCUDA backward (:346-406) with
 * computeCov3D backward (:278-341).  SH branch is not restated (GScream always passes
 * shs=None, gaussian_renderer/__init__.py:152).
 * dL_dmeans/dL_dcov3D/dL_dscale/dL_drot must arrive zero-filled; rows of culled Gaussians stay 0.
 * cov3Ds: the forward's cached cov3D (or cov3D_precomp).  scales may be NULL (precomputed cov).
 * ------------------------------------------------------------------------------------------ */
void orc_preprocess_backward(int P, const float *means3D, const int *radii, const float *scales,
                             const float *rotations, float scale_modifier, const real *cov3Ds,
                             const float *view, const float *proj, int W, int H, float tan_fovx, float tan_fovy,
                             const real *dL_dmean2D, const real *dL_dconic, const real *dL_ddepth,
                             real *dL_dmeans, real *dL_dcov3D, real *dL_dscale, real *dL_drot)
{
	const real h_y = (real)((float)H / (2.0f * tan_fovy));
	const real h_x = (real)((float)W / (2.0f * tan_fovx));
	for (int idx = 0; idx < P; idx++) {
		if (!(radii[idx] > 0)) continue;
		real mean[3] = { means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2] };
		const real *cov3D = cov3Ds + 6 * (size_t)idx;
		/* ---- K8 ---- */
		real dLc[3] = { dL_dconic[4 * (size_t)idx], dL_dconic[4 * (size_t)idx + 1], dL_dconic[4 * (size_t)idx + 3] };
		cov2d_ctx c;
		cov2d_common(mean, h_x, h_y, (real)tan_fovx, (real)tan_fovy, cov3D, view, &c);
		const real x_grad_mul = (c.txtz < -c.limx || c.txtz > c.limx) ? 0 : 1;
		const real y_grad_mul = (c.tytz < -c.limy || c.tytz > c.limy) ? 0 : 1;
		const mat3 *T = &c.T, *Vrk = &c.Vrk, *Wm = &c.W;
		real a = c.cov.m[0][0] + (real)0.3f;
		real b = c.cov.m[0][1];
		real cc = c.cov.m[1][1] + (real)0.3f;
		real denom = a * cc - b * b;
		real dL_da = 0, dL_db = 0, dL_dc = 0;
		real denom2inv = (real)1 / ((denom * denom) + (real)0.0000001f);
		real *dcov = dL_dcov3D + 6 * (size_t)idx;
		if (denom2inv != 0) {
			dL_da = denom2inv * (-cc * cc * dLc[0] + 2 * b * cc * dLc[1] + (denom - a * cc) * dLc[2]);
			dL_dc = denom2inv * (-a * a * dLc[2] + 2 * a * b * dLc[1] + (denom - a * cc) * dLc[0]);
			dL_db = denom2inv * 2 * (b * cc * dLc[0] - (denom + 2 * b * b) * dLc[1] + a * b * dLc[2]);
			dcov[0] = (T->m[0][0] * T->m[0][0] * dL_da + T->m[0][0] * T->m[1][0] * dL_db + T->m[1][0] * T->m[1][0] * dL_dc);
			dcov[3] = (T->m[0][1] * T->m[0][1] * dL_da + T->m[0][1] * T->m[1][1] * dL_db + T->m[1][1] * T->m[1][1] * dL_dc);
			dcov[5] = (T->m[0][2] * T->m[0][2] * dL_da + T->m[0][2] * T->m[1][2] * dL_db + T->m[1][2] * T->m[1][2] * dL_dc);
			dcov[1] = 2 * T->m[0][0] * T->m[0][1] * dL_da + (T->m[0][0] * T->m[1][1] + T->m[0][1] * T->m[1][0]) * dL_db + 2 * T->m[1][0] * T->m[1][1] * dL_dc;
			dcov[2] = 2 * T->m[0][0] * T->m[0][2] * dL_da + (T->m[0][0] * T->m[1][2] + T->m[0][2] * T->m[1][0]) * dL_db + 2 * T->m[1][0] * T->m[1][2] * dL_dc;
			dcov[4] = 2 * T->m[0][2] * T->m[0][1] * dL_da + (T->m[0][1] * T->m[1][2] + T->m[0][2] * T->m[1][1]) * dL_db + 2 * T->m[1][1] * T->m[1][2] * dL_dc;
		} else {
			for (int i = 0; i < 6; i++) dcov[i] = 0;
		}
		real dL_dT00 = 2 * (T->m[0][0] * Vrk->m[0][0] + T->m[0][1] * Vrk->m[0][1] + T->m[0][2] * Vrk->m[0][2]) * dL_da +
		               (T->m[1][0] * Vrk->m[0][0] + T->m[1][1] * Vrk->m[0][1] + T->m[1][2] * Vrk->m[0][2]) * dL_db;
		real dL_dT01 = 2 * (T->m[0][0] * Vrk->m[1][0] + T->m[0][1] * Vrk->m[1][1] + T->m[0][2] * Vrk->m[1][2]) * dL_da +
		               (T->m[1][0] * Vrk->m[1][0] + T->m[1][1] * Vrk->m[1][1] + T->m[1][2] * Vrk->m[1][2]) * dL_db;
		real dL_dT02 = 2 * (T->m[0][0] * Vrk->m[2][0] + T->m[0][1] * Vrk->m[2][1] + T->m[0][2] * Vrk->m[2][2]) * dL_da +
		               (T->m[1][0] * Vrk->m[2][0] + T->m[1][1] * Vrk->m[2][1] + T->m[1][2] * Vrk->m[2][2]) * dL_db;
		real dL_dT10 = 2 * (T->m[1][0] * Vrk->m[0][0] + T->m[1][1] * Vrk->m[0][1] + T->m[1][2] * Vrk->m[0][2]) * dL_dc +
		               (T->m[0][0] * Vrk->m[0][0] + T->m[0][1] * Vrk->m[0][1] + T->m[0][2] * Vrk->m[0][2]) * dL_db;
		real dL_dT11 = 2 * (T->m[1][0] * Vrk->m[1][0] + T->m[1][1] * Vrk->m[1][1] + T->m[1][2] * Vrk->m[1][2]) * dL_dc +
		               (T->m[0][0] * Vrk->m[1][0] + T->m[0][1] * Vrk->m[1][1] + T->m[0][2] * Vrk->m[1][2]) * dL_db;
		real dL_dT12 = 2 * (T->m[1][0] * Vrk->m[2][0] + T->m[1][1] * Vrk->m[2][1] + T->m[1][2] * Vrk->m[2][2]) * dL_dc +
		               (T->m[0][0] * Vrk->m[2][0] + T->m[0][1] * Vrk->m[2][1] + T->m[0][2] * Vrk->m[2][2]) * dL_db;
		real dL_dJ00 = Wm->m[0][0] * dL_dT00 + Wm->m[0][1] * dL_dT01 + Wm->m[0][2] * dL_dT02;
		real dL_dJ02 = Wm->m[2][0] * dL_dT00 + Wm->m[2][1] * dL_dT01 + Wm->m[2][2] * dL_dT02;
		real dL_dJ11 = Wm->m[1][0] * dL_dT10 + Wm->m[1][1] * dL_dT11 + Wm->m[1][2] * dL_dT12;
		real dL_dJ12 = Wm->m[2][0] * dL_dT10 + Wm->m[2][1] * dL_dT11 + Wm->m[2][2] * dL_dT12;
		real tz = (real)1 / c.t[2];
		real tz2 = tz * tz;
		real tz3 = tz2 * tz;
		real dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
		real dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
		real dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * c.t[0]) * tz3 * dL_dJ02 + (2 * h_y * c.t[1]) * tz3 * dL_dJ12;
		/* transformVec4x3Transpose, CR/auxiliary.h:90-97; assignment, backward.cu:273 */
		real dm[3] = {
			view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz,
			view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz,
			view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz,
		};

		/* ---- K9 ---- */
		real m_hom[4];
		transform_point_4x4(mean, proj, m_hom);
		real m_w = (real)1 / (m_hom[3] + (real)0.0000001f);
		real mul1 = (proj[0] * mean[0] + proj[4] * mean[1] + proj[8] * mean[2] + proj[12]) * m_w * m_w;
		real mul2 = (proj[1] * mean[0] + proj[5] * mean[1] + proj[9] * mean[2] + proj[13]) * m_w * m_w;
		real g2x = dL_dmean2D[3 * (size_t)idx], g2y = dL_dmean2D[3 * (size_t)idx + 1];
		dm[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
		dm[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
		dm[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
		/* depth path, backward.cu:391-396 */
		dm[0] += view[2] * dL_ddepth[idx];
		dm[1] += view[6] * dL_ddepth[idx];
		dm[2] += view[10] * dL_ddepth[idx];
		dL_dmeans[3 * (size_t)idx + 0] = dm[0];
		dL_dmeans[3 * (size_t)idx + 1] = dm[1];
		dL_dmeans[3 * (size_t)idx + 2] = dm[2];

		if (scales) { /* computeCov3D backward, backward.cu:278-341 */
			real q[4] = { rotations[4 * idx], rotations[4 * idx + 1], rotations[4 * idx + 2], rotations[4 * idx + 3] };
			real r = q[0], x = q[1], y = q[2], z = q[3];
			mat3 R, S, dL_dSigma;
			build_rotation(q, &R);
			memset(&S, 0, sizeof(S));
			real s[3] = { (real)scale_modifier * scales[3 * idx], (real)scale_modifier * scales[3 * idx + 1], (real)scale_modifier * scales[3 * idx + 2] };
			S.m[0][0] = s[0]; S.m[1][1] = s[1]; S.m[2][2] = s[2];
			mat3 M = mat3_mul(&S, &R);
			dL_dSigma.m[0][0] = dcov[0];              dL_dSigma.m[0][1] = (real)0.5f * dcov[1]; dL_dSigma.m[0][2] = (real)0.5f * dcov[2];
			dL_dSigma.m[1][0] = (real)0.5f * dcov[1]; dL_dSigma.m[1][1] = dcov[3];              dL_dSigma.m[1][2] = (real)0.5f * dcov[4];
			dL_dSigma.m[2][0] = (real)0.5f * dcov[2]; dL_dSigma.m[2][1] = (real)0.5f * dcov[4]; dL_dSigma.m[2][2] = dcov[5];
			mat3 M2;
			for (int cI = 0; cI < 3; cI++) for (int w = 0; w < 3; w++) M2.m[cI][w] = (real)2 * M.m[cI][w]; /* 2.0f * M */
			mat3 dL_dM = mat3_mul(&M2, &dL_dSigma);
			mat3 Rt = mat3_transpose(&R);
			mat3 dL_dMt = mat3_transpose(&dL_dM);
			for (int k = 0; k < 3; k++)
				dL_dscale[3 * (size_t)idx + k] = Rt.m[k][0] * dL_dMt.m[k][0] + Rt.m[k][1] * dL_dMt.m[k][1] + Rt.m[k][2] * dL_dMt.m[k][2];
			for (int k = 0; k < 3; k++) for (int w = 0; w < 3; w++) dL_dMt.m[k][w] *= s[k];
			real (*D)[3] = dL_dMt.m;
			real *dq = dL_drot + 4 * (size_t)idx;
			dq[0] = 2 * z * (D[0][1] - D[1][0]) + 2 * y * (D[2][0] - D[0][2]) + 2 * x * (D[1][2] - D[2][1]);
			dq[1] = 2 * y * (D[1][0] + D[0][1]) + 2 * z * (D[2][0] + D[0][2]) + 2 * r * (D[1][2] - D[2][1]) - 4 * x * (D[2][2] + D[1][1]);
			dq[2] = 2 * x * (D[1][0] + D[0][1]) + 2 * r * (D[2][0] - D[0][2]) + 2 * z * (D[1][2] + D[2][1]) - 4 * y * (D[2][2] + D[0][0]);
			dq[3] = 2 * r * (D[0][1] - D[1][0]) + 2 * x * (D[2][0] + D[0][2]) + 2 * y * (D[1][2] + D[2][1]) - 4 * z * (D[1][1] + D[0][0]);
		}
	}
}

int orc_real_bytes(void) { return (int)sizeof(real); }
