// gsr_preprocess.cu — per-Gaussian kernels: frustum cull + EWA projection (forward), the anchor
// pre-filters, and the fused per-Gaussian backward (cov2D -> cov3D -> scale/rotation, mean paths).
//
// Replaces, in W-Ted/GScream submodules/diff-gaussian-rasterization (CR/ = cuda_rasterizer/):
//   preprocessCUDA fwd            CR/forward.cu:157-267      (+ in_frustum CR/auxiliary.h:139-164,
//                                 computeCov3D :120-154, computeCov2D :76-115, getRect, ndc2Pix)
//   filter_preprocessCUDA         CR/forward.cu:271-346
//   position2D_preprocessCUDA     CR/forward.cu:352-433
//   checkFrustum                  CR/rasterizer_impl.cu:54-66
//   computeCov2DCUDA + preprocessCUDA bwd   CR/backward.cu:144-274, 346-406 (+ computeCov3D :278-341)
//
// Bit-exactness: tile/key indexing must equal the reference's, and keys embed fp32 depth bits and
// rects derived from fp32 radii/centres.  The projection arithmetic below therefore keeps the
// reference's expression trees (GLM's column-major mat3 product order, the literal zeros in J and S,
// the double-precision ndc2Pix) so that nvcc's FMA contraction produces the same fp32 results; it is
// NOT compiled with fast-math.  Everything that is new here (bounding extents for the blend kernels'
// culling) is computed from already-rounded values with explicit intrinsics, so it cannot perturb the
// contraction of the reference expressions.
#include "gsr_internal.cuh"
#include <cstdio>

namespace gsr {

struct M3 { float m[3][3]; }; // m[col][row], as GLM

// GLM operator*(mat3, mat3), third_party/glm/glm/detail/type_mat3x3.inl:486-519
__device__ __forceinline__ M3 m3mul(const M3 &a, const M3 &b)
{
	M3 r;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int w = 0; w < 3; w++)
			r.m[c][w] = a.m[0][w] * b.m[c][0] + a.m[1][w] * b.m[c][1] + a.m[2][w] * b.m[c][2];
	return r;
}
__device__ __forceinline__ M3 m3t(const M3 &a)
{
	M3 r;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int w = 0; w < 3; w++)
			r.m[c][w] = a.m[w][c];
	return r;
}

// CR/auxiliary.h:41-44: the 1.0 / 0.5 literals make this a double-precision evaluation.
__device__ __forceinline__ float ndc2pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

__device__ __forceinline__ float3 xform4x3(const float3 &p, const float *m)
{
	float3 t = {
	    m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
	    m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
	    m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
	};
	return t;
}
__device__ __forceinline__ float4 xform4x4(const float3 &p, const float *m)
{
	float4 t = {
	    m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
	    m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
	    m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
	    m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]};
	return t;
}

__device__ __forceinline__ M3 rotation_of(const float4 q)
{
	// CR/forward.cu:130-140 — quaternion used as given (r,x,y,z), no normalisation (:129).
	// The mixed terms a*b +- c*d are where ptxas (not nvcc's front end) picks which product to keep rounded
	// and which to fuse; the choice below is the one in the reference build's SASS for sm_100a (identical in
	// preprocessCUDA, filter_preprocessCUDA and position2D_preprocessCUDA), pinned with explicit
	// round-to-nearest intrinsics so that it cannot drift with code motion in this kernel.
	const float r = q.x, x = q.y, y = q.z, z = q.w;
	const float xz = __fmul_rn(x, z), rx = __fmul_rn(r, x), rz = __fmul_rn(r, z);
	const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
	const float xz_p_ry = __fmaf_rn(r, y, xz);   // x*z + r*y
	const float xz_m_ry = __fmaf_rn(-r, y, xz);  // x*z - r*y
	const float yz_m_rx = __fmaf_rn(y, z, -rx);  // y*z - r*x
	const float yz_p_rx = __fmaf_rn(y, z, rx);   // y*z + r*x
	const float xy_m_rz = __fmaf_rn(x, y, -rz);  // x*y - r*z
	const float xy_p_rz = __fmaf_rn(x, y, rz);   // x*y + r*z
	const float yy_zz = __fadd_rn(yy, zz);
	const float xx_zz = __fmaf_rn(x, x, zz);
	const float xx_yy = __fmaf_rn(x, x, yy);
	M3 R;
	R.m[0][0] = __fsub_rn(1.f, __fadd_rn(yy_zz, yy_zz)); R.m[0][1] = __fadd_rn(xy_m_rz, xy_m_rz);             R.m[0][2] = __fadd_rn(xz_p_ry, xz_p_ry);
	R.m[1][0] = __fadd_rn(xy_p_rz, xy_p_rz);             R.m[1][1] = __fsub_rn(1.f, __fadd_rn(xx_zz, xx_zz)); R.m[1][2] = __fadd_rn(yz_m_rx, yz_m_rx);
	R.m[2][0] = __fadd_rn(xz_m_ry, xz_m_ry);             R.m[2][1] = __fadd_rn(yz_p_rx, yz_p_rx);             R.m[2][2] = __fsub_rn(1.f, __fadd_rn(xx_yy, xx_yy));
	return R;
}

// CR/forward.cu:120-154
__device__ __forceinline__ void cov3d_of(const float3 scale, float mod, const float4 rot, float *cov3D)
{
	M3 S;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int w = 0; w < 3; w++)
			S.m[c][w] = 0.0f;
	S.m[0][0] = mod * scale.x;
	S.m[1][1] = mod * scale.y;
	S.m[2][2] = mod * scale.z;
	M3 R = rotation_of(rot);
	M3 M = m3mul(S, R);
	M3 Sigma = m3mul(m3t(M), M);
	cov3D[0] = Sigma.m[0][0];
	cov3D[1] = Sigma.m[0][1];
	cov3D[2] = Sigma.m[0][2];
	cov3D[3] = Sigma.m[1][1];
	cov3D[4] = Sigma.m[1][2];
	cov3D[5] = Sigma.m[2][2];
}

// Front half shared by CR/forward.cu:76-115 and CR/backward.cu:155-199.
struct Cov2DCtx {
	float3 t;
	float txtz, tytz, limx, limy;
	M3 W, T, Vrk, cov;
};
__device__ __forceinline__ void cov2d_common(const float3 &mean, float focal_x, float focal_y, float tan_fovx, float tan_fovy,
                                             const float *cov3D, const float *view, Cov2DCtx &c)
{
	float3 t = xform4x3(mean, view);
	const float limx = 1.3f * tan_fovx;
	const float limy = 1.3f * tan_fovy;
	const float txtz = t.x / t.z;
	const float tytz = t.y / t.z;
	t.x = min(limx, max(-limx, txtz)) * t.z;
	t.y = min(limy, max(-limy, tytz)) * t.z;
	M3 J;
	J.m[0][0] = focal_x / t.z; J.m[0][1] = 0.0f;          J.m[0][2] = -(focal_x * t.x) / (t.z * t.z);
	J.m[1][0] = 0.0f;          J.m[1][1] = focal_y / t.z; J.m[1][2] = -(focal_y * t.y) / (t.z * t.z);
	J.m[2][0] = 0;             J.m[2][1] = 0;             J.m[2][2] = 0;
	c.W.m[0][0] = view[0]; c.W.m[0][1] = view[4]; c.W.m[0][2] = view[8];
	c.W.m[1][0] = view[1]; c.W.m[1][1] = view[5]; c.W.m[1][2] = view[9];
	c.W.m[2][0] = view[2]; c.W.m[2][1] = view[6]; c.W.m[2][2] = view[10];
	c.T = m3mul(c.W, J);
	c.Vrk.m[0][0] = cov3D[0]; c.Vrk.m[0][1] = cov3D[1]; c.Vrk.m[0][2] = cov3D[2];
	c.Vrk.m[1][0] = cov3D[1]; c.Vrk.m[1][1] = cov3D[3]; c.Vrk.m[1][2] = cov3D[4];
	c.Vrk.m[2][0] = cov3D[2]; c.Vrk.m[2][1] = cov3D[4]; c.Vrk.m[2][2] = cov3D[5];
	c.cov = m3mul(m3mul(m3t(c.T), m3t(c.Vrk)), c.T);
	c.t = t; c.txtz = txtz; c.tytz = tytz; c.limx = limx; c.limy = limy;
}

// SH evaluation, CR/forward.cu:22-73 (GScream never uses it: shs=None; kept for the drop-in surface).
__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                  -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator*(float s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }

__device__ V3 color_from_sh(int idx, int deg, int max_coeffs, float3 pos, const float *campos, const float *shs, uint8_t *clamped)
{
	V3 dir = {pos.x - campos[0], pos.y - campos[1], pos.z - campos[2]};
	float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
	dir = {dir.x / len, dir.y / len, dir.z / len};
	const V3 *sh = ((const V3 *)shs) + (size_t)idx * max_coeffs;
	V3 result = SH_C0 * sh[0];
	if (deg > 0) {
		float x = dir.x, y = dir.y, z = dir.z;
		result = result - SH_C1 * y * sh[1] + SH_C1 * z * sh[2] - SH_C1 * x * sh[3];
		if (deg > 1) {
			float xx = x * x, yy = y * y, zz = z * z;
			float xy = x * y, yz = y * z, xz = x * z;
			result = result + SH_C2[0] * xy * sh[4] + SH_C2[1] * yz * sh[5] + SH_C2[2] * (2.0f * zz - xx - yy) * sh[6] +
			         SH_C2[3] * xz * sh[7] + SH_C2[4] * (xx - yy) * sh[8];
			if (deg > 2) {
				result = result + SH_C3[0] * y * (3.0f * xx - yy) * sh[9] + SH_C3[1] * xy * z * sh[10] +
				         SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11] + SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12] +
				         SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13] + SH_C3[5] * z * (xx - yy) * sh[14] +
				         SH_C3[6] * x * (xx - 3.0f * yy) * sh[15];
			}
		}
	}
	result = {result.x + 0.5f, result.y + 0.5f, result.z + 0.5f};
	clamped[3 * (size_t)idx + 0] = (result.x < 0);
	clamped[3 * (size_t)idx + 1] = (result.y < 0);
	clamped[3 * (size_t)idx + 2] = (result.z < 0);
	return {fmaxf(result.x, 0.0f), fmaxf(result.y, 0.0f), fmaxf(result.z, 0.0f)};
}

// Cooperative, 128-bit staged read of `rows` consecutive N-float rows starting at row `first`
// (AoS [P][N] fp32 input such as xyz / scales): float4 global loads into shared memory, then each
// thread picks its own row (stride-N shared reads are conflict-free for odd N).
template <int N>
__device__ __forceinline__ void stage_rows(const float *__restrict__ g, int first, int rows, float *s)
{
	const int nfl = rows * N;
	const float *base = g + (size_t)first * N;
	if ((reinterpret_cast<uintptr_t>(base) & 15) == 0) {
		const int nv4 = nfl >> 2;
		const float4 *g4 = reinterpret_cast<const float4 *>(base);
		for (int i = threadIdx.x; i < nv4; i += blockDim.x)
			reinterpret_cast<float4 *>(s)[i] = __ldg(g4 + i);
		for (int i = (nv4 << 2) + threadIdx.x; i < nfl; i += blockDim.x)
			s[i] = __ldg(base + i);
	} else {
		for (int i = threadIdx.x; i < nfl; i += blockDim.x)
			s[i] = __ldg(base + i);
	}
}


// MODE 0: full preprocess (K1); 1: visible_filter (radii only); 2: position2D_filter (radii + pixel x,y)
template <int MODE>
__global__ void __launch_bounds__(256) preprocess_kernel(const PreArgs a)
{
	__shared__ __align__(16) float s_xyz[256 * 3];
	__shared__ __align__(16) float s_scale[256 * 3];
	__shared__ float s_view[16], s_proj[16];

	const int first = blockIdx.x * 256;
	const int rows = min(256, a.P - first);
	stage_rows<3>(a.means3D, first, rows, s_xyz);
	if (a.cov3D_precomp == nullptr) {
		if (a.scales_stride == 3) {
			stage_rows<3>(a.scales, first, rows, s_scale);
		} else if ((int)threadIdx.x < rows) {
			// a row-strided view: the anchor filters are called with `get_scaling[:, :3]` of a [A, 6] tensor
			// (gaussian_renderer/__init__.py:298); read in place instead of through a contiguous copy
			const float *row = a.scales + (size_t)(first + threadIdx.x) * a.scales_stride;
			s_scale[3 * threadIdx.x] = __ldg(row); s_scale[3 * threadIdx.x + 1] = __ldg(row + 1); s_scale[3 * threadIdx.x + 2] = __ldg(row + 2);
		}
	}
	if (threadIdx.x < 16) s_view[threadIdx.x] = a.view[threadIdx.x];
	else if (threadIdx.x < 32) s_proj[threadIdx.x - 16] = a.proj[threadIdx.x - 16];
	__syncthreads();

	const int idx = first + threadIdx.x;
	if (idx >= a.P) return;

	// Defaults for a culled Gaussian (CR/forward.cu:191-197, 385-390)
	int out_radius = 0;
	uint32_t out_tiles = 0;
	uint32_t out_key = 0xFFFFFFFFu; // culled Gaussians sort behind every visible one
	float out_px = 0.f, out_py = 0.f;

	const float3 p_orig = {s_xyz[3 * threadIdx.x], s_xyz[3 * threadIdx.x + 1], s_xyz[3 * threadIdx.x + 2]};
	const float *viewmatrix = s_view, *projmatrix = s_proj;

	// in_frustum, CR/auxiliary.h:139-164: only the near test survives (z <= 0.2 -> culled)
	float3 p_view = xform4x3(p_orig, viewmatrix);
	bool alive = !(p_view.z <= 0.2f);
	if (!alive && a.prefiltered) {
		printf("Point is filtered although prefiltered is set. This shouldn't happen!");
		__trap();
	}

	if (alive) {
		float4 p_hom = xform4x4(p_orig, projmatrix);
		float p_w = 1.0f / (p_hom.w + 0.0000001f);
		float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};

		float cov3D[6];
		if (a.cov3D_precomp != nullptr) {
#pragma unroll
			for (int k = 0; k < 6; k++) cov3D[k] = a.cov3D_precomp[6 * (size_t)idx + k];
		} else {
			const float3 sc = {s_scale[3 * threadIdx.x], s_scale[3 * threadIdx.x + 1], s_scale[3 * threadIdx.x + 2]};
			const float4 q = __ldg(reinterpret_cast<const float4 *>(a.rotations) + idx);
			cov3d_of(sc, a.scale_modifier, q, cov3D);
		}

		Cov2DCtx c;
		cov2d_common(p_orig, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy, cov3D, viewmatrix, c);
		// low-pass: every Gaussian at least ~one pixel wide (CR/forward.cu:112-113)
		c.cov.m[0][0] += 0.3f;
		c.cov.m[1][1] += 0.3f;
		const float3 cov = {float(c.cov.m[0][0]), float(c.cov.m[0][1]), float(c.cov.m[1][1])};

		// det and mid^2 - det: fused as in the reference build's SASS (FMUL cov.y^2 ; FFMA cov.x*cov.z - .)
		float det = __fmaf_rn(cov.x, cov.z, -__fmul_rn(cov.y, cov.y));
		if (det != 0.0f) {
			float det_inv = 1.f / det;
			float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};

			float mid = 0.5f * (cov.x + cov.z);
			const float disc = max(0.1f, __fmaf_rn(mid, mid, -det));
			float lambda1 = mid + sqrt(disc);
			float lambda2 = mid - sqrt(disc);
			float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
			float2 point_image = {ndc2pix(p_proj.x, a.W), ndc2pix(p_proj.y, a.H)};
			int x0, y0, x1, y1;
			get_rect(point_image.x, point_image.y, (int)my_radius, a.gx, a.gy, x0, y0, x1, y1);
			if ((x1 - x0) * (y1 - y0) != 0) {
				out_radius = (int)my_radius;
				out_px = point_image.x;
				out_py = point_image.y;
				if (MODE == 0) {
					out_tiles = (uint32_t)((y1 - y0) * (x1 - x0));
					out_key = __float_as_uint(p_view.z);
					const float opacity = __ldg(a.opacities + idx);
					const float unc = __ldg(a.uncertainties + idx);

					// --- new: conservative half extents of {alpha >= 1/255} for in-kernel culling ---
					// alpha = opacity * exp(-0.5 d^T Q d) >= 1/255  <=>  d^T Q d <= 2 ln(255 opacity) =: tau,
					// Q = fp32 conic actually used by the blend kernels.  bbox half extent along x is
					// sqrt(tau * (Q^-1)_xx) = sqrt(tau * c / (ac - b^2)); evaluated in double from the
					// rounded conic so that cancellation in (ac - b^2) cannot make the box too small.
					float hx = -1.f, hy = -1.f; // negative: nothing can contribute
					float pow_min = __int_as_float(0xff800000); // -inf: "power >= pow_min" never culls
					if (!(opacity == opacity)) { // NaN opacity: never cull
						hx = hy = __int_as_float(0x7f800000);
					} else if (opacity >= 1.0f / 255.0f) {
						const double qa = conic.x, qb = conic.y, qc = conic.z;
						const double dq = qa * qc - qb * qb;
						// necessary condition for alpha >= 1/255: power >= -log(255 opacity), with slack
						pow_min = __double2float_rd(-(log(255.0 * (double)opacity) * 1.0005 + 0.005));
						if (dq > 0.0 && qa > 0.0 && qc > 0.0) {
							const double tau = 2.0 * log(255.0 * (double)opacity) * 1.001 + 0.02;
							hx = __double2float_ru(sqrt(tau * qc / dq) * 1.0005 + 0.02);
							hy = __double2float_ru(sqrt(tau * qa / dq) * 1.0005 + 0.02);
						} else {
							hx = hy = __int_as_float(0x7f800000);
						}
					}

					float cr = 0.f, cg = 0.f, cb = 0.f;
					if (a.colors_precomp == nullptr) {
						V3 col = color_from_sh(idx, a.D, a.M, p_orig, a.campos, a.shs, a.clamped);
						cr = col.x; cg = col.y; cb = col.z;
						a.rgb[3 * (size_t)idx + 0] = cr; a.rgb[3 * (size_t)idx + 1] = cg; a.rgb[3 * (size_t)idx + 2] = cb;
					} else if (a.C <= 3) {
						cr = __ldg(a.colors_precomp + (size_t)idx * a.C);
						if (a.C > 1) cg = __ldg(a.colors_precomp + (size_t)idx * a.C + 1);
						if (a.C > 2) cb = __ldg(a.colors_precomp + (size_t)idx * a.C + 2);
					}
					float4 *r4 = reinterpret_cast<float4 *>(a.rec + (size_t)idx * GSR_REC_FLOATS);
					r4[0] = make_float4(point_image.x, point_image.y, conic.x, conic.y);
					r4[1] = make_float4(conic.z, opacity, p_view.z, unc);
					r4[2] = make_float4(hx, hy, cr, cg);
					r4[3] = make_float4(cb, my_radius, pow_min, 0.f); // [13] radius (exact in fp32) for the instance emission
				}
			}
		}
	}

	a.radii[idx] = out_radius;
	if (MODE == 0) {
		a.tiles_touched[idx] = out_tiles;
		a.depth_key[idx] = out_key;
		a.depth_val[idx] = (uint32_t)idx;
	}
	if (MODE == 2) {
		a.pos_x[idx] = out_px;
		a.pos_y[idx] = out_py;
	}
}

// K12 checkFrustum (CR/rasterizer_impl.cu:54-66)
__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float *__restrict__ means3D, const float *__restrict__ view,
                                                           uint8_t *__restrict__ present)
{
	__shared__ __align__(16) float s_xyz[256 * 3];
	const int first = blockIdx.x * 256;
	const int rows = min(256, P - first);
	stage_rows<3>(means3D, first, rows, s_xyz);
	__syncthreads();
	const int idx = first + threadIdx.x;
	if (idx >= P) return;
	const float3 p = {s_xyz[3 * threadIdx.x], s_xyz[3 * threadIdx.x + 1], s_xyz[3 * threadIdx.x + 2]};
	float3 p_view = xform4x3(p, view);
	present[idx] = !(p_view.z <= 0.2f);
}

// ------------------------------------------------------------------------------------------------
// Fused per-Gaussian backward: computeCov2DCUDA (CR/backward.cu:144-274) + preprocessCUDA backward
// (:346-406) + computeCov3D backward (:278-341) in ONE pass over the Gaussians (the reference runs two
// kernels and round-trips dL_dcov3D / dL_dmeans through HBM between them).
// gacc[P][8] holds what the blend backward accumulated: {dmean2D.x, dmean2D.y, dconic.x, dconic.y,
// dconic.w, dopacity, ddepth, duncertainty}.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void put(float *p, float v, int accumulate)
{
	if (p == nullptr) return;
	if (accumulate) *p += v; else *p = v;
}

__device__ float3 dnormvdv3(float3 v, float3 dv)
{
	float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
	float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
	float3 r;
	r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
	r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
	r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
	return r;
}

// SH backward, CR/backward.cu:20-139.  Returns dL_dmean contribution; writes dL_dsh.
__device__ float3 color_from_sh_bwd(int idx, int deg, int max_coeffs, float3 pos, const float *campos, const float *shs,
                                    const uint8_t *clamped, V3 dL_dRGB, float *dL_dshs, int accumulate)
{
	V3 dir_orig = {pos.x - campos[0], pos.y - campos[1], pos.z - campos[2]};
	float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
	V3 dir = {dir_orig.x / len, dir_orig.y / len, dir_orig.z / len};
	const V3 *sh = ((const V3 *)shs) + (size_t)idx * max_coeffs;
	dL_dRGB.x *= clamped[3 * (size_t)idx + 0] ? 0 : 1;
	dL_dRGB.y *= clamped[3 * (size_t)idx + 1] ? 0 : 1;
	dL_dRGB.z *= clamped[3 * (size_t)idx + 2] ? 0 : 1;
	V3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
	float x = dir.x, y = dir.y, z = dir.z;
	float *out = dL_dshs + (size_t)idx * max_coeffs * 3;
	auto store = [&](int k, float w) {
		put(out + 3 * k + 0, w * dL_dRGB.x, accumulate);
		put(out + 3 * k + 1, w * dL_dRGB.y, accumulate);
		put(out + 3 * k + 2, w * dL_dRGB.z, accumulate);
	};
	store(0, SH_C0);
	if (deg > 0) {
		store(1, -SH_C1 * y);
		store(2, SH_C1 * z);
		store(3, -SH_C1 * x);
		dRGBdx = -SH_C1 * sh[3];
		dRGBdy = -SH_C1 * sh[1];
		dRGBdz = SH_C1 * sh[2];
		if (deg > 1) {
			float xx = x * x, yy = y * y, zz = z * z;
			float xy = x * y, yz = y * z, xz = x * z;
			store(4, SH_C2[0] * xy);
			store(5, SH_C2[1] * yz);
			store(6, SH_C2[2] * (2.f * zz - xx - yy));
			store(7, SH_C2[3] * xz);
			store(8, SH_C2[4] * (xx - yy));
			dRGBdx = dRGBdx + (SH_C2[0] * y * sh[4] + SH_C2[2] * 2.f * -x * sh[6] + SH_C2[3] * z * sh[7] + SH_C2[4] * 2.f * x * sh[8]);
			dRGBdy = dRGBdy + (SH_C2[0] * x * sh[4] + SH_C2[1] * z * sh[5] + SH_C2[2] * 2.f * -y * sh[6] + SH_C2[4] * 2.f * -y * sh[8]);
			dRGBdz = dRGBdz + (SH_C2[1] * y * sh[5] + SH_C2[2] * 2.f * 2.f * z * sh[6] + SH_C2[3] * x * sh[7]);
			if (deg > 2) {
				store(9, SH_C3[0] * y * (3.f * xx - yy));
				store(10, SH_C3[1] * xy * z);
				store(11, SH_C3[2] * y * (4.f * zz - xx - yy));
				store(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
				store(13, SH_C3[4] * x * (4.f * zz - xx - yy));
				store(14, SH_C3[5] * z * (xx - yy));
				store(15, SH_C3[6] * x * (xx - 3.f * yy));
				dRGBdx = dRGBdx + (SH_C3[0] * 3.f * 2.f * xy * sh[9] + SH_C3[1] * yz * sh[10] + SH_C3[2] * -2.f * xy * sh[11] +
				                   SH_C3[3] * -3.f * 2.f * xz * sh[12] + SH_C3[4] * (-3.f * xx + 4.f * zz - yy) * sh[13] +
				                   SH_C3[5] * 2.f * xz * sh[14] + SH_C3[6] * 3.f * (xx - yy) * sh[15]);
				dRGBdy = dRGBdy + (SH_C3[0] * 3.f * (xx - yy) * sh[9] + SH_C3[1] * xz * sh[10] + SH_C3[2] * (-3.f * yy + 4.f * zz - xx) * sh[11] +
				                   SH_C3[3] * -3.f * 2.f * yz * sh[12] + SH_C3[4] * -2.f * xy * sh[13] + SH_C3[5] * -2.f * yz * sh[14] +
				                   SH_C3[6] * -3.f * 2.f * xy * sh[15]);
				dRGBdz = dRGBdz + (SH_C3[1] * xy * sh[10] + SH_C3[2] * 4.f * 2.f * yz * sh[11] + SH_C3[3] * 3.f * (2.f * zz - xx - yy) * sh[12] +
				                   SH_C3[4] * 4.f * 2.f * xz * sh[13] + SH_C3[5] * (xx - yy) * sh[14]);
			}
		}
	}
	float3 dL_ddir = {dRGBdx.x * dL_dRGB.x + dRGBdx.y * dL_dRGB.y + dRGBdx.z * dL_dRGB.z,
	                  dRGBdy.x * dL_dRGB.x + dRGBdy.y * dL_dRGB.y + dRGBdy.z * dL_dRGB.z,
	                  dRGBdz.x * dL_dRGB.x + dRGBdz.y * dL_dRGB.y + dRGBdz.z * dL_dRGB.z};
	return dnormvdv3(float3{dir_orig.x, dir_orig.y, dir_orig.z}, dL_ddir);
}

__global__ void __launch_bounds__(256) preprocess_backward_kernel(const PreBwdArgs a)
{
	__shared__ __align__(16) float s_xyz[256 * 3];
	__shared__ __align__(16) float s_scale[256 * 3];
	__shared__ float s_view[16], s_proj[16];
	const int first = blockIdx.x * 256;
	const int rows = min(256, a.P - first);
	stage_rows<3>(a.means3D, first, rows, s_xyz);
	if (a.scales != nullptr) stage_rows<3>(a.scales, first, rows, s_scale);
	if (threadIdx.x < 16) s_view[threadIdx.x] = a.view[threadIdx.x];
	else if (threadIdx.x < 32) s_proj[threadIdx.x - 16] = a.proj[threadIdx.x - 16];
	__syncthreads();
	const int idx = first + threadIdx.x;
	if (idx >= a.P) return;
	const float *view = s_view, *proj = s_proj;
	const size_t i = (size_t)idx;

	float3 dmean = {0.f, 0.f, 0.f};
	float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
	float3 dscale = {0.f, 0.f, 0.f};
	float4 drot = {0.f, 0.f, 0.f, 0.f};
	float g_m2x = 0.f, g_m2y = 0.f, g_op = 0.f, g_unc = 0.f;

	if (a.radii[idx] > 0) {
		const float4 ga = __ldg(reinterpret_cast<const float4 *>(a.gacc) + 2 * i);
		const float4 gb = __ldg(reinterpret_cast<const float4 *>(a.gacc) + 2 * i + 1);
		g_m2x = ga.x; g_m2y = ga.y;
		const float3 dL_dconic = {ga.z, ga.w, gb.x};
		g_op = gb.y;
		const float dL_ddepth = gb.z;
		g_unc = gb.w;

		const float3 mean = {s_xyz[3 * threadIdx.x], s_xyz[3 * threadIdx.x + 1], s_xyz[3 * threadIdx.x + 2]};
		float cov3D[6];
		float3 sc = {0.f, 0.f, 0.f};
		float4 q = {0.f, 0.f, 0.f, 0.f};
		if (a.cov3D_precomp != nullptr) {
#pragma unroll
			for (int k = 0; k < 6; k++) cov3D[k] = a.cov3D_precomp[6 * i + k];
		} else {
			sc = {s_scale[3 * threadIdx.x], s_scale[3 * threadIdx.x + 1], s_scale[3 * threadIdx.x + 2]};
			q = __ldg(reinterpret_cast<const float4 *>(a.rotations) + idx);
			cov3d_of(sc, a.scale_modifier, q, cov3D); // recomputed, bit-identical to the forward's
		}

		// ---- computeCov2DCUDA, CR/backward.cu:155-273 ----
		Cov2DCtx c;
		cov2d_common(mean, a.h_x, a.h_y, a.tan_fovx, a.tan_fovy, cov3D, view, c);
		const float x_grad_mul = c.txtz < -c.limx || c.txtz > c.limx ? 0 : 1;
		const float y_grad_mul = c.tytz < -c.limy || c.tytz > c.limy ? 0 : 1;
		const M3 &T = c.T, &Vrk = c.Vrk, &W = c.W;
		float ca = c.cov.m[0][0] + 0.3f;
		float cb = c.cov.m[0][1];
		float cc = c.cov.m[1][1] + 0.3f;
		float denom = ca * cc - cb * cb;
		float dL_da = 0, dL_db = 0, dL_dc = 0;
		float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
		if (denom2inv != 0) {
			dL_da = denom2inv * (-cc * cc * dL_dconic.x + 2 * cb * cc * dL_dconic.y + (denom - ca * cc) * dL_dconic.z);
			dL_dc = denom2inv * (-ca * ca * dL_dconic.z + 2 * ca * cb * dL_dconic.y + (denom - ca * cc) * dL_dconic.x);
			dL_db = denom2inv * 2 * (cb * cc * dL_dconic.x - (denom + 2 * cb * cb) * dL_dconic.y + ca * cb * dL_dconic.z);
			dcov[0] = (T.m[0][0] * T.m[0][0] * dL_da + T.m[0][0] * T.m[1][0] * dL_db + T.m[1][0] * T.m[1][0] * dL_dc);
			dcov[3] = (T.m[0][1] * T.m[0][1] * dL_da + T.m[0][1] * T.m[1][1] * dL_db + T.m[1][1] * T.m[1][1] * dL_dc);
			dcov[5] = (T.m[0][2] * T.m[0][2] * dL_da + T.m[0][2] * T.m[1][2] * dL_db + T.m[1][2] * T.m[1][2] * dL_dc);
			dcov[1] = 2 * T.m[0][0] * T.m[0][1] * dL_da + (T.m[0][0] * T.m[1][1] + T.m[0][1] * T.m[1][0]) * dL_db + 2 * T.m[1][0] * T.m[1][1] * dL_dc;
			dcov[2] = 2 * T.m[0][0] * T.m[0][2] * dL_da + (T.m[0][0] * T.m[1][2] + T.m[0][2] * T.m[1][0]) * dL_db + 2 * T.m[1][0] * T.m[1][2] * dL_dc;
			dcov[4] = 2 * T.m[0][2] * T.m[0][1] * dL_da + (T.m[0][1] * T.m[1][2] + T.m[0][2] * T.m[1][1]) * dL_db + 2 * T.m[1][1] * T.m[1][2] * dL_dc;
		}
		float dL_dT00 = 2 * (T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_da +
		                (T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_db;
		float dL_dT01 = 2 * (T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_da +
		                (T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_db;
		float dL_dT02 = 2 * (T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_da +
		                (T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_db;
		float dL_dT10 = 2 * (T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_dc +
		                (T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_db;
		float dL_dT11 = 2 * (T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_dc +
		                (T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_db;
		float dL_dT12 = 2 * (T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_dc +
		                (T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_db;
		float dL_dJ00 = W.m[0][0] * dL_dT00 + W.m[0][1] * dL_dT01 + W.m[0][2] * dL_dT02;
		float dL_dJ02 = W.m[2][0] * dL_dT00 + W.m[2][1] * dL_dT01 + W.m[2][2] * dL_dT02;
		float dL_dJ11 = W.m[1][0] * dL_dT10 + W.m[1][1] * dL_dT11 + W.m[1][2] * dL_dT12;
		float dL_dJ12 = W.m[2][0] * dL_dT10 + W.m[2][1] * dL_dT11 + W.m[2][2] * dL_dT12;
		float tz = 1.f / c.t.z;
		float tz2 = tz * tz;
		float tz3 = tz2 * tz;
		float dL_dtx = x_grad_mul * -a.h_x * tz2 * dL_dJ02;
		float dL_dty = y_grad_mul * -a.h_y * tz2 * dL_dJ12;
		float dL_dtz = -a.h_x * tz2 * dL_dJ00 - a.h_y * tz2 * dL_dJ11 + (2 * a.h_x * c.t.x) * tz3 * dL_dJ02 + (2 * a.h_y * c.t.y) * tz3 * dL_dJ12;
		// transformVec4x3Transpose, CR/auxiliary.h:90-97
		dmean = {view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz,
		         view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz,
		         view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz};

		// ---- preprocessCUDA backward, CR/backward.cu:368-405 ----
		float4 m_hom = xform4x4(mean, proj);
		float m_w = 1.0f / (m_hom.w + 0.0000001f);
		float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
		float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
		dmean.x += (proj[0] * m_w - proj[3] * mul1) * g_m2x + (proj[1] * m_w - proj[3] * mul2) * g_m2y;
		dmean.y += (proj[4] * m_w - proj[7] * mul1) * g_m2x + (proj[5] * m_w - proj[7] * mul2) * g_m2y;
		dmean.z += (proj[8] * m_w - proj[11] * mul1) * g_m2x + (proj[9] * m_w - proj[11] * mul2) * g_m2y;
		dmean.x += view[2] * dL_ddepth;
		dmean.y += view[6] * dL_ddepth;
		dmean.z += view[10] * dL_ddepth;

		if (a.shs != nullptr) {
			// dL_dcolors for the SH path was accumulated into dL_dcolors[P,3] by the blend backward
			V3 dL_dRGB = {a.dL_dcolors[3 * i], a.dL_dcolors[3 * i + 1], a.dL_dcolors[3 * i + 2]};
			float3 dm = color_from_sh_bwd(idx, a.D, a.M, mean, a.campos, a.shs, a.clamped, dL_dRGB, a.dL_dsh, a.accumulate);
			dmean.x += dm.x; dmean.y += dm.y; dmean.z += dm.z;
		}

		if (a.cov3D_precomp == nullptr) {
			// computeCov3D backward, CR/backward.cu:278-341 (gradient w.r.t. the UN-normalised quaternion, :340)
			float r = q.x, x = q.y, y = q.z, z = q.w;
			M3 R = rotation_of(q);
			M3 S;
#pragma unroll
			for (int cI = 0; cI < 3; cI++)
#pragma unroll
				for (int w = 0; w < 3; w++) S.m[cI][w] = 0.0f;
			float3 s = {a.scale_modifier * sc.x, a.scale_modifier * sc.y, a.scale_modifier * sc.z};
			S.m[0][0] = s.x; S.m[1][1] = s.y; S.m[2][2] = s.z;
			M3 M = m3mul(S, R);
			M3 dL_dSigma;
			dL_dSigma.m[0][0] = dcov[0];        dL_dSigma.m[0][1] = 0.5f * dcov[1]; dL_dSigma.m[0][2] = 0.5f * dcov[2];
			dL_dSigma.m[1][0] = 0.5f * dcov[1]; dL_dSigma.m[1][1] = dcov[3];        dL_dSigma.m[1][2] = 0.5f * dcov[4];
			dL_dSigma.m[2][0] = 0.5f * dcov[2]; dL_dSigma.m[2][1] = 0.5f * dcov[4]; dL_dSigma.m[2][2] = dcov[5];
			M3 M2;
#pragma unroll
			for (int cI = 0; cI < 3; cI++)
#pragma unroll
				for (int w = 0; w < 3; w++) M2.m[cI][w] = 2.0f * M.m[cI][w];
			M3 dL_dM = m3mul(M2, dL_dSigma);
			M3 Rt = m3t(R);
			M3 D = m3t(dL_dM);
			dscale.x = Rt.m[0][0] * D.m[0][0] + Rt.m[0][1] * D.m[0][1] + Rt.m[0][2] * D.m[0][2];
			dscale.y = Rt.m[1][0] * D.m[1][0] + Rt.m[1][1] * D.m[1][1] + Rt.m[1][2] * D.m[1][2];
			dscale.z = Rt.m[2][0] * D.m[2][0] + Rt.m[2][1] * D.m[2][1] + Rt.m[2][2] * D.m[2][2];
#pragma unroll
			for (int w = 0; w < 3; w++) { D.m[0][w] *= s.x; D.m[1][w] *= s.y; D.m[2][w] *= s.z; }
			drot.x = 2 * z * (D.m[0][1] - D.m[1][0]) + 2 * y * (D.m[2][0] - D.m[0][2]) + 2 * x * (D.m[1][2] - D.m[2][1]);
			drot.y = 2 * y * (D.m[1][0] + D.m[0][1]) + 2 * z * (D.m[2][0] + D.m[0][2]) + 2 * r * (D.m[1][2] - D.m[2][1]) - 4 * x * (D.m[2][2] + D.m[1][1]);
			drot.z = 2 * x * (D.m[1][0] + D.m[0][1]) + 2 * r * (D.m[2][0] - D.m[0][2]) + 2 * z * (D.m[1][2] + D.m[2][1]) - 4 * y * (D.m[2][2] + D.m[0][0]);
			drot.w = 2 * r * (D.m[0][1] - D.m[1][0]) + 2 * x * (D.m[2][0] + D.m[0][2]) + 2 * y * (D.m[1][2] + D.m[2][1]) - 4 * z * (D.m[1][1] + D.m[0][0]);
		}
	}

	const int acc = a.accumulate;
	put(a.dL_dmeans3D + 3 * i + 0, dmean.x, acc);
	put(a.dL_dmeans3D + 3 * i + 1, dmean.y, acc);
	put(a.dL_dmeans3D + 3 * i + 2, dmean.z, acc);
	put(a.dL_dmeans2D + 3 * i + 0, g_m2x, acc);
	put(a.dL_dmeans2D + 3 * i + 1, g_m2y, acc);
	put(a.dL_dmeans2D + 3 * i + 2, 0.f, acc);
	put(a.dL_dopacity + i, g_op, acc);
	put(a.dL_duncertainty + i, g_unc, acc);
	if (a.dL_dcov3D != nullptr) {
#pragma unroll
		for (int k = 0; k < 6; k++) put(a.dL_dcov3D + 6 * i + k, dcov[k], acc);
	}
	if (a.dL_dscales != nullptr) {
		put(a.dL_dscales + 3 * i + 0, dscale.x, acc);
		put(a.dL_dscales + 3 * i + 1, dscale.y, acc);
		put(a.dL_dscales + 3 * i + 2, dscale.z, acc);
	}
	if (a.dL_drotations != nullptr) {
		put(a.dL_drotations + 4 * i + 0, drot.x, acc);
		put(a.dL_drotations + 4 * i + 1, drot.y, acc);
		put(a.dL_drotations + 4 * i + 2, drot.z, acc);
		put(a.dL_drotations + 4 * i + 3, drot.w, acc);
	}
}

// ---- host launchers -------------------------------------------------------------------------
cudaError_t launch_preprocess(int mode, const PreArgs &a, cudaStream_t stream)
{
	if (a.P <= 0) return cudaSuccess;
	const int blocks = (a.P + 255) / 256;
	if (mode == 0) preprocess_kernel<0><<<blocks, 256, 0, stream>>>(a);
	else if (mode == 1) preprocess_kernel<1><<<blocks, 256, 0, stream>>>(a);
	else preprocess_kernel<2><<<blocks, 256, 0, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}
cudaError_t launch_mark_visible(int P, const float *means3D, const float *view, uint8_t *present, cudaStream_t stream)
{
	if (P <= 0) return cudaSuccess;
	mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, view, present);
	count_launch();
	return cudaGetLastError();
}
cudaError_t launch_preprocess_backward(const PreBwdArgs &a, cudaStream_t stream)
{
	if (a.P <= 0) return cudaSuccess;
	preprocess_backward_kernel<<<(a.P + 255) / 256, 256, 0, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}

} // namespace gsr
