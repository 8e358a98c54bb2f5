// gsr_torch_glue.cpp — compiled torch glue above the C ABI: the pybind11 module the reference calls `_C`.
//
// W-Ted/GScream's rasterizer reaches its CUDA through five pybind11 functions (submodules/diff-gaussian-rasterization/ext.cpp:16-20)
// implemented in rasterize_points.cu:35-373 on libtorch.  This file is their B200 counterpart: the same five names (typos included),
// the same positional arguments and the same return tuples, but every byte of compute goes through include/gsr_b200.h
// (libgsr_b200.so); libtorch only owns memory and names the current stream.  gscream_b200/_C.py is the same glue over ctypes and
// stays the default; `GSR_GLUE=cpp` selects this module (gscream_b200/_glue.py builds and loads it).
//
// Differences from the reference glue, all behind the same interface:
//   * one synchronisation per forward, on the caller's stream only (the reference blocks the device in cudaMemcpy on the legacy
//     default stream, CR/rasterizer_impl.cu:287); num_rendered arrives in a pinned int64 written asynchronously by stage 1;
//   * scratch is sized by gsr_*_bytes() instead of resize lambdas (rasterize_points.cu:27-33, 77-82);
//   * gradient tensors are torch::empty — the library writes every element, zeros for culled Gaussians — instead of nine
//     zero-fills (rasterize_points.cu:160-170); only dL_dsh (M > 0) and the scale / rotation gradients of the cov3D_precomp
//     path, which the library does not touch, are zero-filled;
//   * the anchor filters allocate no scratch at all (rasterize_points.cu:262-281 allocates geometry + image buffers).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <map>
#include <ATen/cuda/CUDAEvent.h>
#include <mutex>
#include <tuple>

#include "../../include/gsr_b200.h"

namespace {

using torch::Tensor;

void check_rc(int rc)
{
	if (rc == 0) return;
	// GSR_E_SH_CHANNELS carries the reference's own message (CR/rasterizer_impl.cu:246-249); both surface as RuntimeError
	TORCH_CHECK(false, rc == GSR_E_SH_CHANNELS ? "" : "libgsr_b200 call failed: ", gsr_error_string(rc));
}

void check_means(const Tensor &means3D)
{
	if (means3D.ndimension() != 2 || means3D.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)"); // rasterize_points.cu:58-60
}

// "not provided" travels as an empty tensor (diff_gaussian_rasterization/__init__.py:230-240) and becomes a null pointer
struct F32 {
	Tensor t;
	F32(const Tensor &src, const char *name)
	{
		if (src.numel() == 0) return;
		TORCH_CHECK_TYPE(src.scalar_type() == torch::kFloat32, name, " must be float32");
		TORCH_CHECK_VALUE(src.is_cuda(), name, " must be a CUDA tensor");
		t = src.contiguous();
	}
	const float *ptr() const { return t.defined() ? t.data_ptr<float>() : nullptr; }
};

gsr_stream_t current_stream() { return (gsr_stream_t)at::cuda::getCurrentCUDAStream().stream(); }

// a pinned int64 for one call's num_rendered: slots rotate through a small per-device ring, so that concurrent calls on one
// device never share a counter (a call waits for its own value before it returns, long before the ring wraps)
constexpr int kPinnedSlots = 64;
int64_t *pinned_counter(int device)
{
	static std::mutex mu;
	static std::map<int, std::pair<Tensor, int64_t>> rings;
	std::lock_guard<std::mutex> lock(mu);
	auto it = rings.find(device);
	if (it == rings.end())
		it = rings.emplace(device, std::make_pair(torch::zeros({kPinnedSlots}, torch::dtype(torch::kInt64).pinned_memory(true)), (int64_t)0)).first;
	return it->second.first.data_ptr<int64_t>() + (it->second.second++ % kPinnedSlots);
}

// instance capacity to size the binning buffer with before num_rendered has reached the host: 25 % above the (slowly decaying)
// running maximum of what this problem shape produced so far; -1: no history yet
std::mutex g_hint_mu;
std::map<std::tuple<int, int, int, int>, double> g_hint;
// (keyed by the magnitude of P, not P itself: in training P changes a little every iteration and num_rendered follows it smoothly)
int magnitude(int P) { int b = 0; while (P > 0) { b++; P >>= 1; } return b; }
int64_t capacity_guess(int device, int P, int W, int H)
{
	std::lock_guard<std::mutex> lock(g_hint_mu);
	auto it = g_hint.find(std::make_tuple(device, magnitude(P), W, H));
	return it == g_hint.end() ? -1 : (int64_t)(it->second * 1.25) + 65536;
}
void capacity_update(int device, int P, int W, int H, int64_t R)
{
	std::lock_guard<std::mutex> lock(g_hint_mu);
	double &h = g_hint[std::make_tuple(device, magnitude(P), W, H)];
	h = std::max((double)R, 0.98 * h);
}

Tensor bytes(size_t n, const Tensor &like) { return torch::empty({(int64_t)n}, like.options().dtype(torch::kUInt8)); }

// ---- RasterizeGaussiansCUDA, rasterize_points.cu:35-122 ---------------------------------------------------------------
std::tuple<int64_t, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> rasterize_gaussians(
    const Tensor &background, const Tensor &means3D, const Tensor &colors, const Tensor &opacity, const Tensor &uncertaintys,
    const Tensor &scales, const Tensor &rotations, const float scale_modifier, const Tensor &cov3D_precomp, const Tensor &viewmatrix,
    const Tensor &projmatrix, const float tan_fovx, const float tan_fovy, const int image_height, const int image_width,
    const Tensor &sh, const int degree, const Tensor &campos, const bool prefiltered, const bool debug)
{
	check_means(means3D);
	const int P = (int)means3D.size(0), H = image_height, W = image_width;
	const bool has_colors = colors.numel() != 0;
	const int C = has_colors ? (int)colors.size(1) : 3;
	const int M = sh.numel() != 0 ? (int)sh.size(1) : 0;
	const auto f32 = means3D.options().dtype(torch::kFloat32);
	if (P == 0) { // rasterize_points.cu:85: nothing is launched, the images are the reference's torch::full(0)
		Tensor e = bytes(0, means3D);
		return std::make_tuple((int64_t)0, torch::zeros({C, H, W}, f32), torch::zeros({1, H, W}, f32), torch::zeros({1, H, W}, f32),
		                       torch::zeros({0}, means3D.options().dtype(torch::kInt32)), e, e.clone(), e.clone());
	}
	// every element of the four outputs is written by the kernels (the reference zero-fills them first, rasterize_points.cu:69-72)
	Tensor out_color = torch::empty({C, H, W}, f32), out_depth = torch::empty({1, H, W}, f32), out_unc = torch::empty({1, H, W}, f32);
	Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
	TORCH_CHECK_VALUE(means3D.is_cuda(), "means3D must be a CUDA tensor (there is no CPU rasterizer)");
	const c10::cuda::CUDAGuard guard(means3D.device());
	const F32 m3(means3D, "means3D"), col(colors, "colors"), op(opacity, "opacity"), un(uncertaintys, "uncertainties"), sc(scales, "scales"),
	    ro(rotations, "rotations"), cov(cov3D_precomp, "cov3D_precomp"), view(viewmatrix, "viewmatrix"), proj(projmatrix, "projmatrix"),
	    cam(campos, "campos"), bg(background, "bg"), shs(sh, "sh");
	Tensor geom = bytes(gsr_geom_bytes(P), means3D), img = bytes(gsr_image_bytes(W, H), means3D);
	int64_t *R_host = pinned_counter(means3D.get_device());
	const gsr_stream_t stream = current_stream();
	check_rc(gsr_forward_stage1(P, C, degree, M, m3.ptr(), shs.ptr(), col.ptr(), op.ptr(), un.ptr(), sc.ptr(), scale_modifier, ro.ptr(),
	                            cov.ptr(), view.ptr(), proj.ptr(), cam.ptr(), W, H, tan_fovx, tan_fovy, prefiltered ? 1 : 0,
	                            radii.data_ptr<int>(), geom.data_ptr(), (size_t)geom.numel(), R_host, stream));
	// The reference returns num_rendered and sizes the binning buffer from it: one blocking copy in the middle of the forward
	// (CR/rasterizer_impl.cu:287).  Here the second half is launched right behind the first with a buffer sized from an estimate
	// (its kernels read num_rendered from device memory) and the host waits for stage 1's counter only; a wrong estimate costs
	// one repeat of the second half.
	at::cuda::CUDAEvent counted;
	counted.record(at::cuda::getCurrentCUDAStream());
	auto second_half = [&](int64_t num_rendered, int64_t capacity) {
		Tensor buf = bytes(gsr_binning_bytes(P, capacity, W, H), means3D);
		check_rc(gsr_forward_stage2(P, C, num_rendered, col.ptr(), bg.ptr(), W, H, geom.data_ptr(), (size_t)geom.numel(), buf.data_ptr(),
		                            (size_t)buf.numel(), img.data_ptr(), (size_t)img.numel(), out_color.data_ptr<float>(),
		                            out_depth.data_ptr<float>(), out_unc.data_ptr<float>(), stream));
		return buf;
	};
	const int dev = (int)means3D.get_device();
	const int64_t guess = capacity_guess(dev, P, W, H);
	Tensor binning;
	if (guess >= 0) binning = second_half(-1, guess);
	counted.synchronize();
	const int64_t R = *R_host;
	if (guess < 0 || R > gsr_binning_capacity(P, W, H, (size_t)binning.numel())) binning = second_half(R, R);
	capacity_update(dev, P, W, H, R);
	if (debug) AT_CUDA_CHECK(cudaDeviceSynchronize()); // CHECK_CUDA(debug), CR/auxiliary.h:166-173
	return std::make_tuple(R, out_color, out_depth, out_unc, radii, geom, binning, img);
}

// ---- RasterizeGaussiansBackwardCUDA, rasterize_points.cu:124-211 -----------------------------------------------------
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> rasterize_gaussians_backward(
    const Tensor &background, const Tensor &means3D, const Tensor &radii, const Tensor &colors, const Tensor &scales,
    const Tensor &rotations, const float scale_modifier, const Tensor &cov3D_precomp, const Tensor &viewmatrix, const Tensor &projmatrix,
    const float tan_fovx, const float tan_fovy, const Tensor &dL_dout_color, const Tensor &dL_dout_depth,
    const Tensor &dL_dout_uncertainty, const Tensor &sh, const int degree, const Tensor &campos, const Tensor &geomBuffer,
    const int64_t R, const Tensor &binningBuffer, const Tensor &imageBuffer, const bool debug)
{
	const int P = (int)means3D.size(0);
	const int C = (int)dL_dout_color.size(0), H = (int)dL_dout_color.size(1), W = (int)dL_dout_color.size(2); // rasterize_points.cu:151-152
	const int M = sh.numel() != 0 ? (int)sh.size(1) : 0;
	const bool has_cov = cov3D_precomp.numel() != 0;
	const auto f32 = means3D.options().dtype(torch::kFloat32);
	Tensor dL_dmeans2D = torch::empty({P, 3}, f32), dL_dcolors = torch::empty({P, C}, f32), dL_dopacity = torch::empty({P, 1}, f32),
	       dL_duncertainty = torch::empty({P, 1}, f32), dL_dmeans3D = torch::empty({P, 3}, f32), dL_dcov3D = torch::empty({P, 6}, f32),
	       dL_dsh = torch::zeros({P, M, 3}, f32);
	Tensor dL_dscales = has_cov ? torch::zeros({P, 3}, f32) : torch::empty({P, 3}, f32);
	Tensor dL_drotations = has_cov ? torch::zeros({P, 4}, f32) : torch::empty({P, 4}, f32);
	if (P != 0) {
		const c10::cuda::CUDAGuard guard(means3D.device());
		const F32 m3(means3D, "means3D"), col(colors, "colors"), sc(scales, "scales"), ro(rotations, "rotations"), cov(cov3D_precomp, "cov3D_precomp"),
		    view(viewmatrix, "viewmatrix"), proj(projmatrix, "projmatrix"), cam(campos, "campos"), bg(background, "bg"), shs(sh, "sh"),
		    gc(dL_dout_color, "dL_dout_color"), gd(dL_dout_depth, "dL_dout_depth"), gu(dL_dout_uncertainty, "dL_dout_uncertainty");
		check_rc(gsr_backward(P, C, degree, M, R, bg.ptr(), W, H, m3.ptr(), shs.ptr(), col.ptr(), sc.ptr(), scale_modifier, ro.ptr(), cov.ptr(),
		                      view.ptr(), proj.ptr(), cam.ptr(), tan_fovx, tan_fovy, radii.data_ptr<int>(), geomBuffer.data_ptr(),
		                      (size_t)geomBuffer.numel(), binningBuffer.data_ptr(), (size_t)binningBuffer.numel(), imageBuffer.data_ptr(),
		                      (size_t)imageBuffer.numel(), gc.ptr(), gd.ptr(), gu.ptr(), dL_dmeans2D.data_ptr<float>(),
		                      dL_dcolors.data_ptr<float>(), dL_dopacity.data_ptr<float>(), dL_duncertainty.data_ptr<float>(),
		                      dL_dmeans3D.data_ptr<float>(), dL_dcov3D.data_ptr<float>(), M ? dL_dsh.data_ptr<float>() : nullptr,
		                      has_cov ? nullptr : dL_dscales.data_ptr<float>(), has_cov ? nullptr : dL_drotations.data_ptr<float>(),
		                      /*accumulate=*/0, current_stream()));
		if (debug) AT_CUDA_CHECK(cudaDeviceSynchronize());
	}
	return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_duncertainty, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations);
}

// ---- RasterizeGaussiansfilterCUDA (sic), rasterize_points.cu:235-299 -------------------------------------------------
Tensor rasterize_aussians_filter(const Tensor &means3D, const Tensor &scales, const Tensor &rotations, const float scale_modifier,
                                 const Tensor &cov3D_precomp, const Tensor &viewmatrix, const Tensor &projmatrix, const float tan_fovx,
                                 const float tan_fovy, const int image_height, const int image_width, const bool prefiltered, const bool debug)
{
	check_means(means3D);
	const int P = (int)means3D.size(0);
	Tensor radii = torch::zeros({P}, means3D.options().dtype(torch::kInt32));
	if (P != 0) {
		const c10::cuda::CUDAGuard guard(means3D.device());
		const F32 m3(means3D, "means3D"), sc(scales, "scales"), ro(rotations, "rotations"), cov(cov3D_precomp, "cov3D_precomp"),
		    view(viewmatrix, "viewmatrix"), proj(projmatrix, "projmatrix");
		check_rc(gsr_visible_filter(P, m3.ptr(), sc.ptr(), 3, scale_modifier, ro.ptr(), cov.ptr(), view.ptr(), proj.ptr(), image_width, image_height,
		                            tan_fovx, tan_fovy, prefiltered ? 1 : 0, radii.data_ptr<int>(), current_stream()));
		if (debug) AT_CUDA_CHECK(cudaDeviceSynchronize());
	}
	return radii;
}

// ---- RasterizeGaussiansfilterPositionCUDA, rasterize_points.cu:304-373 -----------------------------------------------
std::tuple<Tensor, Tensor, Tensor> rasterize_aussians_filter_position2D(const Tensor &means3D, const Tensor &scales, const Tensor &rotations,
                                                                        const float scale_modifier, const Tensor &cov3D_precomp,
                                                                        const Tensor &viewmatrix, const Tensor &projmatrix, const float tan_fovx,
                                                                        const float tan_fovy, const int image_height, const int image_width,
                                                                        const bool prefiltered, const bool debug)
{
	check_means(means3D);
	const int P = (int)means3D.size(0);
	const auto f32 = means3D.options().dtype(torch::kFloat32);
	Tensor radii = torch::zeros({P}, means3D.options().dtype(torch::kInt32)), x = torch::zeros({P}, f32), y = torch::zeros({P}, f32);
	if (P != 0) {
		const c10::cuda::CUDAGuard guard(means3D.device());
		const F32 m3(means3D, "means3D"), sc(scales, "scales"), ro(rotations, "rotations"), cov(cov3D_precomp, "cov3D_precomp"),
		    view(viewmatrix, "viewmatrix"), proj(projmatrix, "projmatrix");
		check_rc(gsr_position2d_filter(P, m3.ptr(), sc.ptr(), 3, scale_modifier, ro.ptr(), cov.ptr(), view.ptr(), proj.ptr(), image_width,
		                               image_height, tan_fovx, tan_fovy, prefiltered ? 1 : 0, radii.data_ptr<int>(), x.data_ptr<float>(),
		                               y.data_ptr<float>(), current_stream()));
		if (debug) AT_CUDA_CHECK(cudaDeviceSynchronize());
	}
	return std::make_tuple(radii, x, y);
}

// ---- markVisible, rasterize_points.cu:213-232 -----------------------------------------------------------------------
Tensor mark_visible(const Tensor &means3D, const Tensor &viewmatrix, const Tensor &projmatrix)
{
	const int P = (int)means3D.size(0);
	Tensor present = torch::zeros({P}, means3D.options().dtype(torch::kBool));
	if (P != 0) {
		const c10::cuda::CUDAGuard guard(means3D.device());
		const F32 m3(means3D, "means3D"), view(viewmatrix, "viewmatrix"), proj(projmatrix, "projmatrix");
		check_rc(gsr_mark_visible(P, m3.ptr(), view.ptr(), proj.ptr(), reinterpret_cast<uint8_t *>(present.data_ptr<bool>()), current_stream()));
	}
	return present;
}

} // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
	m.doc() = "compiled torch glue of libgsr_b200.so: the five functions of the reference's diff_gaussian_rasterization._C (ext.cpp:16-20)";
	m.def("rasterize_gaussians", &rasterize_gaussians);
	m.def("rasterize_gaussians_backward", &rasterize_gaussians_backward);
	m.def("rasterize_aussians_filter", &rasterize_aussians_filter);
	m.def("rasterize_aussians_filter_position2D", &rasterize_aussians_filter_position2D);
	m.def("mark_visible", &mark_visible);
}
