"""Python stand-in for the reference's pybind11 module `diff_gaussian_rasterization._C`.

Same five functions, same positional arguments, same return tuples — including the
`rasterize_aussians_*` spellings — as submodules/diff-gaussian-rasterization/ext.cpp:16-20 and
rasterize_points.cu:35-373 of W-Ted/GScream, but every byte of compute goes through the C ABI of
libgsr_b200.so (include/gsr_b200.h).  Torch is used only to own device memory and to name the
current stream, which is what rasterize_points.cu does with libtorch.
"""
import itertools
import os
import threading

import torch

from . import _lib

_PINNED = {}
_PINNED_LOCK = threading.Lock()
_PINNED_SLOTS = 64
_R_HINT = {}


def _compiled():
    """`GSR_GLUE=cpp`: route the five reference entry points through the compiled pybind11 glue (csrc/gsr_torch_glue.cpp, the
    C++ counterpart of the reference's rasterize_points.cu above the same C ABI) instead of this module's ctypes marshalling.
    Read at call time so that tests can compare the two glues in one process."""
    if os.environ.get("GSR_GLUE", "ctypes") != "cpp":
        return None
    from . import _glue
    return _glue.load()


def _ptr(t):
    """Device pointer of a tensor; empty tensors become NULL (the reference's convention for
    'not provided', diff_gaussian_rasterization/__init__.py:230-240)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t, name):
    if t is None or t.numel() == 0:
        return t
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor" % name)
    return t.contiguous()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _pinned_i64(device):
    """A pinned int64 for one call's num_rendered.  Slots rotate through a small per-device ring, so that concurrent calls
    (threads, streams) on one device never share a counter: a call waits for its own value before it returns, which frees
    its slot long before the ring wraps."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    with _PINNED_LOCK:
        if key not in _PINNED:
            _PINNED[key] = (torch.zeros(_PINNED_SLOTS, dtype=torch.int64).pin_memory(), itertools.count())
        ring, counter = _PINNED[key]
        return ring[next(counter) % _PINNED_SLOTS:][:1]


def _capacity_guess(key):
    """Instance capacity to size the binning buffer with BEFORE this call's num_rendered has reached the host: 25 % above the
    (slowly decaying) running maximum of what this problem shape produced so far.  None: no history, take the exact path."""
    hint = _R_HINT.get(key)
    return None if hint is None else int(hint * 1.25) + 65536


def _capacity_update(key, R):
    _R_HINT[key] = max(float(R), 0.98 * _R_HINT.get(key, 0.0))


def _check_means(means3D):
    if means3D.ndimension() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:58-60


def rasterize_gaussians(background, means3D, colors, opacity, uncertaintys, scales, rotations, scale_modifier,
                        cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width,
                        sh, degree, campos, prefiltered, debug):
    """RasterizeGaussiansCUDA, rasterize_points.cu:35-122.  Returns
    (num_rendered, color[C,H,W], depth[1,H,W], uncertainty[1,H,W], radii[P], geomBuffer, binningBuffer, imgBuffer)."""
    native = _compiled()
    if native is not None:
        return native.rasterize_gaussians(background, means3D, colors, opacity, uncertaintys, scales, rotations, float(scale_modifier),
                                          cov3D_precomp, viewmatrix, projmatrix, float(tan_fovx), float(tan_fovy), int(image_height),
                                          int(image_width), sh, int(degree), campos, bool(prefiltered), bool(debug))
    lib = _lib.load()
    _check_means(means3D)
    if not means3D.is_cuda:
        raise ValueError("means3D must be a CUDA tensor (there is no CPU rasterizer)")
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    dev = means3D.device
    with torch.cuda.device(dev):
        means3D = _f32c(means3D, "means3D")
        colors, opacity, uncertaintys = _f32c(colors, "colors"), _f32c(opacity, "opacity"), _f32c(uncertaintys, "uncertainties")
        scales, rotations, cov3D_precomp = _f32c(scales, "scales"), _f32c(rotations, "rotations"), _f32c(cov3D_precomp, "cov3D_precomp")
        viewmatrix, projmatrix, campos = _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix"), _f32c(campos, "campos")
        background, sh = _f32c(background, "bg"), _f32c(sh, "sh")
        has_colors = colors is not None and colors.numel() != 0
        C = colors.size(1) if has_colors else 3
        M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
        if not has_colors and C != 3:
            raise RuntimeError("For non-RGB, provide precomputed Gaussian colors!")

        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        if P == 0:  # rasterize_points.cu:85: nothing is launched, the images are the reference's torch::full(0)
            empty = torch.empty((0,), **u8)
            return (0, torch.zeros((C, H, W), **f32), torch.zeros((1, H, W), **f32), torch.zeros((1, H, W), **f32),
                    torch.zeros((0,), dtype=torch.int32, device=dev), empty, empty.clone(), empty.clone())
        # every element of the four outputs is written by the kernels (the reference fills them with zeros first,
        # rasterize_points.cu:69-72)
        out_color = torch.empty((C, H, W), **f32)
        out_depth = torch.empty((1, H, W), **f32)
        out_unc = torch.empty((1, H, W), **f32)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)

        geom = torch.empty((lib.gsr_geom_bytes(P),), **u8)
        img = torch.empty((lib.gsr_image_bytes(W, H),), **u8)
        pinned = _pinned_i64(dev)
        stream = _stream()
        _lib.check(lib.gsr_forward_stage1(
            P, C, int(degree), M, _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity), _ptr(uncertaintys),
            _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix),
            _ptr(campos), W, H, float(tan_fovx), float(tan_fovy), int(bool(prefiltered)), radii.data_ptr(),
            geom.data_ptr(), geom.numel(), pinned.data_ptr(), stream))
        # The reference API returns num_rendered as a Python int and sizes the binning buffer from it: one blocking copy in the
        # middle of the forward (CR/rasterizer_impl.cu:287) that drains the GPU's queue.  Here the second half is launched right
        # behind the first with a buffer sized from an estimate (its kernels read num_rendered from device memory), and the host
        # then waits for stage 1's counter only — the GPU keeps running.  A wrong estimate costs one repeat of the second half.
        counted = torch.cuda.Event()
        counted.record()

        def second_half(num_rendered, capacity):
            buf = torch.empty((lib.gsr_binning_bytes(P, capacity, W, H),), **u8)
            _lib.check(lib.gsr_forward_stage2(
                P, C, num_rendered, _ptr(colors), _ptr(background), W, H, geom.data_ptr(), geom.numel(), buf.data_ptr(), buf.numel(),
                img.data_ptr(), img.numel(), out_color.data_ptr(), out_depth.data_ptr(), out_unc.data_ptr(), stream))
            return buf

        # history per device, image size and magnitude of P: in training P changes a little every iteration (the kept offsets),
        # num_rendered follows it smoothly, and the table must not grow by one entry per distinct P
        key = (dev.index, P.bit_length(), W, H)
        guess = None if os.environ.get("GSR_EXACT_BINNING") else _capacity_guess(key)
        binning = second_half(-1, guess) if guess is not None else None
        counted.synchronize()
        R = int(pinned.item())
        if binning is None or R > lib.gsr_binning_capacity(P, W, H, binning.numel()):
            binning = second_half(R, R)
        _capacity_update(key, R)
        if debug:  # CHECK_CUDA(debug), auxiliary.h:166-173
            torch.cuda.synchronize(dev)
    return R, out_color, out_depth, out_unc, radii, geom, binning, img


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp,
                                 viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_depth,
                                 dL_dout_uncertainty, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug,
                                 out=None, accumulate=False, want_cov3D=True):
    """RasterizeGaussiansBackwardCUDA, rasterize_points.cu:124-211.  Returns
    (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_duncertainty, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations).
    The reference signature is the first 23 positional arguments.  Extensions: `out`/`accumulate` let
    gscream_b200.dist accumulate several views into one flat gradient bucket; `want_cov3D=False` skips
    writing dL_dcov3D when scales/rotations were given (the reference always materialises it even though
    nothing consumes it in that case) and returns None in its place."""
    native = _compiled() if (out is None and not accumulate) else None
    if native is not None:   # the reference's 23 positional arguments; dL_dcov3D is always materialised, like upstream
        return native.rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, float(scale_modifier), cov3D_precomp,
                                                   viewmatrix, projmatrix, float(tan_fovx), float(tan_fovy), dL_dout_color, dL_dout_depth,
                                                   dL_dout_uncertainty, sh, int(degree), campos, geomBuffer, int(R), binningBuffer,
                                                   imageBuffer, bool(debug))
    lib = _lib.load()
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    C = dL_dout_color.size(0)
    dev = means3D.device
    with torch.cuda.device(dev):
        means3D = _f32c(means3D, "means3D")
        colors, scales, rotations = _f32c(colors, "colors"), _f32c(scales, "scales"), _f32c(rotations, "rotations")
        cov3D_precomp, sh = _f32c(cov3D_precomp, "cov3D_precomp"), _f32c(sh, "sh")
        viewmatrix, projmatrix, campos = _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix"), _f32c(campos, "campos")
        background = _f32c(background, "bg")
        g_color, g_depth, g_unc = _f32c(dL_dout_color, "dL_dout_color"), _f32c(dL_dout_depth, "dL_dout_depth"), _f32c(dL_dout_uncertainty, "dL_dout_uncertainty")
        M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
        has_cov = cov3D_precomp is not None and cov3D_precomp.numel() != 0

        opts = dict(dtype=torch.float32, device=dev)
        if out is None:
            # the library overwrites every element (zeros for culled Gaussians), so no zero-fill pass
            # is needed for the per-Gaussian outputs; dL_dcolors is zeroed inside gsr_backward.
            out = dict(
                dL_dmeans2D=torch.empty((P, 3), **opts), dL_dcolors=torch.empty((P, C), **opts),
                dL_dopacity=torch.empty((P, 1), **opts), dL_duncertainty=torch.empty((P, 1), **opts),
                dL_dmeans3D=torch.empty((P, 3), **opts),
                dL_dcov3D=torch.empty((P, 6), **opts) if (has_cov or want_cov3D) else None,
                dL_dsh=torch.zeros((P, M, 3), **opts),
                dL_dscales=torch.zeros((P, 3), **opts) if has_cov else torch.empty((P, 3), **opts),
                dL_drotations=torch.zeros((P, 4), **opts) if has_cov else torch.empty((P, 4), **opts))
        if P != 0:
            _lib.check(lib.gsr_backward(
                P, C, int(degree), M, int(R), _ptr(background), W, H, _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(scales),
                float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                float(tan_fovx), float(tan_fovy), radii.data_ptr(), geomBuffer.data_ptr(), geomBuffer.numel(),
                binningBuffer.data_ptr(), binningBuffer.numel(), imageBuffer.data_ptr(), imageBuffer.numel(),
                g_color.data_ptr(), g_depth.data_ptr(), g_unc.data_ptr(),
                out["dL_dmeans2D"].data_ptr(), out["dL_dcolors"].data_ptr(), out["dL_dopacity"].data_ptr(),
                out["dL_duncertainty"].data_ptr(), out["dL_dmeans3D"].data_ptr(),
                out["dL_dcov3D"].data_ptr() if out.get("dL_dcov3D") is not None else None,
                _ptr(out.get("dL_dsh")),
                None if has_cov else out["dL_dscales"].data_ptr(),
                None if has_cov else out["dL_drotations"].data_ptr(),
                int(bool(accumulate)), _stream()))
            if debug:
                torch.cuda.synchronize(dev)
    return (out["dL_dmeans2D"], out["dL_dcolors"], out["dL_dopacity"], out["dL_duncertainty"], out["dL_dmeans3D"],
            out.get("dL_dcov3D"), out.get("dL_dsh"), out["dL_dscales"], out["dL_drotations"])


def _rows3(t, name):
    """(tensor, row stride in floats) of a [P,3] float32 CUDA tensor.  A row-strided view with unit inner stride — GScream hands
    the filters `get_scaling[:, :3]` of a [A,6] tensor (gaussian_renderer/__init__.py:298) — is read in place by the kernel
    instead of through the contiguous copy the reference makes (rasterize_points.cu:280)."""
    if t is None or t.numel() == 0:
        return t, 3
    if (t.dim() == 2 and t.size(1) == 3 and t.dtype == torch.float32 and t.is_cuda and t.stride(1) == 1
            and t.stride(0) >= 3 and t.data_ptr() % 4 == 0):
        return t, int(t.stride(0))
    return _f32c(t, name), 3


def _filter_common(means3D, scales, rotations, cov3D_precomp, viewmatrix, projmatrix):
    _check_means(means3D)
    scales, stride = _rows3(scales, "scales")
    return (_f32c(means3D, "means3D"), scales, stride, _f32c(rotations, "rotations"),
            _f32c(cov3D_precomp, "cov3D_precomp"), _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix"))


def rasterize_aussians_filter(means3D, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                              tan_fovx, tan_fovy, image_height, image_width, prefiltered, debug):
    """RasterizeGaussiansfilterCUDA (sic), rasterize_points.cu:235-299 -> radii[P] int32."""
    native = _compiled()
    if native is not None:
        return native.rasterize_aussians_filter(means3D, scales, rotations, float(scale_modifier), cov3D_precomp, viewmatrix, projmatrix,
                                                float(tan_fovx), float(tan_fovy), int(image_height), int(image_width), bool(prefiltered), bool(debug))
    lib = _lib.load()
    _check_means(means3D)
    dev = means3D.device
    with torch.cuda.device(dev):
        means3D, scales, scales_stride, rotations, cov3D_precomp, viewmatrix, projmatrix = _filter_common(
            means3D, scales, rotations, cov3D_precomp, viewmatrix, projmatrix)
        P = means3D.size(0)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)  # every element is written
        if P != 0:
            _lib.check(lib.gsr_visible_filter(
                P, _ptr(means3D), _ptr(scales), scales_stride, float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix),
                _ptr(projmatrix), int(image_width), int(image_height), float(tan_fovx), float(tan_fovy), int(bool(prefiltered)),
                radii.data_ptr(), _stream()))
            if debug:
                torch.cuda.synchronize(dev)
    return radii


def rasterize_aussians_filter_position2D(means3D, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                                         tan_fovx, tan_fovy, image_height, image_width, prefiltered, debug):
    """RasterizeGaussiansfilterPositionCUDA, rasterize_points.cu:304-373 -> (radii, x, y)."""
    native = _compiled()
    if native is not None:
        return native.rasterize_aussians_filter_position2D(means3D, scales, rotations, float(scale_modifier), cov3D_precomp, viewmatrix,
                                                           projmatrix, float(tan_fovx), float(tan_fovy), int(image_height), int(image_width),
                                                           bool(prefiltered), bool(debug))
    lib = _lib.load()
    _check_means(means3D)
    dev = means3D.device
    with torch.cuda.device(dev):
        means3D, scales, scales_stride, rotations, cov3D_precomp, viewmatrix, projmatrix = _filter_common(
            means3D, scales, rotations, cov3D_precomp, viewmatrix, projmatrix)
        P = means3D.size(0)
        radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        x = torch.zeros((P,), dtype=torch.float32, device=dev)
        y = torch.zeros((P,), dtype=torch.float32, device=dev)
        if P != 0:
            _lib.check(lib.gsr_position2d_filter(
                P, _ptr(means3D), _ptr(scales), scales_stride, float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix),
                _ptr(projmatrix), int(image_width), int(image_height), float(tan_fovx), float(tan_fovy), int(bool(prefiltered)),
                radii.data_ptr(), x.data_ptr(), y.data_ptr(), _stream()))
            if debug:
                torch.cuda.synchronize(dev)
    return radii, x, y


def mark_visible(means3D, viewmatrix, projmatrix):
    """markVisible, rasterize_points.cu:213-232 -> bool[P]."""
    native = _compiled()
    if native is not None:
        return native.mark_visible(means3D, viewmatrix, projmatrix)
    lib = _lib.load()
    dev = means3D.device
    with torch.cuda.device(dev):
        P = means3D.size(0)
        present = torch.zeros((P,), dtype=torch.bool, device=dev)
        if P != 0:
            means3D, viewmatrix, projmatrix = _f32c(means3D, "means3D"), _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix")
            _lib.check(lib.gsr_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix), present.data_ptr(), _stream()))
    return present


def debug_export(P, R, W, H, geomBuffer, binningBuffer, imageBuffer):
    """Parity-test helper: copies of the intermediates hidden in the opaque scratch buffers."""
    lib = _lib.load()
    dev = geomBuffer.device
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    with torch.cuda.device(dev):
        o = dict(
            xy=torch.zeros((P, 2), dtype=torch.float32, device=dev), depths=torch.zeros((P,), dtype=torch.float32, device=dev),
            conic_opacity=torch.zeros((P, 4), dtype=torch.float32, device=dev),
            tiles_touched=torch.zeros((P,), dtype=torch.int32, device=dev),
            point_list=torch.zeros((max(R, 1),), dtype=torch.int32, device=dev),
            ranges=torch.zeros((tiles, 2), dtype=torch.int32, device=dev),
            final_T=torch.zeros((H * W,), dtype=torch.float32, device=dev),
            n_contrib=torch.zeros((H * W,), dtype=torch.int32, device=dev))
        _lib.check(lib.gsr_debug_export(
            P, R, W, H, geomBuffer.data_ptr(), binningBuffer.data_ptr(), binningBuffer.numel(), imageBuffer.data_ptr(), o["xy"].data_ptr(),
            o["depths"].data_ptr(), o["conic_opacity"].data_ptr(), o["tiles_touched"].data_ptr(), o["point_list"].data_ptr(),
            o["ranges"].data_ptr(), o["final_T"].data_ptr(), o["n_contrib"].data_ptr(), _stream()))
        o["point_list"] = o["point_list"][:R]
    return o
