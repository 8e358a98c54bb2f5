"""Host-side mirror of the reference's `diff_gaussian_rasterization` Python package.

Same names, argument order, return tuples and error behaviour as
submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py of W-Ted/GScream
(GaussianRasterizationSettings :189-201, GaussianRasterizer :203-312, _RasterizeGaussians :46-187,
rasterize_gaussians :21-44), so gaussian_renderer/__init__.py:15,131-158,265-300 runs unmodified.
The native side is libgsr_b200.so through gscream_b200._C (a stand-in for the pybind11 `_C`).
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _C


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def _call_native(fn, args, debug, dump_name, message, **kwargs):
    """Debug mode of the reference binding (diff_gaussian_rasterization/__init__.py:87-95, 150-158): CPU-clone the positional
    arguments before the call can corrupt them, and on any exception save them as a snapshot, print the hint and re-raise."""
    if not debug:
        return fn(*args, **kwargs)
    cpu_args = cpu_deep_copy_tuple(args)
    try:
        return fn(*args, **kwargs)
    except Exception as ex:
        torch.save(cpu_args, dump_name)
        print(message)
        raise ex


def _or_empty(t):
    """`None` inputs travel as empty CPU tensors -> null pointers in the native call (reference :230-240)."""
    return torch.Tensor([]) if t is None else t


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, uncertainties, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, uncertainties, scales,
                                     rotations, cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, uncertainties, scales, rotations,
                cov3Ds_precomp, raster_settings):
        # argument order of RasterizeGaussiansCUDA (rasterize_points.cu:35-56)
        args = (
            raster_settings.bg, means3D, colors_precomp, opacities, uncertainties, scales, rotations,
            raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix, raster_settings.projmatrix,
            raster_settings.tanfovx, raster_settings.tanfovy, raster_settings.image_height, raster_settings.image_width,
            sh, raster_settings.sh_degree, raster_settings.campos, raster_settings.prefiltered, raster_settings.debug,
        )
        num_rendered, color, depth, uncertainty, radii, geomBuffer, binningBuffer, imgBuffer = _call_native(
            _C.rasterize_gaussians, args, raster_settings.debug, "snapshot_fw.dump",
            "\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")

        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, depth, uncertainty, radii

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_depth, grad_out_uncertainty, _):
        num_rendered = ctx.num_rendered
        raster_settings = ctx.raster_settings
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer = ctx.saved_tensors

        # argument order of RasterizeGaussiansBackwardCUDA (rasterize_points.cu:124-148)
        args = (raster_settings.bg, means3D, radii, colors_precomp, scales, rotations, raster_settings.scale_modifier,
                cov3Ds_precomp, raster_settings.viewmatrix, raster_settings.projmatrix, raster_settings.tanfovx,
                raster_settings.tanfovy, grad_out_color, grad_out_depth, grad_out_uncertainty, sh,
                raster_settings.sh_degree, raster_settings.campos, geomBuffer, num_rendered, binningBuffer, imgBuffer,
                raster_settings.debug)
        want_cov = cov3Ds_precomp.numel() != 0
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_uncertainties, grad_means3D, grad_cov3Ds_precomp,
         grad_sh, grad_scales, grad_rotations) = _call_native(
            _C.rasterize_gaussians_backward, args, raster_settings.debug, "snapshot_bw.dump",
            "\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n", want_cov3D=want_cov)

        # an input that was passed as an empty placeholder gets no gradient
        def _g(grad, inp):
            return grad if inp.numel() != 0 else None

        grads = (
            grad_means3D,
            grad_means2D,
            _g(grad_sh, sh),
            _g(grad_colors_precomp, colors_precomp),
            grad_opacities,
            grad_uncertainties,
            _g(grad_scales, scales),
            _g(grad_rotations, rotations),
            _g(grad_cov3Ds_precomp, cov3Ds_precomp) if grad_cov3Ds_precomp is not None else None,
            None,
        )
        return grads


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, uncertainties, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        return rasterize_gaussians(means3D, means2D, _or_empty(shs), _or_empty(colors_precomp), opacities, uncertainties,
                                   _or_empty(scales), _or_empty(rotations), _or_empty(cov3D_precomp), raster_settings)

    def visible_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
        with torch.no_grad():
            return _C.rasterize_aussians_filter(*self._filter_args(means3D, scales, rotations, cov3D_precomp))

    def position2D_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
        with torch.no_grad():
            radii, position2D_x, position2D_y = _C.rasterize_aussians_filter_position2D(
                *self._filter_args(means3D, scales, rotations, cov3D_precomp))
        return radii, position2D_x, position2D_y

    def _filter_args(self, means3D, scales, rotations, cov3D_precomp):
        """The 13 positional arguments shared by the two anchor filters (rasterize_points.h:74-104)."""
        rs = self.raster_settings
        return (means3D, _or_empty(scales), _or_empty(rotations), rs.scale_modifier, _or_empty(cov3D_precomp), rs.viewmatrix,
                rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, rs.prefiltered, rs.debug)
