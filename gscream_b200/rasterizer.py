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
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # copy them before they can be corrupted
            try:
                num_rendered, color, depth, uncertainty, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, depth, uncertainty, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)

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
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                (grad_means2D, grad_colors_precomp, grad_opacities, grad_uncertainties, grad_means3D, grad_cov3Ds_precomp,
                 grad_sh, grad_scales, grad_rotations) = _C.rasterize_gaussians_backward(*args, want_cov3D=want_cov)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            (grad_means2D, grad_colors_precomp, grad_opacities, grad_uncertainties, grad_means3D, grad_cov3Ds_precomp,
             grad_sh, grad_scales, grad_rotations) = _C.rasterize_gaussians_backward(*args, want_cov3D=want_cov)

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

        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, uncertainties, scales, rotations,
                                   cov3D_precomp, raster_settings)

    def visible_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        with torch.no_grad():
            radii = _C.rasterize_aussians_filter(
                means3D, scales, rotations, raster_settings.scale_modifier, cov3D_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy, raster_settings.image_height,
                raster_settings.image_width, raster_settings.prefiltered, raster_settings.debug)
        return radii

    def position2D_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        with torch.no_grad():
            radii, position2D_x, position2D_y = _C.rasterize_aussians_filter_position2D(
                means3D, scales, rotations, raster_settings.scale_modifier, cov3D_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy, raster_settings.image_height,
                raster_settings.image_width, raster_settings.prefiltered, raster_settings.debug)
        return radii, position2D_x, position2D_y
