"""Drop-in for W-Ted/GScream's `gaussian_renderer` package (gaussian_renderer/__init__.py): the callers either side of the
rasterizer, on this repo's kernels.

    from gscream_b200.renderer import render, prefilter_voxel, prefilter_position2D, generate_neural_gaussians

Same names, signatures and return values as the reference module, so train.py:433 / :527 / :754 work unchanged:

  render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, visible_mask=None, retain_grad=False)
      gaussian_renderer/__init__.py:104-179 — fused anchor decode (gscream_b200.decode, one kernel each way instead of ~40 eager
      launches) -> GaussianRasterizer (libgsr_b200) -> the same result dict.
  prefilter_voxel(...) -> bool[A]                         gaussian_renderer/__init__.py:190-246
  prefilter_position2D(...) -> (bool[A], x[A], y[A])      gaussian_renderer/__init__.py:248-302
  prefilter_position2D_debug(...) -> (radii, x, y)        gaussian_renderer/__init__.py:306-359
      the anchor filters allocate no scratch (the reference allocates full geometry + image buffers there,
      CR/rasterizer_impl.cu:493-508) and never build the unused `screenspace_points` tensor of the reference's prefilters
      (:197-201, an autograd leaf nothing reads); the `pc.get_scaling[:, :3]` slice is read in place through its row stride
      (the reference copies it, rasterize_points.cu:280).

`viewpoint_camera` needs the attributes the reference reads: FoVx, FoVy, image_height, image_width, world_view_transform,
full_proj_transform, camera_center (scene/cameras.py:64-69); `pipe`: debug, compute_cov3D_python; `pc`: the GaussianModel
attributes of gscream_b200.decode.generate_neural_gaussians plus get_rotation (and get_covariance when compute_cov3D_python).
No CPU path: everything below raises if libgsr_b200.so is missing.
"""
import math

import torch

from .decode import generate_neural_gaussians
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer

__all__ = ["render", "prefilter_voxel", "prefilter_position2D", "prefilter_position2D_debug", "generate_neural_gaussians"]


def _settings(viewpoint_camera, pipe, bg_color, scaling_modifier):
    """The settings tuple every function of the reference module builds (e.g. gaussian_renderer/__init__.py:127-144)."""
    return GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5),
        tanfovy=math.tan(viewpoint_camera.FoVy * 0.5),
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=1,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=pipe.debug)


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, visible_mask=None, retain_grad=False):
    """gaussian_renderer/__init__.py:104-179.  Background tensor (bg_color) must be on the GPU."""
    is_training = pc.get_color_mlp.training
    if is_training:
        # visible_mask: anchors kept by the prefilter; mask: offsets whose neural opacity is > 0
        xyz, color, opacity, uncertainty, scaling, rot, neural_opacity, mask = generate_neural_gaussians(
            viewpoint_camera, pc, visible_mask, is_training=is_training)
    else:
        xyz, color, opacity, uncertainty, scaling, rot = generate_neural_gaussians(viewpoint_camera, pc, visible_mask, is_training=is_training)

    # the gradient sink for the 2-D means (train.py:599 reads its .grad); `+ 0` makes it a non-leaf exactly like the reference's
    screenspace_points = torch.zeros_like(xyz, dtype=pc.get_anchor.dtype, requires_grad=True, device=xyz.device) + 0
    if retain_grad:
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass

    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, pipe, bg_color, scaling_modifier))
    rendered_image, rendered_depth, uncer, radii = rasterizer(
        means3D=xyz, means2D=screenspace_points, shs=None, colors_precomp=color, opacities=opacity, uncertainties=uncertainty,
        scales=scaling, rotations=rot, cov3D_precomp=None)

    out = {"render": rendered_image, "render_depth": rendered_depth, "uncertainty": uncer, "viewspace_points": screenspace_points,
           "visibility_filter": radii > 0, "radii": radii}
    if is_training:
        out.update({"selection_mask": mask, "neural_opacity": neural_opacity, "scaling": scaling})
    return out


def _anchor_filter_inputs(pc, pipe, scaling_modifier):
    """scales / rotations / cov3D_precomp of the anchors as the reference's prefilters pick them (:229-237)."""
    if pipe.compute_cov3D_python:
        return None, None, pc.get_covariance(scaling_modifier)
    return pc.get_scaling[:, :3], pc.get_rotation, None


def prefilter_voxel(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None):
    """gaussian_renderer/__init__.py:190-246: which anchors project to a non-empty tile rectangle in this view."""
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, pipe, bg_color, scaling_modifier))
    scales, rotations, cov3D_precomp = _anchor_filter_inputs(pc, pipe, scaling_modifier)
    radii_pure = rasterizer.visible_filter(means3D=pc.get_anchor, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)
    return radii_pure > 0


def prefilter_position2D_debug(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None):
    """gaussian_renderer/__init__.py:306-359: (radii, x, y) of every anchor (x, y = projected pixel, 0 where culled)."""
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, pipe, bg_color, scaling_modifier))
    scales, rotations, cov3D_precomp = _anchor_filter_inputs(pc, pipe, scaling_modifier)
    return rasterizer.position2D_filter(means3D=pc.get_anchor, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)


def prefilter_position2D(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None):
    """gaussian_renderer/__init__.py:248-302."""
    radii_pure, x, y = prefilter_position2D_debug(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, override_color)
    return radii_pure > 0, x, y
