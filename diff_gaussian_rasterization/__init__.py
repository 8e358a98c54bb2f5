"""Drop-in package: `from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`
(gaussian_renderer/__init__.py:15 of W-Ted/GScream) resolves here when this repository is on sys.path in place
of the reference's submodules/diff-gaussian-rasterization install.  Everything is served by gscream_b200."""
from gscream_b200 import _C  # noqa: F401  (same attribute name as the reference's pybind11 module)
from gscream_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _RasterizeGaussians,
    cpu_deep_copy_tuple,
    rasterize_gaussians,
)
