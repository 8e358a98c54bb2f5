"""gscream_b200 — B200-native (sm_100a) differentiable Gaussian rasterizer behind the
`diff_gaussian_rasterization` surface of W-Ted/GScream.  See DESIGN.md."""
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
