"""View-parallel execution of the rasterizer hot path over the GPUs of one node.

The reference (W-Ted/GScream) is single-GPU (utils/general_utils.py:208).  The path shards naturally by
camera view: Gaussian parameters are replicated, view v of a batch runs on rank
v % world_size, every rank accumulates the per-Gaussian gradients of its views into ONE flat fp32 bucket
(the backward kernels add straight into it, `accumulate=1` in the C ABI — no per-view temporaries), and a
single sum-allreduce of that bucket per step (NCCL over NVLink/NVSwitch; gloo in the CPU tests) makes the
gradients identical on every rank.  There is no other data-path collective.

Bucket order: the colour block first (two thirds of the bucket at C = 32; final as soon as the blend backward is done), then
the other parameter gradients, then — outside the all-reduced range — the screen-space means2D gradients: they are the
per-view signal of the densification statistics (scene/gaussian_model.py:755), not a parameter gradient, and SURVEY 8e's
bucket does not carry them (44 floats per Gaussian at C = 32).  (Starting the collective on the colour block on a side stream
while the per-Gaussian backward kernel still runs was built and measured: two collectives cost more than the 0.08 ms they hide,
profiles/r2_scaling.md.)
"""
from typing import Dict, List, Sequence

import torch

# (name, columns) of the per-Gaussian gradient blocks in bucket order; C is filled in at construction.
_BLOCKS = (("colors", None), ("means3D", 3), ("opacities", 1), ("uncertainties", 1), ("scales", 3), ("rotations", 4),
           ("means2D", 3))
_LOCAL_ONLY = ("means2D",)   # kept per rank, not all-reduced


def shard_views(n_views: int, world_size: int, rank: int) -> List[int]:
    """Round-robin assignment of views to ranks (view v -> rank v % world_size)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return [v for v in range(n_views) if v % world_size == rank]


class GradBucket:
    """One flat fp32 buffer holding every gradient the rasterizer returns, as contiguous [P,k] blocks."""

    def __init__(self, P: int, C: int, device="cpu"):
        self.P, self.C = int(P), int(C)
        self.layout = []
        off = 0
        for name, cols in _BLOCKS:
            cols = self.C if cols is None else cols
            self.layout.append((name, off, cols))
            off += self.P * cols
        self.numel = off
        self.reduced_numel = sum(self.P * cols for name, _, cols in self.layout if name not in _LOCAL_ONLY)  # a prefix of the buffer
        self.colors_numel = self.P * self.C                                                                  # ... and its first block
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.views: Dict[str, torch.Tensor] = {
            name: self.flat[o:o + self.P * cols].view(self.P, cols) for name, o, cols in self.layout}

    def zero_(self):
        self.flat.zero_()
        return self

    def nbytes(self) -> int:
        return self.numel * 4

    def reduced_nbytes(self) -> int:
        """Bytes the per-step collective moves (the parameter gradients; the per-view means2D block stays local)."""
        return self.reduced_numel * 4

    def as_backward_out(self) -> Dict[str, torch.Tensor]:
        """Mapping expected by gscream_b200._C.rasterize_gaussians_backward(out=..., accumulate=True)."""
        v = self.views
        return dict(dL_dmeans2D=v["means2D"], dL_dcolors=v["colors"], dL_dopacity=v["opacities"],
                    dL_duncertainty=v["uncertainties"], dL_dmeans3D=v["means3D"], dL_dcov3D=None, dL_dsh=None,
                    dL_dscales=v["scales"], dL_drotations=v["rotations"])


def allreduce_bucket(bucket: GradBucket, group=None, average: bool = False):
    """The one collective of the path: sum (or mean) of the flat gradient bucket over all ranks."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return bucket
    red = bucket.flat[:bucket.reduced_numel]
    dist.all_reduce(red, op=dist.ReduceOp.SUM, group=group)
    if average:
        red.div_(dist.get_world_size(group))
    return bucket


def render_views_into_bucket(scene: Dict[str, torch.Tensor], cameras: Sequence[dict], upstream: Sequence[tuple],
                             bucket: GradBucket, keep_outputs: bool = False, overwrite: bool = False):
    """Forward + backward of every local view, gradients accumulated into `bucket` (not zeroed here).
    overwrite=True: the first view's backward OVERWRITES the bucket (the kernels write every element, zeros for culled
    Gaussians) and only the later views add to it, so the caller need not zero 4 * (15 + C) * P bytes per step.

    scene: device tensors means3D/colors/opacities/uncertainties/scales/rotations/bg;
    cameras: dicts as produced by scenes.make_camera with device matrices;
    upstream: per view (dL_dcolor[C,H,W], dL_ddepth[1,H,W], dL_dunc[1,H,W]).
    Goes through the same `_C` entry points as GaussianRasterizer, minus autograd bookkeeping.
    """
    from . import _C
    outs = []
    empty = torch.empty(0)
    out_map = bucket.as_backward_out()
    for i, (cam, (gc, gd, gu)) in enumerate(zip(cameras, upstream)):
        R, color, depth, unc, radii, geom, binning, img = _C.rasterize_gaussians(
            scene["bg"], scene["means3D"], scene["colors"], scene["opacities"], scene["uncertainties"], scene["scales"],
            scene["rotations"], 1.0, empty, cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"],
            cam["H"], cam["W"], empty, 1, cam["campos"], False, False)
        _C.rasterize_gaussians_backward(
            scene["bg"], scene["means3D"], radii, scene["colors"], scene["scales"], scene["rotations"], 1.0, empty,
            cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"], gc, gd, gu, empty, 1, cam["campos"],
            geom, R, binning, img, False, out=out_map, accumulate=not (overwrite and i == 0), want_cov3D=False)
        if keep_outputs:
            outs.append((color, depth, unc, radii, R))
    return outs
