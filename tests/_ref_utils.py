"""TEST INFRASTRUCTURE — access to the reference rasterizer's own CUDA build (oracle/_ref).

oracle/_ref/dgr{3,32}/diff_gaussian_rasterization/ are installed-package layouts produced by
oracle/build_ref.py from the unmodified sources under /root/reference.  They are imported under
private module names so they can live next to this repo's drop-in `diff_gaussian_rasterization`.
Also parses the reference's opaque scratch buffers (layout: cuda_rasterizer/rasterizer_impl.cu:155-195,
rasterizer_impl.h:21-27) so integer artefacts can be compared bit for bit.
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
_MODS = {}


def ref_available(channels):
    d = os.path.join(REF_DIR, "dgr%d" % channels, "diff_gaussian_rasterization")
    return os.path.exists(os.path.join(d, "_C.so")) and os.path.exists(os.path.join(d, "__init__.py"))


def load_ref(channels):
    """Import the reference package built with NUM_CHANNELS == channels (3 or 32)."""
    if channels in _MODS:
        return _MODS[channels]
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    d = os.path.join(REF_DIR, "dgr%d" % channels, "diff_gaussian_rasterization")
    name = "ref_dgr%d" % channels
    spec = importlib.util.spec_from_file_location(name, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    _MODS[channels] = mod
    return mod


def _al(off, a=128):
    return (off + a - 1) // a * a


def parse_ref_geom(buf, P):
    """buf: uint8 numpy array of the reference geomBuffer."""
    o = 0
    out = {}

    def take(name, count, dtype):
        nonlocal o
        o = _al(o)
        n = count * np.dtype(dtype).itemsize
        out[name] = buf[o:o + n].view(dtype).copy()
        o += n

    take("depths", P, np.float32)
    take("clamped", 3 * P, np.uint8)
    take("internal_radii", P, np.int32)
    take("means2D", 2 * P, np.float32)
    take("cov3D", 6 * P, np.float32)
    take("conic_opacity", 4 * P, np.float32)
    take("uncertainty", P, np.float32)
    take("rgb", 3 * P, np.float32)
    take("tiles_touched", P, np.uint32)
    out["means2D"] = out["means2D"].reshape(P, 2)
    out["cov3D"] = out["cov3D"].reshape(P, 6)
    out["conic_opacity"] = out["conic_opacity"].reshape(P, 4)
    return out


def parse_ref_image(buf, W, H):
    N = W * H
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    o = 0
    out = {}
    for name, count, dtype in (("final_T", N, np.float32), ("n_contrib", N, np.uint32), ("ranges", 2 * N, np.uint32)):
        o = _al(o)
        n = count * np.dtype(dtype).itemsize
        out[name] = buf[o:o + n].view(dtype).copy()
        o += n
    out["ranges"] = out["ranges"].reshape(N, 2)[:tiles]
    return out


def parse_ref_binning(buf, R):
    o = 0
    out = {}
    for name, dtype in (("point_list", np.uint32), ("point_list_unsorted", np.uint32), ("keys_sorted", np.uint64), ("keys_unsorted", np.uint64)):
        o = _al(o)
        n = R * np.dtype(dtype).itemsize
        out[name] = buf[o:o + n].view(dtype).copy()
        o += n
    return out


def run_impl(mod, scene, cam, grads=None, device="cuda", C_module=None, debug=False):
    """Run forward (+ backward if grads given) through `mod.GaussianRasterizer` (reference or ours).
    Returns a dict of CPU numpy arrays including the raw scratch buffers."""
    import torch
    dev = torch.device(device)
    t = {k: v.to(dev) for k, v in scene.items()}
    leaves = {k: t[k].clone().requires_grad_(grads is not None) for k in ("means3D", "colors", "opacities", "uncertainties", "scales", "rotations")}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=grads is not None)
    settings = mod.GaussianRasterizationSettings(
        image_height=cam["H"], image_width=cam["W"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=t["bg"],
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=1,
        campos=cam["campos"].to(dev), prefiltered=False, debug=debug)
    rast = mod.GaussianRasterizer(raster_settings=settings)
    color, depth, unc, radii = rast(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
                                    uncertainties=leaves["uncertainties"], shs=None, colors_precomp=leaves["colors"],
                                    scales=leaves["scales"], rotations=leaves["rotations"], cov3D_precomp=None)
    out = dict(color=color.detach().cpu().numpy(), depth=depth.detach().cpu().numpy(), uncertainty=unc.detach().cpu().numpy(),
               radii=radii.cpu().numpy())
    fn = color.grad_fn
    if fn is not None:
        saved = fn.saved_tensors
        out["num_rendered"] = int(fn.num_rendered)
        out["_geom"], out["_binning"], out["_img"] = saved[7], saved[8], saved[9]
    if grads is not None:
        gc, gd, gu = (g.to(dev) for g in grads)
        torch.autograd.backward((color, depth, unc), (gc, gd, gu))
        out["dL_dmeans3D"] = leaves["means3D"].grad.cpu().numpy()
        out["dL_dmeans2D"] = means2D.grad.cpu().numpy()
        out["dL_dcolors"] = leaves["colors"].grad.cpu().numpy()
        out["dL_dopacity"] = leaves["opacities"].grad.cpu().numpy()
        out["dL_duncertainty"] = leaves["uncertainties"].grad.cpu().numpy()
        out["dL_dscales"] = leaves["scales"].grad.cpu().numpy()
        out["dL_drotations"] = leaves["rotations"].grad.cpu().numpy()
    return out
