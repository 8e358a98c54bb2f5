"""gscream_b200.renderer — the drop-in for the reference's `gaussian_renderer` package (gaussian_renderer/__init__.py:104-359).

CPU: names, signatures (copied here from the reference file, lines cited) and loud failure without the CUDA library's inputs.
GPU: render() / prefilter_*() against the torch restatement of the reference glue (tests/_anchor_decode.py, itself pinned to the
reference's own function by tests/golden/decode_*.npz) driving the same rasterizer: same dict keys, same selection mask, images
within 1e-4 except threshold-flip pixels (the fused decode's MLPs differ from cuBLAS in the last bits), gradients to the anchor
parameters within 2e-3 of their scale.
"""
import inspect
import math
import types

import numpy as np
import pytest
import torch

import _anchor_decode as ad
from gscream_b200 import scenes


def _planes_close(got, ref, tol, flip, name):
    """Two decodes that differ in the last bits feed the same rasterizer: a (pixel, Gaussian) pair whose alpha sits on the 1/255
    threshold (CR/forward.cu:531) may blend in one run and not in the other, which moves that pixel by up to alpha * value.
    So: all but a 1e-4 fraction of the pixels within `tol`, and no pixel further than one threshold contribution (`flip`)."""
    err = np.abs(got - ref)
    assert float((err > tol).mean()) <= 1e-4, "%s: %d pixels beyond %.1e" % (name, int((err > tol).sum()), tol)
    assert float(err.max()) <= flip, "%s: max err %.3e" % (name, float(err.max()))

# gaussian_renderer/__init__.py:104, :190, :248, :306
SIGNATURES = {
    "render": "(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, visible_mask=None, retain_grad=False)",
    "prefilter_voxel": "(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None)",
    "prefilter_position2D": "(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None)",
    "prefilter_position2D_debug": "(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None)",
    "generate_neural_gaussians": "(viewpoint_camera, pc, visible_mask=None, is_training=False)",
}


def _camera(W, H, device="cpu", **kw):
    """An object with the attributes of scene/cameras.py:64-69 that the renderer reads."""
    cam = scenes.make_camera(W, H, **kw)
    return types.SimpleNamespace(FoVx=2.0 * math.atan(cam["tanfovx"]), FoVy=2.0 * math.atan(cam["tanfovy"]), image_height=H, image_width=W,
                                 world_view_transform=cam["viewmatrix"].to(device), full_proj_transform=cam["projmatrix"].to(device),
                                 camera_center=cam["campos"].to(device)), cam


PIPE = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)


def test_renderer_surface_matches_reference_signatures():
    from gscream_b200 import renderer
    for name, sig in SIGNATURES.items():
        assert str(inspect.signature(getattr(renderer, name))) == sig, name
    assert set(renderer.__all__) == set(SIGNATURES)


def test_renderer_has_no_cpu_path():
    from gscream_b200 import renderer
    pc = ad.SyntheticAnchors(50, seed=1)
    camera, _ = _camera(64, 48)
    with pytest.raises(Exception):
        renderer.prefilter_voxel(camera, pc, PIPE, torch.zeros(3))
    with pytest.raises(Exception):
        renderer.render(camera, pc, PIPE, torch.zeros(3))


@pytest.mark.gpu
def test_gpu_render_and_prefilters_match_reference_glue():
    from gscream_b200 import rasterizer as ours
    from gscream_b200 import renderer
    dev = torch.device("cuda")
    W, H = 640, 360
    camera, cam = _camera(W, H, dev, yaw_deg=4.0)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    pc = ad.SyntheticAnchors(8000, seed=12, tanfov=(cam["tanfovx"], cam["tanfovy"])).to(dev)
    pc.train()
    params = dict(anchor=pc._anchor, offset=pc._offset, feat=pc._anchor_feat, scaling=pc._scaling, w=pc.mlp_color[2].weight)

    # prefilters: the same kernels behind the same rasterizer methods -> identical
    vis, x, y = renderer.prefilter_position2D(camera, pc, PIPE, bg)
    vis_r, x_r, y_r = ad.prefilter_position2D(ours, cam, pc, bg)
    assert torch.equal(vis, vis_r) and torch.equal(x, x_r) and torch.equal(y, y_r)
    assert torch.equal(renderer.prefilter_voxel(camera, pc, PIPE, bg), vis_r)
    radii_dbg, _, _ = renderer.prefilter_position2D_debug(camera, pc, PIPE, bg)
    assert radii_dbg.dtype == torch.int32 and torch.equal(radii_dbg > 0, vis_r)
    assert 0 < int(vis.sum()) < vis.numel()

    target = torch.rand(3, H, W, generator=torch.Generator().manual_seed(2)).to(dev)
    res = {}
    for name in ("glue", "ours"):
        for t in params.values():
            t.grad = None
        pkg = renderer.render(camera, pc, PIPE, bg, visible_mask=vis, retain_grad=True) if name == "ours" else ad.render(ours, cam, pc, bg, vis)
        loss = (pkg["render"] - target).abs().mean() + 0.1 * pkg["render_depth"].mean() + 0.05 * pkg["uncertainty"].mean() + 0.01 * pkg["scaling"].prod(dim=1).mean()
        loss.backward()
        res[name] = dict(keys=set(pkg.keys()), image=pkg["render"].detach().cpu().numpy(), depth=pkg["render_depth"].detach().cpu().numpy(),
                         mask=pkg["selection_mask"].cpu().numpy(), radii=pkg["radii"].cpu().numpy(), vsp=pkg["viewspace_points"].grad.cpu().numpy(),
                         nop=pkg["neural_opacity"].detach().cpu().numpy(), **{k: t.grad.detach().cpu().numpy().copy() for k, t in params.items()})
    g, m = res["glue"], res["ours"]
    assert m["keys"] == g["keys"] == {"render", "render_depth", "uncertainty", "viewspace_points", "visibility_filter", "radii",
                                      "selection_mask", "neural_opacity", "scaling"}
    if np.array_equal(m["mask"], g["mask"]):
        assert m["radii"].shape == g["radii"].shape and m["vsp"].shape == g["vsp"].shape
        assert (m["radii"] != g["radii"]).mean() <= 1e-3
    else:   # a |neural_opacity| < 1e-6 offset selected by one decode and not the other: invisible (alpha < 1/255) either way
        flips = np.nonzero(m["mask"] != g["mask"])[0]
        assert (np.abs(g["nop"].reshape(-1)[flips]) < 1e-6).all(), "selection differs away from zero"
    assert np.abs(m["nop"] - g["nop"]).max() <= 2e-5
    _planes_close(m["image"], g["image"], 1e-4, 1.5 / 255.0, "image")
    _planes_close(m["depth"], g["depth"], 1e-3, 1.5 * 12.0 / 255.0, "depth")
    assert m["vsp"].shape[1] == 3 and np.abs(m["vsp"]).max() > 0
    for k in params:
        tol = 2e-3 * np.abs(g[k]).max() + 1e-12   # as tests/test_decode.py: through the rasterizer's atomics and threshold flips
        assert np.abs(m[k] - g[k]).max() <= tol, (k, float(np.abs(m[k] - g[k]).max()), float(np.abs(g[k]).max()))

    # evaluation branch (train.py:754-761): no_grad, six-key dict
    pc.eval()
    with torch.no_grad():
        pkg = renderer.render(camera, pc, PIPE, bg, visible_mask=vis)
    assert set(pkg.keys()) == {"render", "render_depth", "uncertainty", "viewspace_points", "visibility_filter", "radii"}
    _planes_close(pkg["render"].cpu().numpy(), g["image"], 1e-4, 1.5 / 255.0, "eval image")
    assert pkg["render"].shape == (3, H, W) and pkg["render_depth"].shape == (1, H, W) and pkg["uncertainty"].shape == (1, H, W)
