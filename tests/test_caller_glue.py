"""BASELINE.json configs[0] (anchor -> Gaussian decode for 1k anchors on CPU torch, plumbing only) and, on the GPU, the same
caller glue driving this repo's rasterizer and the reference build side by side (gradients back to the anchor parameters)."""
import numpy as np
import pytest
import torch

import _anchor_decode as ad
from gscream_b200 import scenes


def test_config1_decode_1k_anchors_cpu():
    A, k = 1000, 10
    pc = ad.SyntheticAnchors(A, seed=3)
    cam = scenes.make_camera(1008, 567)
    xyz, color, opacity, unc, scaling, rot, neural_opacity, mask = ad.generate_neural_gaussians(cam["campos"], pc)
    n = int(mask.sum())
    assert neural_opacity.shape == (A * k, 1) and mask.shape == (A * k,) and 0 < n < A * k
    assert xyz.shape == (n, 3) and color.shape == (n, 3) and opacity.shape == (n, 1) and unc.shape == (n, 1)
    assert scaling.shape == (n, 3) and rot.shape == (n, 4)
    assert (opacity > 0).all() and (opacity <= 1).all()              # tanh output, masked at > 0
    assert (color >= 0).all() and (color <= 1).all() and (unc >= 0).all() and (unc <= 1).all()
    assert torch.allclose(rot.norm(dim=1), torch.ones(n), atol=1e-5)  # caller normalises; the rasterizer does not (forward.cu:129)
    # neural Gaussian j of anchor a sits at anchor + offset * scaling[:, :3]
    idx = torch.nonzero(mask).flatten()
    a, j = idx // k, idx % k
    expect = pc._anchor[a] + pc._offset[a, j] * pc.get_scaling[a, :3]
    assert torch.allclose(xyz, expect, atol=1e-6)
    assert (scaling < pc.get_scaling[a, 3:]).all()                     # x sigmoid
    # a visible_mask restricts the decode to those anchors
    vm = torch.zeros(A, dtype=torch.bool)
    vm[::3] = True
    out = ad.generate_neural_gaussians(cam["campos"], pc, vm)
    assert out[6].shape == (int(vm.sum()) * k, 1)


@pytest.mark.gpu
def test_caller_glue_same_results_with_reference_build():
    import _ref_utils as ru
    from gscream_b200 import rasterizer as ours
    assert torch.cuda.is_available()
    if not ru.ref_available(3):
        pytest.skip("oracle/_ref/dgr3 not built")
    ref = ru.load_ref(3)
    dev = torch.device("cuda")
    W, H = 1008, 567
    cam = scenes.make_camera(W, H)
    bg = torch.zeros(3, device=dev)
    target = torch.rand(3, H, W, generator=torch.Generator().manual_seed(1)).to(dev)
    # ONE anchor model and ONE decode graph feed both rasterizers, so both see bit-identical Gaussians (two separate
    # decodes may differ in the last bit: cuBLAS picks its algorithm per call).
    pc = ad.SyntheticAnchors(20000, seed=11, tanfov=(cam["tanfovx"], cam["tanfovy"])).to(dev)
    params = dict(g_feat=pc._anchor_feat, g_off=pc._offset, g_anchor=pc._anchor, g_scaling=pc._scaling, g_w=pc.mlp_cov[2].weight)
    res = {}
    filt = {name: ad.prefilter_position2D(mod, cam, pc, bg) for name, mod in (("ref", ref), ("ours", ours))}
    vis = filt["ref"][0]
    dec = ad.generate_neural_gaussians(cam["campos"].to(dev), pc, vis)
    xyz, color, opacity, uncertainty, scaling, rot = dec[:6]
    for name, mod in (("ref", ref), ("ours", ours)):
        ssp = torch.zeros_like(xyz, requires_grad=True)
        rast = mod.GaussianRasterizer(raster_settings=ad.make_settings(mod, cam, bg, dev))
        image, depth, uncer, radii = rast(means3D=xyz, means2D=ssp, shs=None, colors_precomp=color, opacities=opacity,
                                          uncertainties=uncertainty, scales=scaling, rotations=rot, cov3D_precomp=None)
        loss = (image - target).abs().mean() + 0.1 * depth.mean() + 0.05 * uncer.mean()
        for t in params.values():
            t.grad = None
        loss.backward(retain_graph=True)
        res[name] = dict(vis=filt[name][0].cpu().numpy(), x=filt[name][1].cpu().numpy(), y=filt[name][2].cpu().numpy(),
                         image=image.detach().cpu().numpy(), radii=radii.cpu().numpy(), vsp=ssp.grad.cpu().numpy(), loss=float(loss.detach()),
                         **{k: t.grad.detach().cpu().numpy().copy() for k, t in params.items()})
    r, m = res["ref"], res["ours"]
    assert np.array_equal(m["vis"], r["vis"]) and np.array_equal(m["x"], r["x"]) and np.array_equal(m["y"], r["y"])  # prefilter: bit-exact
    assert np.array_equal(m["radii"], r["radii"])
    assert np.abs(m["image"] - r["image"]).max() <= 1e-5
    assert abs(m["loss"] - r["loss"]) <= 1e-6 * abs(r["loss"])
    for k in ("vsp", "g_feat", "g_off", "g_anchor", "g_scaling", "g_w"):
        tol = 1e-4 * np.abs(r[k]).max() + 1e-12   # the reference's own atomics jitter, amplified through the MLP backward
        assert np.abs(m[k] - r[k]).max() <= tol, (k, float(np.abs(m[k] - r[k]).max()), float(np.abs(r[k]).max()))
