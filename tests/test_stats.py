"""Fused densification statistics (gscream_b200/stats.py + training_statis_kernel).

tests/golden/statis_*.npz hold the four statistics buffers before and after the REFERENCE's own
GaussianModel.training_statis (scene/gaussian_model.py:729-757), executed from /root/reference by
tests/golden/make_statis_golden.py.  Counters must match exactly; the accumulated gradient norms within 1e-6 relative
(sqrt(x*x + y*y) with or without FMA contraction)."""
import os

import numpy as np
import pytest
import torch

GOLDEN = ("statis_k10", "statis_k4_sparse", "statis_all")
BUFS = ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom")


class _Model:
    pass


def _numpy_restatement(z):
    """Index-level restatement of the reference method (numpy), the CPU check of the golden files' self-consistency."""
    k = int(z["k"])
    out = {n: z["in." + n].copy() for n in BUFS}
    vis_ids = np.nonzero(z["vis"])[0]
    op = np.maximum(z["opacity"].reshape(-1, k), 0)
    out["opacity_accum"][vis_ids, 0] += op.sum(1)
    out["anchor_demon"][vis_ids, 0] += 1
    sel_pos = np.nonzero(z["sel"])[0]                      # positions in the (visible anchor, offset) grid
    kept = sel_pos[z["upd"]]
    glob = vis_ids[kept // k] * k + kept % k
    norm = np.sqrt((z["grad"][z["upd"], :2].astype(np.float32) ** 2).sum(1, dtype=np.float32))
    out["offset_gradient_accum"][glob, 0] += norm
    out["offset_denom"][glob, 0] += 1
    return out


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_is_what_the_method_documents(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    out = _numpy_restatement(z)
    for n in BUFS:
        np.testing.assert_allclose(out[n], z["out." + n], rtol=1e-6, atol=1e-7, err_msg=n)


def test_stats_fail_loudly_on_cpu():
    from gscream_b200 import stats
    m = _Model()
    m.n_offsets = 2
    m.opacity_accum = torch.zeros(4, 1)
    with pytest.raises(TypeError):
        stats.training_statis(m, torch.zeros(1, 3), torch.zeros(2, 1), torch.zeros(1, dtype=torch.bool), torch.zeros(2, dtype=torch.bool),
                              torch.zeros(4, dtype=torch.bool))


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_gpu_training_statis_matches_reference_golden(golden_dir, name):
    from gscream_b200 import stats
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    dev = torch.device("cuda")
    m = _Model()
    m.n_offsets = int(z["k"])
    for n in BUFS:
        setattr(m, n, torch.from_numpy(z["in." + n]).to(dev))
    pts = torch.zeros(z["grad"].shape, device=dev, requires_grad=True)
    pts.grad = torch.from_numpy(z["grad"]).to(dev)
    stats.training_statis(m, pts, torch.from_numpy(z["opacity"]).to(dev), torch.from_numpy(z["upd"]).to(dev),
                          torch.from_numpy(z["sel"]).to(dev), torch.from_numpy(z["vis"]).to(dev))
    assert np.array_equal(m.anchor_demon.cpu().numpy(), z["out.anchor_demon"])
    assert np.array_equal(m.offset_denom.cpu().numpy(), z["out.offset_denom"])
    np.testing.assert_allclose(m.opacity_accum.cpu().numpy(), z["out.opacity_accum"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(m.offset_gradient_accum.cpu().numpy(), z["out.offset_gradient_accum"], rtol=2e-6, atol=1e-9)


@pytest.mark.gpu
def test_gpu_training_statis_after_a_real_render():
    """Wired to the real producers: prefilter -> fused decode -> rasterizer -> backward -> statistics, against the eager index
    code of the reference method (restated in numpy above) on the same tensors."""
    import _anchor_decode as ad
    from gscream_b200 import decode, scenes, stats
    from gscream_b200 import rasterizer as ours
    dev = torch.device("cuda")
    W, H, A, k = 320, 192, 4000, 10
    cam = scenes.make_camera(W, H)
    bg = torch.zeros(3, device=dev)
    pc = ad.SyntheticAnchors(A, n_offsets=k, seed=5, tanfov=(cam["tanfovx"], cam["tanfovy"])).to(dev)
    vis, _, _ = ad.prefilter_position2D(ours, cam, pc, bg)

    class Cam:
        camera_center = cam["campos"].to(dev)

    xyz, color, opacity, unc, scaling, rot, nop, mask = decode.generate_neural_gaussians(Cam, pc, vis, is_training=True)
    ssp = torch.zeros_like(xyz, requires_grad=True)
    rast = ours.GaussianRasterizer(raster_settings=ad.make_settings(ours, cam, bg, dev))
    image, depth, uncer, radii = rast(means3D=xyz, means2D=ssp, shs=None, colors_precomp=color, opacities=opacity, uncertainties=unc,
                                      scales=scaling, rotations=rot, cov3D_precomp=None)
    (image.mean() + depth.mean()).backward()
    m = _Model()
    m.n_offsets = k
    g = torch.Generator().manual_seed(9)
    z = {"k": k, "vis": vis.cpu().numpy(), "opacity": nop.detach().cpu().numpy(), "sel": mask.cpu().numpy(), "upd": (radii > 0).cpu().numpy(),
         "grad": ssp.grad.cpu().numpy()}
    for n, shape in (("opacity_accum", (A, 1)), ("anchor_demon", (A, 1)), ("offset_gradient_accum", (A * k, 1)), ("offset_denom", (A * k, 1))):
        t = torch.rand(shape, generator=g)
        z["in." + n] = t.numpy().copy()
        setattr(m, n, t.to(dev))
    stats.training_statis(m, ssp, nop, radii > 0, mask, vis)
    ref = _numpy_restatement(z)
    for n in BUFS:
        np.testing.assert_allclose(getattr(m, n).cpu().numpy(), ref[n], rtol=2e-6, atol=1e-9, err_msg=n)
    assert int((ref["offset_denom"] != z["in.offset_denom"]).sum()) > 100


@pytest.mark.parametrize("name", GOLDEN)
def test_eager_restatement_is_bit_identical_to_the_reference_method(golden_dir, name):
    """Pins tests/_anchor_decode.training_statis_eager (the reference arm of bench.py's train-step loop) to the reference method."""
    import _anchor_decode as ad
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    m = _Model()
    m.n_offsets = int(z["k"])
    for n in BUFS:
        setattr(m, n, torch.from_numpy(z["in." + n]).clone())
    pts = torch.zeros(z["grad"].shape, requires_grad=True)
    pts.grad = torch.from_numpy(z["grad"])
    ad.training_statis_eager(m, pts, torch.from_numpy(z["opacity"]), torch.from_numpy(z["upd"]), torch.from_numpy(z["sel"]), torch.from_numpy(z["vis"]))
    for n in BUFS:
        assert np.array_equal(getattr(m, n).numpy(), z["out." + n]), n
