"""Fused anchor -> neural-Gaussian decode (gscream_b200/decode.py + gsr_decode.cu).

Oracle chain: tests/golden/decode_*.npz hold the outputs and parameter gradients of the REFERENCE's own
`generate_neural_gaussians` (gaussian_renderer/__init__.py:18-102), executed from /root/reference by
tests/golden/make_decode_golden.py in fp32 and fp64.  CPU tests pin the torch restatement (tests/_anchor_decode.py) to
those vectors; GPU tests compare the CUDA path (through the C ABI) with the vectors and, on larger / edge-case inputs,
with the restatement in fp64.

Tolerances (floating point, fp32 compute): forward |ours - ref64| <= 2e-5 * max|ref64|; gradients
<= 1e-4 * max|ref64| (sums over up to A*k terms in a different order than cuBLAS / eager torch).  The selection mask
(neural_opacity > 0) must be identical; a flip is only tolerated where |neural_opacity| < 1e-6 (never observed).
"""
import os

import numpy as np
import pytest
import torch

import _anchor_decode as ad

GOLDEN = ("decode_k10_vis", "decode_k10_all", "decode_k5_vis")
OUT_NAMES = ("xyz", "color", "opacity", "uncertainty", "scaling", "rot", "neural_opacity", "mask")


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    A, k = int(z["A"]), int(z["k"])
    pc = ad.SyntheticAnchors(A, n_offsets=k, seed=0)
    with torch.no_grad():
        for n, p in pc.named_parameters():
            p.copy_(torch.from_numpy(z["p." + n]))
    vis = None if z["vis"].size == 0 else torch.from_numpy(z["vis"])
    return z, pc, vis


def _loss(outs, z, device="cpu"):
    return sum((o * torch.from_numpy(z["u." + n]).to(device=device, dtype=o.dtype)).sum() for n, o in zip(OUT_NAMES[:7], outs[:7]))


@pytest.mark.parametrize("name", GOLDEN)
def test_restatement_matches_reference_function(golden_dir, name):
    """Pins tests/_anchor_decode.generate_neural_gaussians to the reference's own function (same torch ops on CPU)."""
    z, pc, vis = _load(golden_dir, name)
    outs = ad.generate_neural_gaussians(torch.from_numpy(z["campos"]), pc, vis)
    for n, o in zip(OUT_NAMES, outs):
        ref = z["o." + n]
        assert o.shape == ref.shape, n
        if n == "mask":
            assert np.array_equal(o.numpy(), ref)
        else:
            np.testing.assert_allclose(o.detach().numpy(), ref, rtol=1e-6, atol=1e-7, err_msg=n)
    params = dict(pc.named_parameters())
    grads = torch.autograd.grad(_loss(outs, z), list(params.values()), allow_unused=True)
    for (n, p), g in zip(params.items(), grads):
        ref = z["g." + n]
        got = np.zeros_like(ref) if g is None else g.numpy()
        assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-6), n


def test_decode_fails_loudly_without_cuda_tensors():
    from gscream_b200 import decode
    pc = ad.SyntheticAnchors(8)

    class Cam:
        camera_center = torch.zeros(3)

    with pytest.raises(TypeError):
        decode.generate_neural_gaussians(Cam(), pc, None, is_training=True)   # CPU tensors: no CPU decode exists
    pc.use_feat_bank = True
    with pytest.raises(NotImplementedError):
        decode.generate_neural_gaussians(Cam(), pc, None)


# ---- GPU ------------------------------------------------------------------------------------------------------------
class _Cam:
    def __init__(self, c):
        self.camera_center = c


def _ours(pc, campos, vis, dev):
    from gscream_b200 import decode
    return decode.generate_neural_gaussians(_Cam(campos.to(dev)), pc, None if vis is None else vis.to(dev), is_training=True)


def _close(got, ref, tol, name, extra=0.0):
    scale = max(float(np.abs(ref).max()), 1e-6) if ref.size else 1.0
    err = float(np.abs(got - ref).max()) if ref.size else 0.0
    path = os.environ.get("GSR_PARITY_REPORT")     # one JSON line per comparison: the figure behind the assertion
    if path:
        import json
        with open(path, "a") as fh:
            fh.write(json.dumps(dict(case=os.environ.get("PYTEST_CURRENT_TEST", "").split("::")[-1].split(" ")[0], tensor=name,
                                     err_over_scale=err / scale, bound=tol, slack_over_scale=extra / scale)) + "\n")
    assert err <= tol * scale + extra, "%s: max err %.3e vs scale %.3e" % (name, err, scale)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_gpu_decode_matches_reference_golden(golden_dir, name):
    z, pc, vis = _load(golden_dir, name)
    dev = torch.device("cuda")
    pc = pc.to(dev)
    outs = _ours(pc, torch.from_numpy(z["campos"]), vis, dev)
    assert np.array_equal(outs[7].cpu().numpy(), z["o.mask"])                       # selection: exact
    for n, o in zip(OUT_NAMES[:7], outs[:7]):
        ref64, ref32 = z["f64.o." + n], z["o." + n]
        assert tuple(o.shape) == ref32.shape, n
        _close(o.detach().cpu().numpy().astype(np.float64), ref64, 2e-5, n, extra=2 * float(np.abs(ref32 - ref64).max()))
    params = dict(pc.named_parameters())
    grads = torch.autograd.grad(_loss(outs, z, dev), list(params.values()), allow_unused=True)
    for (n, p), g in zip(params.items(), grads):
        ref64, ref32 = z["f64.g." + n], z["g." + n]
        got = np.zeros_like(ref64) if g is None else g.cpu().numpy().astype(np.float64)
        _close(got, ref64, 1e-4, "grad " + n, extra=4 * float(np.abs(ref32 - ref64).max()))


def _compare_with_restatement(A, k, seed, vis_frac, dev, bias_shift=None):
    pc = ad.SyntheticAnchors(A, n_offsets=k, seed=seed)
    if bias_shift is not None:
        with torch.no_grad():
            pc.mlp_opacity[2].bias += bias_shift
    g = torch.Generator().manual_seed(seed + 5)
    vis = None if vis_frac is None else (torch.rand(A, generator=g) < vis_frac)
    campos = torch.tensor([0.05, 0.1, -0.3])
    pc64 = ad.SyntheticAnchors(A, n_offsets=k, seed=seed).double()
    pc64.load_state_dict({n: v.double() for n, v in pc.state_dict().items()})
    ref = ad.generate_neural_gaussians(campos.double(), pc64, vis)
    pcd = pc.to(dev)
    outs = _ours(pcd, campos, vis, dev)
    ref_mask = ref[7].numpy()
    got_mask = outs[7].cpu().numpy()
    if not np.array_equal(got_mask, ref_mask):
        flips = np.nonzero(got_mask != ref_mask)[0]
        assert (np.abs(ref[6].detach().numpy().reshape(-1)[flips]) < 1e-6).all(), "mask differs away from zero"
        pytest.skip("selection flipped on a |neural_opacity| < 1e-6 element (seed-dependent); nothing else is comparable")
    ups = [torch.randn(o.shape, generator=g, dtype=torch.float64) for o in ref[:7]]
    for n, o, r in zip(OUT_NAMES[:7], outs[:7], ref[:7]):
        assert tuple(o.shape) == tuple(r.shape), n
        _close(o.detach().cpu().numpy().astype(np.float64), r.detach().numpy(), 2e-5, n)
    p64 = dict(pc64.named_parameters())
    g64 = torch.autograd.grad(sum((o * u).sum() for o, u in zip(ref[:7], ups)), list(p64.values()), allow_unused=True)
    pd = dict(pcd.named_parameters())
    gd = torch.autograd.grad(sum((o * u.to(dev).float()).sum() for o, u in zip(outs[:7], ups)), list(pd.values()), allow_unused=True)
    for (n, _), a, b in zip(p64.items(), g64, gd):
        ra = np.zeros(tuple(p64[n].shape)) if a is None else a.numpy()
        rb = np.zeros(tuple(p64[n].shape)) if b is None else b.cpu().numpy().astype(np.float64)
        _close(rb, ra, 1e-4, "grad " + n)
    return int(ref_mask.sum())


@pytest.mark.gpu
@pytest.mark.parametrize("A,k,seed,vis_frac", [(20000, 10, 21, 0.7), (4099, 10, 22, None), (1, 10, 23, None), (5, 16, 24, 0.9),
                                                (37, 1, 25, None), (1000, 4, 26, 0.05),
                                                # k > 10: four tiles per CTA iteration and the 7-row-tile instantiation of the backward
                                                (3001, 16, 27, 0.8), (2050, 11, 28, None)])
def test_gpu_decode_matches_fp64_restatement(A, k, seed, vis_frac):
    P = _compare_with_restatement(A, k, seed, vis_frac, torch.device("cuda"))
    assert P >= 0


@pytest.mark.gpu
def test_gpu_decode_empty_selections():
    dev = torch.device("cuda")
    # no offset survives the opacity test: P == 0, gradients only through neural_opacity
    assert _compare_with_restatement(64, 10, 31, None, dev, bias_shift=-50.0) == 0
    # no anchor is visible
    pc = ad.SyntheticAnchors(16, seed=1).to(dev)
    outs = _ours(pc, torch.zeros(3), torch.zeros(16, dtype=torch.bool), dev)
    assert outs[0].shape == (0, 3) and outs[6].shape == (0, 1) and outs[7].shape == (0,)
    (outs[0].sum() + outs[6].sum()).backward()
    assert float(pc._anchor_feat.grad.abs().max()) == 0.0


@pytest.mark.gpu
def test_gpu_decode_feeds_rasterizer_end_to_end():
    """decode -> GaussianRasterizer -> loss -> backward to the anchor parameters, against the torch decode feeding the same
    rasterizer (the caller glue of gaussian_renderer/__init__.py:104-179)."""
    from gscream_b200 import decode, scenes
    from gscream_b200 import rasterizer as ours
    dev = torch.device("cuda")
    W, H = 320, 192
    cam = scenes.make_camera(W, H)
    bg = torch.zeros(3, device=dev)
    target = torch.rand(3, H, W, generator=torch.Generator().manual_seed(1)).to(dev)
    res = {}
    for which in ("torch", "fused"):
        pc = ad.SyntheticAnchors(3000, seed=41, tanfov=(cam["tanfovx"], cam["tanfovy"])).to(dev)
        vis, _, _ = ad.prefilter_position2D(ours, cam, pc, bg)
        if which == "torch":
            dec = ad.generate_neural_gaussians(cam["campos"].to(dev), pc, vis)
        else:
            dec = decode.generate_neural_gaussians(_Cam(cam["campos"].to(dev)), pc, vis, is_training=True)
        xyz, color, opacity, unc, scaling, rot, nop, mask = dec
        ssp = torch.zeros_like(xyz, requires_grad=True)
        rast = ours.GaussianRasterizer(raster_settings=ad.make_settings(ours, cam, bg, dev))
        image, depth, uncer, radii = rast(means3D=xyz, means2D=ssp, shs=None, colors_precomp=color, opacities=opacity, uncertainties=unc,
                                          scales=scaling, rotations=rot, cov3D_precomp=None)
        loss = (image - target).abs().mean() + 0.1 * depth.mean() + 0.05 * uncer.mean() + 0.01 * scaling.prod(dim=1).mean()
        loss.backward()
        res[which] = dict(loss=float(loss), image=image.detach().cpu().numpy(), mask=mask.cpu().numpy(),
                          grads={n: p.grad.detach().cpu().numpy() for n, p in pc.named_parameters() if p.grad is not None})
    assert np.array_equal(res["torch"]["mask"], res["fused"]["mask"])
    assert abs(res["torch"]["loss"] - res["fused"]["loss"]) <= 1e-5 * abs(res["torch"]["loss"])
    _close(res["fused"]["image"], res["torch"]["image"], 1e-4, "image")
    for n, g in res["torch"]["grads"].items():
        _close(res["fused"]["grads"][n], g, 2e-3, "grad " + n)   # through the rasterizer's atomics and a loss that re-tiles on 1-ulp xyz changes
