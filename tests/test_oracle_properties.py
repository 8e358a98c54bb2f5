"""CPU: self-consistency of the oracle — its backward is the gradient of its forward, its float32 and
float64 builds agree, the anchor-filter modes agree with the full preprocess, and edge cases behave as
the reference's code says they must."""
import numpy as np
import pytest

from gscream_b200 import scenes
from oracle.oracle import Oracle


def _args(P, W, H, C, seed, smult=2.0, bg=0.2, yaw=0.0):
    s = scenes.make_scene(P, W, H, C, seed, scale_mult=smult, bg_value=bg)
    cam = scenes.make_camera(W, H, yaw_deg=yaw)
    a = dict(means3D=s["means3D"].numpy(), colors_precomp=s["colors"].numpy(), opacities=s["opacities"].numpy(),
             uncertainties=s["uncertainties"].numpy(), scales=s["scales"].numpy(), rotations=s["rotations"].numpy(),
             viewmatrix=cam["viewmatrix"].numpy(), projmatrix=cam["projmatrix"].numpy(), bg=s["bg"].numpy(), W=W, H=H,
             tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"])
    return a, [g.numpy() for g in scenes.make_upstream_grads(C, W, H, seed)]


def _bargs(a, grads):
    return dict(means3D=a["means3D"], colors_precomp=a["colors_precomp"], scales=a["scales"], rotations=a["rotations"],
                viewmatrix=a["viewmatrix"], projmatrix=a["projmatrix"], bg=a["bg"], W=a["W"], H=a["H"], tanfovx=a["tanfovx"],
                tanfovy=a["tanfovy"], dL_dcolor=grads[0], dL_ddepth=grads[1], dL_dunc=grads[2])


def test_f32_and_f64_builds_agree():
    a, grads = _args(500, 80, 48, 3, 3)
    f32, f64 = Oracle("f32").forward(**a), Oracle("f64").forward(**a)
    assert np.array_equal(f32["radii"], f64["radii"])
    assert np.array_equal(f32["point_list"], f64["point_list"])
    assert (f32["n_contrib"] != f64["n_contrib"]).sum() <= 2
    for k in ("color", "depth", "uncertainty"):
        assert np.abs(f32[k] - f64[k]).max() <= 2e-5 * max(1.0, np.abs(f64[k]).max())
    g32 = Oracle("f32").backward(f32, **_bargs(a, grads))
    g64 = Oracle("f64").backward(f64, **_bargs(a, grads))
    for k in g64:
        assert np.abs(g32[k] - g64[k]).max() <= 2e-4 * np.abs(g64[k]).max() + 1e-12, k


def test_render_backward_is_gradient_of_render_forward():
    """Central differences on the float64 oracle at the blend stage (tiny steps; the 1/255, 1e-4 and 0.99
    thresholds make the function only piecewise smooth, so a few samples may sit on a discontinuity)."""
    a, grads = _args(300, 64, 48, 3, 9, smult=2.5, bg=0.3)
    o = Oracle("f64")
    f = o.forward(**a)
    b = o.backward(f, **_bargs(a, grads))
    W, H = a["W"], a["H"]

    def loss(co, xy, unc, feat):
        r = o.render_forward(W, H, f["ranges"], f["point_list"], xy, feat, f["depths"], unc, co, a["bg"])
        return (r["color"] * grads[0]).sum() + (r["depth"] * grads[1]).sum() + (r["uncertainty"] * grads[2]).sum()

    vis = np.where(f["radii"] > 0)[0]
    rng = np.random.default_rng(0)
    good = total = 0
    for i in rng.choice(vis, 24, replace=False):
        co = f["conic_opacity"].astype(np.float64)
        checks = []
        for comp, ana in ((3, b["dL_dopacity"][i, 0]), (0, b["dL_dconic"][i, 0]), (2, b["dL_dconic"][i, 3]),
                          (1, 2.0 * b["dL_dconic"][i, 1])):  # the reference stores HALF of d/d(conic.y) (backward.cu:597)
            h = 1e-7 * max(abs(co[i, comp]), 1e-3)
            c1, c2 = co.copy(), co.copy()
            c1[i, comp] += h
            c2[i, comp] -= h
            fd = (loss(c1, f["xy"], f["gauss_uncertainty"], a["colors_precomp"]) - loss(c2, f["xy"], f["gauss_uncertainty"], a["colors_precomp"])) / (2 * h)
            checks.append((fd, ana))
        xy1, xy2 = f["xy"].copy(), f["xy"].copy()
        xy1[i, 0] += 1e-6
        xy2[i, 0] -= 1e-6
        fd = (loss(co, xy1, f["gauss_uncertainty"], a["colors_precomp"]) - loss(co, xy2, f["gauss_uncertainty"], a["colors_precomp"])) / 2e-6
        checks.append((fd * 0.5 * W, b["dL_dmean2D"][i, 0]))  # stored w.r.t. NDC: x 0.5*W (backward.cu:486,592)
        u1, u2 = f["gauss_uncertainty"].copy(), f["gauss_uncertainty"].copy()
        u1[i] += 1e-6
        u2[i] -= 1e-6
        checks.append(((loss(co, f["xy"], u1, a["colors_precomp"]) - loss(co, f["xy"], u2, a["colors_precomp"])) / 2e-6, b["dL_duncertainty"][i, 0]))
        for fd, ana in checks:
            total += 1
            good += abs(fd - ana) <= 2e-3 * max(abs(fd), abs(ana)) + 1e-12
    assert good >= 0.9 * total, (good, total)


def test_filter_modes_agree_with_full_preprocess():
    a, _ = _args(800, 150, 83, 3, 4, smult=3.0, yaw=6.0)
    o = Oracle("f32")
    kw = dict(means3D=a["means3D"], scales=a["scales"], rotations=a["rotations"], opacities=a["opacities"],
              uncertainties=a["uncertainties"], viewmatrix=a["viewmatrix"], projmatrix=a["projmatrix"], W=a["W"], H=a["H"],
              tanfovx=a["tanfovx"], tanfovy=a["tanfovy"])
    full, vis, pos = o.preprocess(mode=0, **kw), o.preprocess(mode=1, **kw), o.preprocess(mode=2, **kw)
    assert np.array_equal(full["radii"], vis["radii"]) and np.array_equal(full["radii"], pos["radii"])
    m = full["radii"] > 0
    assert np.array_equal(pos["pos2d_x"][m], full["xy"][m, 0]) and np.array_equal(pos["pos2d_y"][m], full["xy"][m, 1])
    assert not pos["pos2d_x"][~m].any() and not pos["pos2d_y"][~m].any()  # zero where culled (forward.cu:385-390)
    # markVisible is only the near-plane test (auxiliary.h:154): a superset of radii > 0
    present = o.mark_visible(a["means3D"], a["viewmatrix"], a["projmatrix"])
    assert present[m].all()
    z_view = a["means3D"] @ a["viewmatrix"][:3, 2] + a["viewmatrix"][3, 2]
    assert np.array_equal(present, ~(z_view.astype(np.float32) <= np.float32(0.2)))


def test_empty_and_fully_culled_inputs():
    o = Oracle("f32")
    W, H = 48, 32
    cam = scenes.make_camera(W, H)
    P = 7
    means = np.zeros((P, 3), np.float32)
    means[:, 2] = -1.0  # behind the camera: every Gaussian is culled
    kw = dict(means3D=means, colors_precomp=np.ones((P, 3), np.float32), opacities=np.full((P, 1), 0.5, np.float32),
              uncertainties=np.zeros((P, 1), np.float32), scales=np.full((P, 3), 0.1, np.float32),
              rotations=np.tile(np.array([[1, 0, 0, 0]], np.float32), (P, 1)), viewmatrix=cam["viewmatrix"].numpy(),
              projmatrix=cam["projmatrix"].numpy(), bg=np.array([0.1, 0.2, 0.3], np.float32), W=W, H=H,
              tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"])
    f = o.forward(**kw)
    assert f["num_rendered"] == 0 and not f["radii"].any() and not f["ranges"].any()
    assert np.allclose(f["color"], np.array([0.1, 0.2, 0.3], np.float32)[:, None, None])  # T*bg with T=1
    assert not f["depth"].any() and not f["uncertainty"].any() and (f["final_T"] == 1).all()


def test_single_gaussian_closed_form():
    """One isotropic splat on the optical axis: alpha at the centre pixel = min(0.99, opacity * exp(power))."""
    o = Oracle("f64")
    W = H = 33
    cam = scenes.make_camera(W, H)
    kw = dict(means3D=np.array([[0, 0, 5.0]], np.float32), colors_precomp=np.array([[1.0, 0.5, 0.25]], np.float32),
              opacities=np.array([[0.6]], np.float32), uncertainties=np.array([[0.7]], np.float32),
              scales=np.array([[0.2, 0.2, 0.2]], np.float32), rotations=np.array([[1, 0, 0, 0]], np.float32),
              viewmatrix=cam["viewmatrix"].numpy(), projmatrix=cam["projmatrix"].numpy(), bg=np.zeros(3, np.float32), W=W, H=H,
              tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"])
    f = o.forward(**kw)
    # ndc2Pix(0, 33) = ((0+1)*33-1)/2 = 16 exactly -> centre pixel (16,16), d = 0, power = 0
    assert np.allclose(f["xy"][0], [16.0, 16.0])
    a0 = np.float32(0.6)
    assert np.isclose(f["color"][0, 16, 16], a0 * 1.0) and np.isclose(f["depth"][0, 16, 16], a0 * 5.0)
    assert np.isclose(f["uncertainty"][0, 16, 16], a0 * np.float32(0.7)) and np.isclose(f["final_T"][16 * W + 16], 1 - a0)
    focal = W / (2 * cam["tanfovx"])
    cov = (focal * 0.2 / 5.0) ** 2 + 0.3  # EWA footprint + 0.3 low-pass (forward.cu:112-113)
    assert np.isclose(f["conic_opacity"][0, 0], 1.0 / cov, rtol=1e-5)
    # isotropic: mid^2 - det = 0, but the reference floors the discriminant at 0.1 (forward.cu:245-246)
    assert f["radii"][0] == int(np.ceil(3 * np.sqrt(cov + np.sqrt(0.1))))
