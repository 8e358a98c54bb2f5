"""Fused L1 + SSIM image loss (gscream_b200/losses.py + gsr_loss.cu).

Oracle chain: tests/golden/loss_*.npz hold values and image gradients of the REFERENCE's own l1_loss / l1_loss_masked /
ssim / ssim_masked (utils/loss_utils.py), executed from /root/reference by tests/golden/make_loss_golden.py in fp32 and
fp64.  CPU tests pin the torch restatement used by bench.py's train-step loop (tests/_train_step.ssim) to them; GPU tests
compare the CUDA path (through the C ABI) with the vectors and, at the train-step size, with the restatement in fp64.

Tolerances: loss values |ours - ref64| <= 1e-5 * |ref64|; image gradients <= 1e-5 * max|ref64| (+ 2x the reference's own
fp32-vs-fp64 difference).
"""
import os

import numpy as np
import pytest
import torch

import _train_step as ts

GOLDEN = ("loss_rgb", "loss_rgb_masked1", "loss_rgb_masked3", "loss_tiny")


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    mask = None if z["mask"].size == 0 else torch.from_numpy(z["mask"])
    return z, torch.from_numpy(z["img"]), torch.from_numpy(z["gt"]), mask


@pytest.mark.parametrize("name", ("loss_rgb", "loss_tiny"))
def test_restatement_matches_reference_ssim(golden_dir, name):
    z, img, gt, _ = _load(golden_dir, name)
    x = img.clone().requires_grad_(True)
    s = ts.ssim(x, gt, ts._gaussian_window(channels=img.shape[0]))
    g, = torch.autograd.grad(s, x)
    assert abs(float(s) - float(z["ssim"])) <= 1e-6
    assert np.abs(g.numpy() - z["g_ssim"]).max() <= 1e-6 * max(np.abs(z["g_ssim"]).max(), 1e-9) + 1e-10


def test_losses_fail_loudly_without_cuda_tensors():
    from gscream_b200 import losses
    with pytest.raises(TypeError):
        losses.ssim(torch.rand(3, 8, 8), torch.rand(3, 8, 8))
    with pytest.raises(NotImplementedError):
        losses.ssim(torch.rand(3, 8, 8), torch.rand(3, 8, 8), window_size=7)
    # the taps are the reference's (loss_utils.py:112-114): normalised, symmetric, 11 of them
    assert losses._TAPS11.shape == (11,) and abs(float(losses._TAPS11.sum()) - 1.0) < 1e-6
    assert np.array_equal(losses._TAPS11, losses._TAPS11[::-1])


def _close(got, ref, tol, name, extra=0.0):
    scale = max(float(np.abs(ref).max()), 1e-12)
    err = float(np.abs(got - ref).max())
    assert err <= tol * scale + extra, "%s: max err %.3e vs scale %.3e" % (name, err, scale)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_gpu_losses_match_reference_golden(golden_dir, name):
    from gscream_b200 import losses
    z, img, gt, mask = _load(golden_dir, name)
    dev = torch.device("cuda")
    x = img.to(dev).requires_grad_(True)
    y = gt.to(dev)
    m = None if mask is None else mask.to(dev)
    if m is None:
        s, l = losses.ssim(x, y), losses.l1_loss(x, y)
    else:
        s, l = losses.ssim_masked(x, y, m), losses.l1_loss_masked(x, y, m)
    gs, = torch.autograd.grad(s, x)
    gl, = torch.autograd.grad(l, x)
    assert abs(float(s) - float(z["f64.ssim"])) <= 1e-5 * abs(float(z["f64.ssim"]))
    assert abs(float(l) - float(z["f64.l1"])) <= 1e-5 * abs(float(z["f64.l1"]))
    _close(gs.cpu().numpy().astype(np.float64), z["f64.g_ssim"], 1e-5, "d ssim / d image", extra=2 * float(np.abs(z["g_ssim"] - z["f64.g_ssim"]).max()))
    _close(gl.cpu().numpy().astype(np.float64), z["f64.g_l1"], 1e-6, "d l1 / d image")
    # both terms from one node, arbitrary upstream weights (train.py:539: (1 - lambda) * L1 + lambda * (1 - ssim))
    x2 = img.to(dev).requires_grad_(True)
    s2, l2 = losses.l1_ssim(x2, y, m)
    (0.8 * l2 + 0.2 * (1.0 - s2)).backward()
    ref = 0.8 * z["f64.g_l1"] - 0.2 * z["f64.g_ssim"]
    _close(x2.grad.cpu().numpy().astype(np.float64), ref, 1e-5, "combined", extra=2 * float(np.abs(z["g_ssim"] - z["f64.g_ssim"]).max()))


@pytest.mark.gpu
def test_gpu_losses_train_step_size_vs_fp64_restatement():
    from gscream_b200 import losses
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    C, H, W = 3, 567, 1008
    img = torch.rand(C, H, W, generator=g)
    gt = (img + 0.1 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    x64 = img.double().to(dev).requires_grad_(True)
    s64 = ts.ssim(x64, gt.double().to(dev), ts._gaussian_window(device=dev).double())
    l64 = (x64 - gt.double().to(dev)).abs().mean()
    g64, = torch.autograd.grad(0.8 * l64 + 0.2 * (1 - s64), x64)
    x = img.to(dev).requires_grad_(True)
    s, l = losses.l1_ssim(x, gt.to(dev))
    (0.8 * l + 0.2 * (1 - s)).backward()
    assert abs(float(s) - float(s64)) <= 1e-5 * abs(float(s64)) and abs(float(l) - float(l64)) <= 1e-5 * abs(float(l64))
    _close(x.grad.double().cpu().numpy(), g64.cpu().numpy(), 1e-5, "train-step size gradient")
    # batched input [B, C, H, W]: planes = B * C
    xb = torch.stack([img, gt]).to(dev).requires_grad_(True)
    yb = torch.stack([gt, img]).to(dev)
    sb, lb = losses.l1_ssim(xb, yb)
    assert abs(float(sb) - float(s64)) <= 2e-5 * abs(float(s64))   # SSIM and L1 are symmetric in (x, y)
    assert abs(float(lb) - float(l64)) <= 2e-5 * abs(float(l64))


# ---- scale / shift aligned depth L1 -------------------------------------------------------------------------------
DEPTH_GOLDEN = ("depth_ref_view", "depth_other_view", "depth_nomask_b2")


@pytest.mark.parametrize("name", DEPTH_GOLDEN)
def test_restatement_of_scale_and_shift_matches_reference(golden_dir, name):
    """Pins tests/_train_step.compute_scale_and_shift (the eager depth term of bench.py's reference arm) to the reference's."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    d, y = torch.from_numpy(z["depth"]), torch.from_numpy(z["target"])
    m = torch.ones_like(d) if z["fit"].size == 0 else torch.from_numpy(z["fit"])
    s, t = ts.compute_scale_and_shift(d, y, m)
    np.testing.assert_allclose(s.abs().numpy(), z["scale"], rtol=1e-6)
    np.testing.assert_allclose(t.numpy(), z["shift"], rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", DEPTH_GOLDEN)
def test_gpu_aligned_depth_l1_matches_reference_golden(golden_dir, name):
    from gscream_b200 import losses
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    dev = torch.device("cuda")
    d = torch.from_numpy(z["depth"]).to(dev).requires_grad_(True)
    y = torch.from_numpy(z["target"]).to(dev)
    fit = None if z["fit"].size == 0 else torch.from_numpy(z["fit"]).to(dev)
    loss = losses.aligned_depth_l1(d, y, fit, fit if int(z["masked_loss"]) else None)
    g, = torch.autograd.grad(3.0 * loss, d)            # a non-trivial upstream factor
    assert abs(float(loss) - float(z["f64.loss"])) <= 1e-5 * abs(float(z["f64.loss"]))
    ref = 3.0 * z["f64.grad"]
    extra = 6.0 * float(np.abs(z["grad"] - z["f64.grad"]).max())
    _close(g.cpu().numpy().astype(np.float64), ref, 1e-5, "d loss / d depth", extra=extra)


@pytest.mark.gpu
def test_gpu_aligned_depth_l1_train_step_size_vs_fp64():
    from gscream_b200 import losses
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(8)
    H, W = 567, 1008
    depth = 2.0 + 8.0 * torch.rand(1, H, W, generator=g)
    target = 2.0 + 10.0 * torch.rand(1, H, W, generator=g)
    valid = (torch.rand(1, H, W, generator=g) > 0.1).float()
    x64 = depth.double().to(dev).requires_grad_(True)
    s, t = ts.compute_scale_and_shift(x64, target.double().to(dev), valid.double().to(dev))
    l64 = (s.abs().view(-1, 1, 1) * x64 + t.view(-1, 1, 1) - target.double().to(dev)).abs().mean()
    g64, = torch.autograd.grad(l64, x64)
    x = depth.to(dev).requires_grad_(True)
    l = losses.aligned_depth_l1(x, target.to(dev), valid.to(dev))
    l.backward()
    assert abs(float(l) - float(l64)) <= 1e-5 * abs(float(l64))
    _close(x.grad.double().cpu().numpy(), g64.cpu().numpy(), 2e-5, "aligned depth gradient at train-step size")


# ---- four-scale gradient-matching loss from the same fit (train.py:232-251, :556-560, :571-574) ------------------------------
DEPTHGRAD_GOLDEN = ("depthgrad_ref_view", "depthgrad_other_view", "depthgrad_b2_holes", "depthgrad_plain_b2", "depthgrad_tiny",
                    "depthgrad_one_scale")


def _load_depthgrad(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    fit = None if z["fit"].size == 0 else z["fit"]
    return z, fit, (fit if int(z["masked_l1"]) else None), z["grad_mask"], int(z["n_scales"]), bool(int(z["aligned"]))


@pytest.mark.parametrize("name", DEPTHGRAD_GOLDEN)
def test_fused_depth_loss_closed_forms_match_reference_autograd(golden_dir, name):
    """CPU: the closed forms the CUDA kernels evaluate (tests/_depthgrad_emul.py), including the hand-derived gradient through
    the scale/shift fit, against the reference's own functions and autograd in fp64."""
    import _depthgrad_emul as em
    z, fit, l1m, gm, n_scales, aligned = _load_depthgrad(golden_dir, name)
    l1, gl, grad = em.fused_depth_losses(z["depth"], z["target"], fit, l1m, gm, n_scales, aligned, up_l1=0.7, up_gl=1.9)
    assert abs(gl - float(z["f64.gl"])) <= 1e-9 * max(abs(float(z["f64.gl"])), 1e-12)
    if aligned:
        assert abs(l1 - float(z["f64.l1"])) <= 1e-9 * abs(float(z["f64.l1"]))
    ref = 0.7 * z["f64.g_l1"] + 1.9 * z["f64.g_gl"]
    _close(grad, ref, 1e-8, "closed-form gradient")


@pytest.mark.parametrize("name", ("depthgrad_plain_b2", "depthgrad_tiny", "depthgrad_one_scale"))
def test_restatement_of_gradient_loss_matches_reference(golden_dir, name):
    """Pins tests/_train_step.gradient_loss (the eager depth-smoothness term of bench.py's reference arm) to the reference's."""
    z, fit, l1m, gm, n_scales, aligned = _load_depthgrad(golden_dir, name)
    assert not aligned
    d, y, m = torch.from_numpy(z["depth"]).double(), torch.from_numpy(z["target"]).double(), torch.from_numpy(gm).double()
    per = [float(ts.gradient_loss(d[:, ::2 ** s, ::2 ** s], y[:, ::2 ** s, ::2 ** s], m[:, ::2 ** s, ::2 ** s])) for s in range(n_scales)]
    np.testing.assert_allclose(per, z["f64.per_scale"], rtol=1e-12, atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("name", DEPTHGRAD_GOLDEN)
def test_gpu_depth_gradient_loss_matches_reference_golden(golden_dir, name):
    from gscream_b200 import losses
    z, fit, l1m, gm, n_scales, aligned = _load_depthgrad(golden_dir, name)
    dev = torch.device("cuda")
    t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    d = torch.from_numpy(z["depth"]).to(dev).requires_grad_(True)
    y = t(z["target"])
    spread = float(np.abs(z["g_gl"] - z["f64.g_gl"]).max()) + float(np.abs(z["g_l1"] - z["f64.g_l1"]).max())
    if aligned:
        l1, gl = losses.aligned_depth_losses(d, y, t(fit), t(l1m), t(gm), n_scales=n_scales)
        assert abs(float(l1) - float(z["f64.l1"])) <= 1e-5 * abs(float(z["f64.l1"]))
        assert abs(float(gl) - float(z["f64.gl"])) <= 1e-5 * abs(float(z["f64.gl"]))
        g, = torch.autograd.grad(0.7 * l1 + 1.9 * gl, d)
        ref = 0.7 * z["f64.g_l1"] + 1.9 * z["f64.g_gl"]
        _close(g.cpu().numpy().astype(np.float64), ref, 1e-5, "d (l1, grad loss) / d depth", extra=6.0 * spread)
        # the L1 half alone is the older single-term node
        d2 = torch.from_numpy(z["depth"]).to(dev).requires_grad_(True)
        l1b = losses.aligned_depth_l1(d2, y, t(fit), t(l1m))
        assert abs(float(l1b) - float(l1)) <= 1e-6 * abs(float(l1))
    else:
        gl = losses.multiscale_gradient_loss(d, y, t(gm), n_scales=n_scales)
        assert abs(float(gl) - float(z["f64.gl"])) <= 1e-5 * abs(float(z["f64.gl"]))
        g, = torch.autograd.grad(1.9 * gl, d)
        _close(g.cpu().numpy().astype(np.float64), 1.9 * z["f64.g_gl"], 1e-5, "d grad loss / d prediction", extra=6.0 * spread)
        # the reference's call pattern: one gradient_loss per strided view (train.py:558-560)
        d3 = torch.from_numpy(z["depth"]).to(dev)
        per = [float(losses.gradient_loss(d3[:, ::2 ** s, ::2 ** s], y[:, ::2 ** s, ::2 ** s], t(gm)[:, ::2 ** s, ::2 ** s])) for s in range(n_scales)]
        np.testing.assert_allclose(per, z["f64.per_scale"], rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_gpu_depth_losses_train_step_size_vs_fp64_closed_forms():
    """567x1008 (the SPIN-NeRF size): CUDA vs the fp64 closed forms pinned above."""
    import _depthgrad_emul as em
    from gscream_b200 import losses
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(9)
    H, W = 567, 1008
    depth = 2.0 + 8.0 * torch.rand(1, H, W, generator=g)
    target = (0.8 * depth + 1.0 + 0.5 * torch.randn(1, H, W, generator=g)).clamp(min=0.1)
    valid = (torch.rand(1, H, W, generator=g) > 0.1).float()
    l1_64, gl_64, g64 = em.fused_depth_losses(depth.numpy(), target.numpy(), valid.numpy(), valid.numpy(), valid.numpy(), 4, True, 1.0, 0.5)
    x = depth.to(dev).requires_grad_(True)
    l1, gl = losses.aligned_depth_losses(x, target.to(dev), valid.to(dev), valid.to(dev), valid.to(dev))
    (l1 + 0.5 * gl).backward()
    assert abs(float(l1) - l1_64) <= 1e-5 * abs(l1_64) and abs(float(gl) - gl_64) <= 1e-5 * abs(gl_64)
    # sign(e_q - e_p) of a near-tie can flip in fp32: allow isolated pixels, bound everything else tightly
    err = np.abs(x.grad.double().cpu().numpy() - g64)
    scale = np.abs(g64).max()
    assert int((err > 2e-5 * scale).sum()) <= 32, int((err > 2e-5 * scale).sum())


def test_depth_losses_validate_arguments_and_have_no_cpu_path():
    from gscream_b200 import losses
    d, y = torch.rand(1, 8, 9), torch.rand(1, 8, 9)
    for fn in (lambda: losses.aligned_depth_losses(d, y), lambda: losses.multiscale_gradient_loss(d, y), lambda: losses.gradient_loss(d, y, torch.ones(1, 8, 9)),
               lambda: losses.aligned_depth_l1(d, y)):
        with pytest.raises(TypeError, match="no CPU loss path"):
            fn()
    with pytest.raises(NotImplementedError):
        losses.aligned_depth_losses(d, y.requires_grad_(True))
    # the reference signature: gradient_loss(prediction, target, mask[, reduction]) — train.py:232
    import inspect
    assert list(inspect.signature(losses.gradient_loss).parameters) == ["prediction", "target", "mask"]
    assert list(inspect.signature(losses.aligned_depth_losses).parameters) == ["depth", "target", "fit_mask", "l1_mask", "grad_mask", "n_scales"]
