"""Fused photometric losses: drop-ins for `l1_loss`, `l1_loss_masked`, `ssim`, `ssim_masked` of W-Ted/GScream's
utils/loss_utils.py (:27-31, :131-207), SURVEY.md section 8f rank 3.

Same names, argument order and return values (0-d tensors) as the reference functions, so train.py:535-545 can import them
from here instead; `l1_ssim(image, gt, mask)` returns both means from ONE forward / ONE backward kernel
(`gsr_l1_ssim_forward/backward` of include/gsr_b200.h), which is what the train-step loop of bench.py uses;
`aligned_depth_l1(depth, target, fit_mask, loss_mask)` is the scale/shift-aligned depth L1 of train.py:548-569;
`aligned_depth_losses(...)` adds the four-scale `gradient_loss` of train.py:232-251, :556-560 from the same fit, and
`gradient_loss` / `multiscale_gradient_loss` are that term alone on an arbitrary prediction.
Only the rendered image receives a gradient (the target and the mask are data).  No CPU / eager fallback: CPU tensors raise.
"""
from math import exp

import numpy as np
import torch

from . import _lib


def _taps(window_size=11, sigma=1.5):
    """The reference's normalised 1-D window, computed the way it computes it (loss_utils.py:112-114: Python-float exp,
    fp32 tensor, fp32 sum); its 2-D window is the outer product of this vector (loss_utils.py:116-121)."""
    g = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return np.ascontiguousarray((g / g.sum()).numpy(), dtype=np.float32)


_TAPS11 = _taps()


def _planes(t, name):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError("%s must be a float32 CUDA tensor (got %s on %s); there is no CPU loss path" % (name, t.dtype, t.device))
    if t.dim() < 2:
        raise ValueError("%s must be [..., H, W]" % name)
    t = t.contiguous()
    H, W = t.shape[-2], t.shape[-1]
    return t, int(t.numel() // (H * W)), int(H), int(W)


class _L1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, target, mask):
        lib = _lib.load()
        x, planes, H, W = _planes(image, "image")
        y, planes_y, Hy, Wy = _planes(target, "target")
        if (planes_y, Hy, Wy) != (planes, H, W):
            raise ValueError("image and target shapes differ: %s vs %s" % (tuple(image.shape), tuple(target.shape)))
        m, mp = None, 0
        if mask is not None:
            m, mp, Hm, Wm = _planes(mask.to(torch.float32) if mask.dtype != torch.float32 else mask, "mask")
            if (Hm, Wm) != (H, W) or mp not in (1, planes):
                raise ValueError("mask must broadcast over the image planes: %s vs %s" % (tuple(mask.shape), tuple(image.shape)))
        dev = x.device
        sums = torch.empty(2, dtype=torch.float64, device=dev)
        need_grad = image.requires_grad
        partials = torch.empty((3, planes, H, W), dtype=torch.float32, device=dev) if need_grad else None
        with torch.cuda.device(dev):
            _lib.check(lib.gsr_l1_ssim_forward(planes, H, W, _TAPS11.ctypes.data, x.data_ptr(), y.data_ptr(),
                                               m.data_ptr() if m is not None else None, mp, sums.data_ptr(),
                                               partials.data_ptr() if partials is not None else None,
                                               torch.cuda.current_stream().cuda_stream))
        means = (sums / float(planes * H * W)).to(torch.float32)
        ctx.save_for_backward(x, y, m if m is not None else torch.empty(0, device=dev), partials if partials is not None else torch.empty(0, device=dev))
        ctx.dims = (planes, H, W, mp, tuple(image.shape))
        return means[0], means[1]   # mean SSIM, mean L1

    @staticmethod
    def backward(ctx, g_ssim, g_l1):
        lib = _lib.load()
        x, y, m, partials = ctx.saved_tensors
        planes, H, W, mp, shape = ctx.dims
        if partials.numel() == 0:
            raise RuntimeError("l1_ssim backward without a forward that saved its partials")
        dev = x.device
        ups = torch.stack([g_ssim.to(torch.float32).reshape(()), g_l1.to(torch.float32).reshape(())]).contiguous()
        grad = torch.empty((planes, H, W), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.gsr_l1_ssim_backward(planes, H, W, _TAPS11.ctypes.data, x.data_ptr(), y.data_ptr(),
                                                m.data_ptr() if m.numel() else None, mp, partials.data_ptr(), ups.data_ptr(),
                                                grad.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return grad.view(shape), None, None


class _AlignedDepthL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, target, fit_mask, loss_mask):
        lib = _lib.load()
        d, B, H, W = _planes(depth, "depth")
        y, By, Hy, Wy = _planes(target, "target")
        if (By, Hy, Wy) != (B, H, W):
            raise ValueError("depth and target shapes differ: %s vs %s" % (tuple(depth.shape), tuple(target.shape)))
        masks = []
        for m, name in ((fit_mask, "fit_mask"), (loss_mask, "loss_mask")):
            if m is None:
                masks.append(None)
                continue
            mm, Bm, Hm, Wm = _planes(m.to(torch.float32) if m.dtype != torch.float32 else m, name)
            if (Bm, Hm, Wm) != (B, H, W):
                raise ValueError("%s must have the shape of depth" % name)
            masks.append(mm)
        dev = d.device
        state = torch.empty(1 + 7 * B, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.gsr_depth_align_l1_forward(B, H, W, d.data_ptr(), y.data_ptr(), masks[0].data_ptr() if masks[0] is not None else None,
                                                      masks[1].data_ptr() if masks[1] is not None else None, state.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream))
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(d, y, masks[0] if masks[0] is not None else empty, masks[1] if masks[1] is not None else empty, state)
        ctx.dims = (B, H, W, tuple(depth.shape))
        return (state[0] / float(B * H * W)).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        d, y, fm, lm, state = ctx.saved_tensors
        B, H, W, shape = ctx.dims
        up = g.to(torch.float32).reshape(1).contiguous()
        grad = torch.empty((B, H, W), dtype=torch.float32, device=d.device)
        with torch.cuda.device(d.device):
            _lib.check(lib.gsr_depth_align_l1_backward(B, H, W, d.data_ptr(), y.data_ptr(), fm.data_ptr() if fm.numel() else None,
                                                       lm.data_ptr() if lm.numel() else None, state.data_ptr(), up.data_ptr(), grad.data_ptr(),
                                                       torch.cuda.current_stream().cuda_stream))
        return grad.view(shape), None, None, None


def _opt_mask(m, name, B, H, W):
    if m is None:
        return None
    mm, Bm, Hm, Wm = _planes(m.to(torch.float32) if m.dtype != torch.float32 else m, name)
    if (Bm, Hm, Wm) != (B, H, W):
        raise ValueError("%s must have the shape of the depth map" % name)
    return mm


def _ptr(t):
    return t.data_ptr() if t is not None and t.numel() else None


class _AlignedDepthLosses(torch.autograd.Function):
    """(aligned L1, multi-scale gradient loss) of one rendered depth map from one scale/shift fit.  `align=False`: no fit, the
    prediction is used as it is and only the gradient term is produced (plain `gradient_loss`)."""

    @staticmethod
    def forward(ctx, depth, target, fit_mask, l1_mask, grad_mask, n_scales, align):
        lib = _lib.load()
        d, B, H, W = _planes(depth, "depth")
        y, By, Hy, Wy = _planes(target, "target")
        if (By, Hy, Wy) != (B, H, W):
            raise ValueError("depth and target shapes differ: %s vs %s" % (tuple(depth.shape), tuple(target.shape)))
        if not 1 <= int(n_scales) <= 4:
            raise ValueError("n_scales must be in 1..4 (train.py uses 4)")
        fm, lm, gm = _opt_mask(fit_mask, "fit_mask", B, H, W), _opt_mask(l1_mask, "l1_mask", B, H, W), _opt_mask(grad_mask, "grad_mask", B, H, W)
        dev = d.device
        state = torch.empty(1 + 7 * B, dtype=torch.float64, device=dev) if align else None
        gstate = torch.empty(1 + 16 * B, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            if align:
                _lib.check(lib.gsr_depth_align_l1_forward(B, H, W, d.data_ptr(), y.data_ptr(), _ptr(fm), _ptr(lm), state.data_ptr(), stream))
            _lib.check(lib.gsr_depth_grad_forward(B, H, W, int(n_scales), d.data_ptr(), y.data_ptr(), _ptr(gm),
                                                  state.data_ptr() if align else None, gstate.data_ptr(), stream))
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(d, y, fm if fm is not None else empty, lm if lm is not None else empty, gm if gm is not None else empty,
                              state if align else empty, gstate)
        ctx.dims = (B, H, W, int(n_scales), bool(align), tuple(depth.shape))
        l1 = (state[0] / float(B * H * W)).to(torch.float32) if align else torch.zeros((), dtype=torch.float32, device=dev)
        return l1, gstate[0].to(torch.float32)

    @staticmethod
    def backward(ctx, g_l1, g_grad):
        lib = _lib.load()
        d, y, fm, lm, gm, state, gstate = ctx.saved_tensors
        B, H, W, n_scales, align, shape = ctx.dims
        dev = d.device
        grad = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        up_g = g_grad.to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            if align:
                up_l = g_l1.to(torch.float32).reshape(1).contiguous()
                _lib.check(lib.gsr_depth_align_l1_backward(B, H, W, d.data_ptr(), y.data_ptr(), _ptr(fm), _ptr(lm), state.data_ptr(),
                                                           up_l.data_ptr(), grad.data_ptr(), stream))
            _lib.check(lib.gsr_depth_grad_backward(B, H, W, n_scales, d.data_ptr(), y.data_ptr(), _ptr(gm), _ptr(fm),
                                                   state.data_ptr() if align else None, gstate.data_ptr(), up_g.data_ptr(), grad.data_ptr(),
                                                   1 if align else 0, stream))
        return grad.view(shape), None, None, None, None, None, None


def aligned_depth_losses(depth, target, fit_mask=None, l1_mask=None, grad_mask=None, n_scales=4):
    """The whole depth branch of train.py:546-560 (reference view) / :563-574 (other views) in one node:

        scale, shift = compute_scale_and_shift(depth, target, fit_mask); scale = torch.abs(scale)
        aligned = scale * depth + shift
        l1   = l1_loss(aligned, target)  (l1_mask None)  or  l1_loss_masked(aligned, target, l1_mask)
        grad = sum(gradient_loss(aligned[:, ::2**s, ::2**s], target[:, ::2**s, ::2**s], grad_mask[:, ::2**s, ::2**s]) for s in range(n_scales))

    Returns (l1, grad); the caller weights them (`refer_depth_lr * l1 + 0.5 * refer_depth_lr_smooth * grad`).  grad_mask None = ones
    (train.py:560 passes torch.ones_like(gt_mask)).  The gradient w.r.t. `depth` flows through the closed-form fit for both terms."""
    if target.requires_grad:
        raise NotImplementedError("fused aligned_depth_losses differentiates w.r.t. the rendered depth only")
    return _AlignedDepthLosses.apply(depth, target, fit_mask, l1_mask, grad_mask, n_scales, True)


def multiscale_gradient_loss(prediction, target, mask=None, n_scales=4):
    """sum over s < n_scales of gradient_loss(prediction[:, ::2**s, ::2**s], ...) for an arbitrary prediction (no alignment)."""
    if target.requires_grad:
        raise NotImplementedError("fused gradient_loss differentiates w.r.t. the prediction only")
    return _AlignedDepthLosses.apply(prediction, target, None, None, mask, n_scales, False)[1]


def gradient_loss(prediction, target, mask):
    """train.py:232-251 (with its default reduction_image_based): same arguments, same value; strided views
    (`aligned_depth[:, ::step, ::step]`) are made contiguous first — use aligned_depth_losses / multiscale_gradient_loss to
    get all four scales from one pass."""
    return multiscale_gradient_loss(prediction, target, mask, n_scales=1)


def aligned_depth_l1(depth, target, fit_mask=None, loss_mask=None):
    """train.py:548-555 / 563-569 in one node: `scale, shift = compute_scale_and_shift(depth, target, fit_mask)`
    (utils/loss_utils.py:80-102), `scale = torch.abs(scale)`, `aligned = scale * depth + shift`, then `l1_loss(aligned, target)`
    (loss_mask None) or `l1_loss_masked(aligned, target, loss_mask)`; the gradient flows through the closed-form fit like the
    reference's.  depth / target / masks: [B, H, W] (B = 1 in train.py)."""
    if target.requires_grad:
        raise NotImplementedError("fused aligned_depth_l1 differentiates w.r.t. the rendered depth only")
    return _AlignedDepthL1.apply(depth, target, fit_mask, loss_mask)


def l1_ssim(image, gt, mask=None):
    """(mean SSIM, mean L1) of `image` against `gt`, both optionally weighted by `mask` exactly like ssim_masked /
    l1_loss_masked (weights multiply the per-element maps, the mean still divides by the element count)."""
    if gt.requires_grad or (mask is not None and mask.requires_grad):
        raise NotImplementedError("fused l1_ssim differentiates w.r.t. the rendered image only")
    return _L1SSIM.apply(image, gt, mask)


def l1_loss(network_output, gt):
    """utils/loss_utils.py:27-28."""
    return l1_ssim(network_output, gt)[1]


def l1_loss_masked(network_output, gt, mask):
    """utils/loss_utils.py:30-31."""
    return l1_ssim(network_output, gt, mask)[1]


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:131-164."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("fused ssim supports window_size=11, size_average=True (the values train.py uses)")
    return l1_ssim(img1, img2)[0]


def ssim_masked(img1, img2, mask, window_size=11, size_average=True):
    """utils/loss_utils.py:166-207."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("fused ssim_masked supports window_size=11, size_average=True (the values train.py uses)")
    return l1_ssim(img1, img2, mask)[0]
