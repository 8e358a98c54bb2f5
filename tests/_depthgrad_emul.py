"""TEST INFRASTRUCTURE — numpy fp64 restatement of the closed forms the fused depth-loss kernels evaluate
(gsr_loss.cu: depth_fit_sums / depth_l1 / depth_l1_backward / depth_grad_forward / depth_grad_backward): the five fit sums, the
aligned L1, the per-(image, scale) gradient-loss sums {sum, M, S1, S0} and the hand-derived backward that carries both terms
through the closed-form scale/shift fit.  tests/test_losses.py pins it to the reference's autograd (tests/golden/depthgrad_*.npz)
on the CPU, so the derivation the CUDA code follows is checked without a GPU; the GPU tests then compare the CUDA path with the
same vectors."""
import numpy as np


def _sgn(v):
    return np.sign(v)


def fused_depth_losses(d, y, fit_mask, l1_mask, grad_mask, n_scales, align, up_l1=1.0, up_gl=1.0):
    """d, y: [B,H,W] float64; masks [B,H,W] or None (= ones).  Returns (l1, gl, d(up_l1*l1 + up_gl*gl)/dd)."""
    d, y = d.astype(np.float64), y.astype(np.float64)
    B, H, W = d.shape
    ones = np.ones_like(d)
    fm = ones if fit_mask is None else fit_mask.astype(np.float64)
    lm = ones if l1_mask is None else l1_mask.astype(np.float64)
    gm = ones if grad_mask is None else grad_mask.astype(np.float64)
    grad = np.zeros_like(d)
    l1_total, gl_total = 0.0, 0.0
    for b in range(B):
        if align:
            a00, a01, a11 = (fm[b] * d[b] * d[b]).sum(), (fm[b] * d[b]).sum(), fm[b].sum()
            b0, b1 = (fm[b] * d[b] * y[b]).sum(), (fm[b] * y[b]).sum()
            det = a00 * a11 - a01 * a01
            x0 = (a11 * b0 - a01 * b1) / det if det != 0 else 0.0
            x1 = (-a01 * b0 + a00 * b1) / det if det != 0 else 0.0
            sc = abs(x0)
        else:
            sc, x0, x1, det = 1.0, 1.0, 0.0, 0.0
        A = sc * d[b] + x1
        dLdA = np.zeros((H, W))
        g0 = g1 = 0.0
        if align:
            r = A - y[b]
            l1_total += (np.abs(r) * lm[b]).sum()
            sg = _sgn(r) * lm[b]
            dLdA += up_l1 * sg / (B * H * W)
        e = gm[b] * (A - y[b])
        for s in range(n_scales):
            st = 1 << s
            es, ms = e[::st, ::st], gm[b][::st, ::st]
            M = ms.sum()
            dx = es[:, 1:] - es[:, :-1]
            wx = ms[:, 1:] * ms[:, :-1]
            dy = es[1:, :] - es[:-1, :]
            wy = ms[1:, :] * ms[:-1, :]
            tot = (np.abs(dx) * wx).sum() + (np.abs(dy) * wy).sum()
            gl_total += (tot / M if M != 0 else tot) / B
            coef = up_gl / B / (M if M != 0 else 1.0)
            ge = np.zeros_like(es)                   # d tot / d e on the sub-sampled grid
            ge[:, 1:] += _sgn(dx) * wx
            ge[:, :-1] -= _sgn(dx) * wx
            ge[1:, :] += _sgn(dy) * wy
            ge[:-1, :] -= _sgn(dy) * wy
            dLdA[::st, ::st] += coef * ge * ms       # de/dA = m
        grad[b] = dLdA * sc
        if align:
            g0 = np.sign(x0) * (dLdA * d[b]).sum()
            g1 = dLdA.sum()
            if det != 0:
                l0 = (a11 * g0 - a01 * g1) / det
                l1_ = (-a01 * g0 + a00 * g1) / det
                grad[b] += fm[b] * (l0 * (y[b] - 2.0 * d[b] * x0 - x1) - l1_ * x0)
    return l1_total / (B * H * W), gl_total, grad
