"""Generates tests/golden/depthgrad_*.npz by running the REFERENCE's own depth branch — compute_scale_and_shift,
gradient_loss, reduction_image_based (/root/reference/train.py:198-251) and l1_loss / l1_loss_masked
(/root/reference/utils/loss_utils.py:27-31) composed exactly as train.py:546-560 (reference view) and :563-574 (other views)
compose them — on CPU torch, fp32 and fp64.

train.py imports the whole project at its top, so only the needed function nodes are taken from the files' ASTs and executed
where they lie; nothing is copied into the repo.

    python tests/golden/make_depthgrad_golden.py          # needs /root/reference; CPU only
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = {"/root/reference/train.py": ("compute_scale_and_shift", "reduction_image_based", "gradient_loss"),
           "/root/reference/utils/loss_utils.py": ("l1_loss", "l1_loss_masked")}


def reference_functions():
    ns = {"torch": torch}
    for path, names in SOURCES.items():
        tree = ast.parse(open(path).read())
        body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
        assert len(body) == len(names), path
        mod = ast.Module(body=body, type_ignores=[])
        ast.fix_missing_locations(mod)
        exec(compile(mod, path, "exec"), ns)
    return ns


def make_inputs(B, H, W, seed, masked_fit, grad_mask_kind):
    g = torch.Generator().manual_seed(seed)
    depth = 2.0 + 6.0 * torch.rand(B, H, W, generator=g)
    # smooth-ish target with noise: first differences of (aligned - target) take both signs everywhere
    target = (0.7 * depth + 1.3 + 0.4 * torch.randn(B, H, W, generator=g)).clamp(min=0.1)
    fit = (torch.rand(B, H, W, generator=g) > 0.25).float() if masked_fit else None
    if grad_mask_kind == "ones":
        gm = torch.ones(B, H, W)
    elif grad_mask_kind == "fit":
        gm = fit.clone()
    else:   # "holes": binary mask with one fully masked image row band (M of coarse grids shrinks)
        gm = (torch.rand(B, H, W, generator=g) > 0.4).float()
        gm[:, : max(1, H // 5)] = 0.0
    return depth, target, fit, gm


def run_aligned(B, H, W, seed, masked_fit, masked_l1, grad_mask_kind, dtype, n_scales=4):
    """train.py:546-560 / :563-574."""
    ns = reference_functions()
    depth, target, fit, gm = make_inputs(B, H, W, seed, masked_fit, grad_mask_kind)
    x = depth.to(dtype).clone().requires_grad_(True)
    y = target.to(dtype)
    fm = torch.ones_like(y) if fit is None else fit.to(dtype)
    scale, shift = ns["compute_scale_and_shift"](x, y, fm)
    scale = torch.abs(scale)
    aligned = scale.view(-1, 1, 1) * x + shift.view(-1, 1, 1)
    l1 = ns["l1_loss_masked"](aligned, y, fm) if masked_l1 else ns["l1_loss"](aligned, y)
    per_scale = []
    for s in range(n_scales):
        step = pow(2, s)
        per_scale.append(ns["gradient_loss"](aligned[:, ::step, ::step], y[:, ::step, ::step], gm.to(dtype)[:, ::step, ::step]))
    gl = sum(per_scale)
    g_l1, = torch.autograd.grad(l1, x, retain_graph=True)
    g_gl, = torch.autograd.grad(gl, x)
    return dict(depth=depth.numpy(), target=target.numpy(), fit=np.zeros(0, np.float32) if fit is None else fit.numpy(),
                grad_mask=gm.numpy(), masked_l1=int(masked_l1), aligned=1, n_scales=n_scales,
                l1=l1.detach().numpy(), gl=gl.detach().numpy(), per_scale=np.array([float(v) for v in per_scale]),
                g_l1=g_l1.numpy(), g_gl=g_gl.numpy())


def run_plain(B, H, W, seed, grad_mask_kind, dtype, n_scales=4):
    """gradient_loss on an arbitrary prediction (no alignment), all scales."""
    ns = reference_functions()
    depth, target, _, gm = make_inputs(B, H, W, seed, False, grad_mask_kind)
    x = depth.to(dtype).clone().requires_grad_(True)
    y = target.to(dtype)
    per_scale = []
    for s in range(n_scales):
        step = pow(2, s)
        per_scale.append(ns["gradient_loss"](x[:, ::step, ::step], y[:, ::step, ::step], gm.to(dtype)[:, ::step, ::step]))
    gl = sum(per_scale)
    g_gl, = torch.autograd.grad(gl, x)
    return dict(depth=depth.numpy(), target=target.numpy(), fit=np.zeros(0, np.float32), grad_mask=gm.numpy(), masked_l1=0, aligned=0,
                n_scales=n_scales, l1=np.zeros(()), gl=gl.detach().numpy(), per_scale=np.array([float(v) for v in per_scale]),
                g_l1=np.zeros_like(depth.numpy()), g_gl=g_gl.numpy())


CASES = {
    # name: (kind, B, H, W, seed, masked_fit, masked_l1, grad mask, n_scales)
    "depthgrad_ref_view": ("aligned", 1, 61, 83, 31, True, False, "ones", 4),      # train.py:546-560
    "depthgrad_other_view": ("aligned", 1, 40, 56, 32, True, True, "fit", 4),      # train.py:563-574
    "depthgrad_b2_holes": ("aligned", 2, 33, 47, 33, False, False, "holes", 4),
    "depthgrad_plain_b2": ("plain", 2, 19, 26, 34, False, False, "holes", 4),
    "depthgrad_tiny": ("plain", 1, 5, 3, 35, False, False, "ones", 4),            # stride 4 / 8 grids are 2x1 and 1x1
    "depthgrad_one_scale": ("plain", 1, 17, 23, 36, False, False, "holes", 1),
}

if __name__ == "__main__":
    for name, (kind, B, H, W, seed, mf, ml, gmk, ns_) in CASES.items():
        if kind == "aligned":
            r32, r64 = (run_aligned(B, H, W, seed, mf, ml, gmk, dt, ns_) for dt in (torch.float32, torch.float64))
        else:
            r32, r64 = (run_plain(B, H, W, seed, gmk, dt, ns_) for dt in (torch.float32, torch.float64))
        out = dict(r32)
        for k in ("l1", "gl", "per_scale", "g_l1", "g_gl"):
            out["f64." + k] = r64[k]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "l1 %.6f gl %.6f per scale %s | fp32-vs-fp64 grad diff %.2e" % (
            float(r32["l1"]), float(r32["gl"]), np.round(r64["per_scale"], 5), float(np.abs(r32["g_gl"] - r64["g_gl"]).max())))
