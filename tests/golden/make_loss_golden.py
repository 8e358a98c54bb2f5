"""Generates tests/golden/loss_*.npz by running the REFERENCE's own l1_loss / l1_loss_masked / ssim / ssim_masked
(/root/reference/utils/loss_utils.py:27-31, 112-207) on CPU torch, fp32 and fp64.

The module imports kornia (absent here) at its top, so — as in make_decode_golden.py — only the needed function nodes
are taken from the file's AST and executed where they lie; nothing is copied into the repo.

    python tests/golden/make_loss_golden.py          # needs /root/reference; CPU only
"""
import ast
import os
from math import exp

import numpy as np
import torch
import torch.nn.functional as F
from torch.autograd import Variable

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/utils/loss_utils.py"
NAMES = ("l1_loss", "l1_loss_masked", "gaussian", "create_window", "ssim", "_ssim", "ssim_masked", "_ssim_masked", "compute_scale_and_shift")


def reference_functions():
    tree = ast.parse(open(REF).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in NAMES]
    assert len(body) == len(NAMES)
    mod = ast.Module(body=body, type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch, "F": F, "Variable": Variable, "exp": exp}
    exec(compile(mod, REF, "exec"), ns)
    return ns


def run_case(C, H, W, seed, masked, mask_planes, dtype):
    ns = reference_functions()
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(C, H, W, generator=g)
    # a target correlated with the image (SSIM in its interesting range) plus flat and saturated regions
    gt = (img + 0.15 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    gt[:, : H // 4, : W // 3] = 0.5
    img[:, H // 2:, W // 2:] = img[:, H // 2:, W // 2:].round()
    mask = (torch.rand(mask_planes, H, W, generator=g) > 0.3).float() * (0.5 + torch.rand(mask_planes, H, W, generator=g)) if masked else None
    x = img.to(dtype).clone().requires_grad_(True)
    y = gt.to(dtype)
    if masked:
        m = mask.to(dtype)
        s, l = ns["ssim_masked"](x, y, m), ns["l1_loss_masked"](x, y, m)
    else:
        s, l = ns["ssim"](x, y), ns["l1_loss"](x, y)
    gs, = torch.autograd.grad(s, x, retain_graph=True)
    gl, = torch.autograd.grad(l, x)
    return dict(img=img.numpy(), gt=gt.numpy(), mask=np.zeros(0, np.float32) if mask is None else mask.numpy(),
                ssim=s.detach().numpy(), l1=l.detach().numpy(), g_ssim=gs.numpy(), g_l1=gl.numpy())


def run_depth_case(B, H, W, seed, masked_fit, masked_loss, dtype):
    """train.py:548-555 (reference view: fit on the valid mask, unmasked L1) and :563-569 (other views: masked L1)."""
    ns = reference_functions()
    g = torch.Generator().manual_seed(seed)
    depth = 2.0 + 6.0 * torch.rand(B, H, W, generator=g)
    target = (0.7 * depth + 1.3 + 0.4 * torch.randn(B, H, W, generator=g)).clamp(min=0.1)
    fit = (torch.rand(B, H, W, generator=g) > 0.25).float() if masked_fit else None
    x = depth.to(dtype).clone().requires_grad_(True)
    y = target.to(dtype)
    fm = None if fit is None else fit.to(dtype)
    scale, shift = ns["compute_scale_and_shift"](x, y, fm)
    scale = torch.abs(scale)
    aligned = scale.view(-1, 1, 1) * x + shift.view(-1, 1, 1)
    loss = ns["l1_loss_masked"](aligned, y, fm) if masked_loss else ns["l1_loss"](aligned, y)
    gr, = torch.autograd.grad(loss, x)
    return dict(depth=depth.numpy(), target=target.numpy(), fit=np.zeros(0, np.float32) if fit is None else fit.numpy(),
                masked_loss=int(masked_loss), loss=loss.detach().numpy(), scale=scale.detach().numpy(), shift=shift.detach().numpy(), grad=gr.numpy())


if __name__ == "__main__":
    for name, (B, H, W, seed, mf, ml) in {"depth_ref_view": (1, 61, 83, 21, True, False), "depth_other_view": (1, 40, 56, 22, True, True),
                                           "depth_nomask_b2": (2, 17, 23, 23, False, False)}.items():
        r32 = run_depth_case(B, H, W, seed, mf, ml, torch.float32)
        r64 = run_depth_case(B, H, W, seed, mf, ml, torch.float64)
        out = dict(r32)
        for k in ("loss", "scale", "shift", "grad"):
            out["f64." + k] = r64[k]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "loss %.6f scale %s shift %s" % (float(r32["loss"]), r32["scale"], r32["shift"]))
    cases = {"loss_rgb": (3, 70, 101, 5, False, 0), "loss_rgb_masked1": (3, 64, 48, 6, True, 1), "loss_rgb_masked3": (3, 33, 57, 7, True, 3),
             "loss_tiny": (1, 7, 9, 8, False, 0)}
    for name, (C, H, W, seed, masked, mp) in cases.items():
        r32 = run_case(C, H, W, seed, masked, mp, torch.float32)
        r64 = run_case(C, H, W, seed, masked, mp, torch.float64)
        out = dict(r32)
        for k in ("ssim", "l1", "g_ssim", "g_l1"):
            out["f64." + k] = r64[k]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "ssim %.6f l1 %.6f" % (float(r32["ssim"]), float(r32["l1"])))
