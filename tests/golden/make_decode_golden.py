"""Generates tests/golden/decode_*.npz by running the REFERENCE's own `generate_neural_gaussians`
(/root/reference/gaussian_renderer/__init__.py:18-102) on CPU torch.

The reference module cannot be imported as a whole here (it pulls in scene.gaussian_model -> StyleGAN / attention
packages that are not installed, and a CUDA-only constructor), so this script parses the file, takes that ONE function's
AST node and executes it — from the file where it lies, nothing is copied into the repo — against a synthetic anchor
model that exposes the attribute names the function uses (tests/_anchor_decode.py:SyntheticAnchors).
Stored: model parameters, camera centre, visible mask, all eight outputs, and the gradients of a fixed scalar
functional of the outputs w.r.t. every parameter (fp32 and fp64 runs).

    python tests/golden/make_decode_golden.py          # needs /root/reference; CPU only
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _anchor_decode as ad  # noqa: E402

REF = "/root/reference/gaussian_renderer/__init__.py"


def reference_function():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "generate_neural_gaussians")
    fn.args.args[1].annotation = None  # `pc : GaussianModel` — the class is not importable here
    mod = ast.Module(body=[fn], type_ignores=[])
    ast.fix_missing_locations(mod)
    from einops import repeat
    ns = {"torch": torch, "repeat": repeat}
    exec(compile(mod, REF, "exec"), ns)
    return ns["generate_neural_gaussians"]


class Cam:
    def __init__(self, c):
        self.camera_center = c


def upstream(outs, seed):
    """Fixed random upstream gradients for the seven differentiable outputs."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(o.shape, generator=g, dtype=torch.float32).to(o.dtype) for o in outs[:7]]


def run_case(name, A, k, seed, vis_frac, dtype):
    pc = ad.SyntheticAnchors(A, n_offsets=k, seed=seed).to(dtype)
    g = torch.Generator().manual_seed(seed + 1000)
    vis = None if vis_frac is None else (torch.rand(A, generator=g) < vis_frac)
    cam = Cam(torch.tensor([0.1, -0.2, 0.3], dtype=dtype))
    fn = reference_function()
    outs = fn(cam, pc, vis, is_training=True)
    ups = upstream(outs, seed + 2000)
    loss = sum((o * u).sum() for o, u in zip(outs[:7], ups))
    params = dict(pc.named_parameters())
    grads = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
    rec = {"A": A, "k": k, "campos": cam.camera_center.numpy(), "vis": np.zeros(0, bool) if vis is None else vis.numpy()}
    for n, p in params.items():
        rec["p." + n] = p.detach().numpy()
    names = ("xyz", "color", "opacity", "uncertainty", "scaling", "rot", "neural_opacity", "mask")
    for n, o in zip(names, outs):
        rec["o." + n] = o.detach().numpy()
    for n, u in zip(names, ups):
        rec["u." + n] = u.numpy()
    for (n, p), gr in zip(params.items(), grads):
        rec["g." + n] = np.zeros_like(p.detach().numpy()) if gr is None else gr.numpy()
    return rec


if __name__ == "__main__":
    cases = {"decode_k10_vis": (301, 10, 11, 0.6), "decode_k10_all": (128, 10, 12, None), "decode_k5_vis": (77, 5, 13, 0.5)}
    for name, (A, k, seed, vf) in cases.items():
        r32 = run_case(name, A, k, seed, vf, torch.float32)
        r64 = run_case(name, A, k, seed, vf, torch.float64)
        out = {k_: v for k_, v in r32.items()}
        for k_, v in r64.items():
            if k_.startswith(("o.", "g.")):
                out["f64." + k_] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "P =", r32["o.xyz"].shape[0], "n_vis*k =", r32["o.mask"].shape[0], "mask equal f32/f64:",
              np.array_equal(r32["o.mask"], r64["o.mask"]))
