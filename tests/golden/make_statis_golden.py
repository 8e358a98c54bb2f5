"""Generates tests/golden/statis_*.npz by running the REFERENCE's own GaussianModel.training_statis
(/root/reference/scene/gaussian_model.py:729-757) on CPU torch: the method's AST node is executed from the file where it lies
(the module itself imports packages that are absent here), bound to a bare object that carries the four statistics buffers.

    python tests/golden/make_statis_golden.py          # needs /root/reference; CPU only
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/scene/gaussian_model.py"


def reference_method():
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GaussianModel")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "training_statis")
    mod = ast.Module(body=[fn], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch}
    exec(compile(mod, REF, "exec"), ns)
    return ns["training_statis"]


class Model:
    pass


def make_case(A, k, seed, vis_frac):
    g = torch.Generator().manual_seed(seed)
    m = Model()
    m.n_offsets = k
    m.opacity_accum = torch.rand(A, 1, generator=g)
    m.anchor_demon = torch.randint(0, 5, (A, 1), generator=g).float()
    m.offset_gradient_accum = torch.rand(A * k, 1, generator=g)
    m.offset_denom = torch.randint(0, 5, (A * k, 1), generator=g).float()
    vis = torch.rand(A, generator=g) < vis_frac
    n_vis = int(vis.sum())
    opacity = torch.tanh(torch.randn(n_vis * k, 1, generator=g))
    sel = (opacity > 0).view(-1)
    P = int(sel.sum())
    upd = torch.rand(P, generator=g) < 0.7
    pts = torch.zeros(P, 3, requires_grad=True)
    pts.grad = torch.randn(P, 3, generator=g) * 1e-3
    return m, pts, opacity, upd, sel, vis


if __name__ == "__main__":
    fn = reference_method()
    for name, (A, k, seed, vf) in {"statis_k10": (500, 10, 1, 0.6), "statis_k4_sparse": (333, 4, 2, 0.1), "statis_all": (64, 10, 3, 1.1)}.items():
        m, pts, opacity, upd, sel, vis = make_case(A, k, seed, vf)
        before = {n: getattr(m, n).clone().numpy() for n in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom")}
        fn(m, pts, opacity, upd, sel, vis)
        out = {"A": A, "k": k, "vis": vis.numpy(), "opacity": opacity.numpy(), "sel": sel.numpy(), "upd": upd.numpy(), "grad": pts.grad.numpy()}
        for n, v in before.items():
            out["in." + n] = v
            out["out." + n] = getattr(m, n).numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "n_vis", int(vis.sum()), "P", int(sel.sum()), "updated offsets", int((out["out.offset_denom"] != out["in.offset_denom"]).sum()))
