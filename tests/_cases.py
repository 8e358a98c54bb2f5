"""Shared helpers for the parity tests: golden loading and oracle invocation."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["c3_small", "c3_ragged", "c32_small", "c32_dense"]
GRAD_KEYS = ["dL_dmeans3D", "dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_duncertainty", "dL_dscales", "dL_drotations"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def oracle_forward(orc, g):
    return orc.forward(
        means3D=g["in_means3D"], colors_precomp=g["in_colors"], opacities=g["in_opacities"], uncertainties=g["in_uncertainties"],
        scales=g["in_scales"], rotations=g["in_rotations"], viewmatrix=g["cam_viewmatrix"], projmatrix=g["cam_projmatrix"],
        bg=g["in_bg"], W=int(g["W"]), H=int(g["H"]), tanfovx=float(g["cam_tanfovx"]), tanfovy=float(g["cam_tanfovy"]))


def oracle_backward(orc, fwd, g):
    return orc.backward(
        fwd, means3D=g["in_means3D"], colors_precomp=g["in_colors"], scales=g["in_scales"], rotations=g["in_rotations"],
        viewmatrix=g["cam_viewmatrix"], projmatrix=g["cam_projmatrix"], bg=g["in_bg"], W=int(g["W"]), H=int(g["H"]),
        tanfovx=float(g["cam_tanfovx"]), tanfovy=float(g["cam_tanfovy"]), dL_dcolor=g["g_color"], dL_ddepth=g["g_depth"], dL_dunc=g["g_unc"])


def rel_err(a, b, floor):
    """max |a-b| / max(|b|, floor) — the 1e-5-relative criterion of SURVEY.md section 8d with a per-plane floor."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
