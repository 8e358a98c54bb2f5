"""The compiled torch glue (gscream_b200/csrc/gsr_torch_glue.cpp, loaded by gscream_b200/_glue.py): the pybind11 counterpart of
the reference's `diff_gaussian_rasterization._C` (ext.cpp:16-20, rasterize_points.cu:35-373) above the C ABI, selected with
GSR_GLUE=cpp.

CPU: it builds, exports the reference's five names, returns the reference's tuples on the P == 0 paths (no CUDA call is made
there, rasterize_points.cu:85,172) and raises the reference's errors.  GPU: the same scenes through both glues — integer outputs
and the forward planes bit-identical (same library calls), gradients within the run-to-run spread of the float atomics.
"""
import numpy as np
import pytest
import torch

E = torch.empty(0)
NAMES = ("rasterize_gaussians", "rasterize_gaussians_backward", "rasterize_aussians_filter", "rasterize_aussians_filter_position2D",
         "mark_visible")


@pytest.fixture(scope="module")
def glue():
    from gscream_b200 import _glue
    _glue.build()
    return _glue.load()


def test_glue_exports_the_reference_names(glue):
    assert sorted(n for n in dir(glue) if not n.startswith("_")) == sorted(NAMES)   # ext.cpp:16-20, typos included


def test_glue_empty_input_tuples_match_reference_layout(glue):
    H, W = 8, 12
    out = glue.rasterize_gaussians(torch.zeros(3), torch.zeros(0, 3), E, E, E, E, E, 1.0, E, torch.eye(4), torch.eye(4), 0.5, 0.5, H, W, E, 1,
                                   torch.zeros(3), False, False)
    assert out[0] == 0 and len(out) == 8                                               # rasterize_points.cu:121
    assert [tuple(t.shape) for t in out[1:5]] == [(3, H, W), (1, H, W), (1, H, W), (0,)]
    assert out[4].dtype == torch.int32 and all(t.dtype == torch.uint8 and t.numel() == 0 for t in out[5:])
    u8 = torch.empty(0, dtype=torch.uint8)
    g = glue.rasterize_gaussians_backward(torch.zeros(3), torch.zeros(0, 3), torch.zeros(0, dtype=torch.int32), E, E, E, 1.0, E, torch.eye(4),
                                          torch.eye(4), 0.5, 0.5, torch.zeros(3, H, W), torch.zeros(1, H, W), torch.zeros(1, H, W), E, 1,
                                          torch.zeros(3), u8, 0, u8, u8, False)
    # (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_duncertainty, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations), :210
    assert [tuple(t.shape) for t in g] == [(0, 3), (0, 3), (0, 1), (0, 1), (0, 3), (0, 6), (0, 0, 3), (0, 3), (0, 4)]
    assert glue.mark_visible(torch.zeros(0, 3), torch.eye(4), torch.eye(4)).dtype == torch.bool
    r, x, y = glue.rasterize_aussians_filter_position2D(torch.zeros(0, 3), E, E, 1.0, E, torch.eye(4), torch.eye(4), 0.5, 0.5, H, W, False, False)
    assert r.dtype == torch.int32 and x.dtype == torch.float32 and y.shape == (0,)


def test_glue_errors_match_reference_behaviour(glue):
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):           # rasterize_points.cu:58-60
        glue.rasterize_aussians_filter(torch.zeros(4, 2), E, E, 1.0, E, torch.eye(4), torch.eye(4), 0.5, 0.5, 16, 16, False, False)
    with pytest.raises(ValueError, match="no CPU rasterizer"):
        glue.rasterize_gaussians(torch.zeros(3), torch.zeros(5, 3), E, E, E, E, E, 1.0, E, torch.eye(4), torch.eye(4), 0.5, 0.5, 8, 8, E, 1,
                                 torch.zeros(3), False, False)


def test_C_routes_through_the_glue_when_selected(glue, monkeypatch):
    from gscream_b200 import _C
    monkeypatch.setenv("GSR_GLUE", "cpp")
    assert _C._compiled() is glue
    out = _C.rasterize_gaussians(torch.zeros(3), torch.zeros(0, 3), E, E, E, E, E, 1.0, E, torch.eye(4), torch.eye(4), 0.5, 0.5, 8, 8, E, 1,
                                 torch.zeros(3), False, False)
    assert out[0] == 0 and len(out) == 8
    monkeypatch.setenv("GSR_GLUE", "ctypes")
    assert _C._compiled() is None


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c32_small", "c3_ragged"])
def test_gpu_compiled_glue_equals_ctypes_glue(glue, monkeypatch, name):
    import _ref_utils as ru
    from _cases import GRAD_KEYS, load_golden
    from gscream_b200 import rasterizer as ours
    import test_gpu_parity as tg
    g = load_golden(name)
    scene, cam, grads = tg._scene_from_golden(g)
    runs = {}
    for which in ("ctypes", "cpp", "ctypes2"):
        monkeypatch.setenv("GSR_GLUE", "cpp" if which == "cpp" else "ctypes")
        runs[which] = ru.run_impl(ours, scene, cam, grads)
    a, b, a2 = runs["ctypes"], runs["cpp"], runs["ctypes2"]
    assert b["num_rendered"] == a["num_rendered"] == int(g["num_rendered"]) and np.array_equal(b["radii"], a["radii"])
    for k in ("color", "depth", "uncertainty"):
        assert np.array_equal(b[k], a[k]), k                                          # the forward is deterministic
    for k in GRAD_KEYS:
        spread = float(np.abs(a2[k] - a[k]).max())
        assert np.abs(b[k] - a[k]).max() <= 8.0 * spread + 1e-6 * np.abs(a[k]).max(), k
    # filters and markVisible
    dev = torch.device("cuda")
    rs = dict(image_height=cam["H"], image_width=cam["W"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=scene["bg"].to(dev), scale_modifier=1.0,
              viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=1, campos=cam["campos"].to(dev),
              prefiltered=False, debug=False)
    rast = ours.GaussianRasterizer(raster_settings=ours.GaussianRasterizationSettings(**rs))
    m3, sc, ro = scene["means3D"].to(dev), scene["scales"].to(dev), scene["rotations"].to(dev)
    res = {}
    for which in ("ctypes", "cpp"):
        monkeypatch.setenv("GSR_GLUE", which)
        res[which] = (rast.visible_filter(m3, sc, ro), *rast.position2D_filter(m3, sc, ro), rast.markVisible(m3))
    for x, y in zip(res["ctypes"], res["cpp"]):
        assert torch.equal(x, y)
