"""CPU: the C-ABI library loads and exports every symbol include/gsr_b200.h declares (no compute calls: there is
no GPU here), argument validation that needs no device, and the host-side mirror of the reference's Python surface."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gscream_b200 import _build, _lib
    _build.build()  # nvcc cross-compiles sm_100a without a GPU
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gsr_b200.h")).read()
    return sorted(set(re.findall(r"GSR_API\s+[\w\s\*]+?\b(gsr_\w+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 14
    raw = ctypes.CDLL(lib._name)
    for n in names:
        assert hasattr(raw, n), "missing export: " + n
    from gscream_b200._lib import PROTOTYPES
    assert sorted(PROTOTYPES) == names  # the ctypes table binds exactly the declared ABI


def test_library_is_sm100a_only():
    from gscream_b200 import _build
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % _build.LIB).read()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_capability_and_size_queries(lib):
    assert lib.gsr_abi_version() == 2
    assert lib.gsr_supported_channels(3) == 1 and lib.gsr_supported_channels(32) == 1
    assert lib.gsr_supported_channels(5) == 0
    g1, g2 = lib.gsr_geom_bytes(1000), lib.gsr_geom_bytes(2000)
    assert 0 < g1 < g2 and g1 >= 1000 * 64
    assert lib.gsr_image_bytes(1920, 1080) >= 1920 * 1080 * 8 + 120 * 68 * 8
    b1, b2 = lib.gsr_binning_bytes(1000, 5000, 640, 480), lib.gsr_binning_bytes(1000, 50000, 640, 480)
    assert 0 < b1 < b2 and b2 >= 50000 * 16
    assert b"success" in lib.gsr_error_string(0)
    assert b"precomputed" in lib.gsr_error_string(-4)  # the reference's runtime_error text (rasterizer_impl.cu:248)


def test_argument_validation_without_device(lib):
    n = ctypes.c_int64(-1)
    # P == 0 short-circuits like rasterize_points.cu:85 and reports zero instances
    rc = lib.gsr_forward_stage1(0, 3, 0, 0, None, None, None, None, None, None, 1.0, None, None, None, None, None,
                                64, 64, 0.5, 0.5, 0, None, None, 0, ctypes.addressof(n), None)
    assert rc == 0 and n.value == 0
    rc = lib.gsr_forward_stage1(10, 3, 0, 0, None, None, None, None, None, None, 1.0, None, None, None, None, None,
                                64, 64, 0.5, 0.5, 0, None, None, 0, ctypes.addressof(n), None)
    assert rc == -1  # null inputs
    rc = lib.gsr_forward_stage1(10, 3, 0, 0, None, None, None, None, None, None, 1.0, None, None, None, None, None,
                                0, 64, 0.5, 0.5, 0, None, None, 0, ctypes.addressof(n), None)
    assert rc == -1  # bad image size
    assert lib.gsr_mark_visible(0, None, None, None, None, None) == 0
    assert lib.gsr_visible_filter(5, None, None, 3, 1.0, None, None, None, None, 64, 64, 0.5, 0.5, 0, None, None) == -1


def test_decode_entry_points_validate_before_launching(lib):
    """gsr_decode_stage1 rejects unsupported shapes, null inputs and misaligned feature rows (they are moved as float4)
    without touching the device."""
    counts = (ctypes.c_int64 * 2)()
    params = (ctypes.c_void_p * 16)(*([0x1000] * 16))
    fake = 0x10000            # never dereferenced: validation returns first
    def stage1(feat_dim, k, feat):
        return lib.gsr_decode_stage1(8, feat_dim, k, fake, feat, None, fake, params, fake, 1 << 20, fake, fake, ctypes.addressof(counts), None)
    assert lib.gsr_decode_supported(32, 10) == 1 and lib.gsr_decode_supported(32, 17) == 0 and lib.gsr_decode_supported(16, 10) == 0
    assert stage1(16, 10, fake) == -1 and stage1(32, 0, fake) == -1       # unsupported feat_dim / n_offsets
    assert stage1(32, 10, None) == -1                                      # null input
    assert stage1(32, 10, fake + 4) == -1                                  # feature rows not 16-byte aligned
    assert lib.gsr_decode_stage1(0, 32, 10, None, None, None, None, None, None, 0, None, None, ctypes.addressof(counts), None) == 0
    assert counts[0] == 0 and counts[1] == 0                               # A == 0: nothing to do, counts zeroed
    assert lib.gsr_decode_scratch_bytes(1000) < lib.gsr_decode_scratch_bytes(100000)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from gscream_b200 import _build, _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_build, "LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GsrError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gscream_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
    for f in ("diff_gaussian_rasterization/__init__.py",):
        assert "oracle" not in open(os.path.join(ROOT, f)).read()


# ---- host-side mirror of diff_gaussian_rasterization/__init__.py ------------------------------------------------
def test_settings_field_order_matches_reference():
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")  # __init__.py:189-201


def _settings():
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    return GaussianRasterizationSettings(16, 16, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 1, torch.zeros(3), False, False)


def test_rasterizer_rejects_bad_argument_combinations():
    from diff_gaussian_rasterization import GaussianRasterizer
    r = GaussianRasterizer(raster_settings=_settings())
    P = 4
    m3, m2, op, un = torch.zeros(P, 3), torch.zeros(P, 3), torch.zeros(P, 1), torch.zeros(P, 1)
    sc, ro, col, sh, cov = torch.ones(P, 3), torch.ones(P, 4), torch.ones(P, 3), torch.ones(P, 4, 3), torch.ones(P, 6)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):  # sic, __init__.py:225
        r(means3D=m3, means2D=m2, opacities=op, uncertainties=un, scales=sc, rotations=ro)
    with pytest.raises(Exception, match="excatly one"):
        r(means3D=m3, means2D=m2, opacities=op, uncertainties=un, shs=sh, colors_precomp=col, scales=sc, rotations=ro)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m3, means2D=m2, opacities=op, uncertainties=un, colors_precomp=col)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m3, means2D=m2, opacities=op, uncertainties=un, colors_precomp=col, scales=sc, rotations=ro, cov3D_precomp=cov)


def test_C_standin_exposes_reference_names():
    import diff_gaussian_rasterization as d
    for name in ("rasterize_gaussians", "rasterize_gaussians_backward", "rasterize_aussians_filter",
                 "rasterize_aussians_filter_position2D", "mark_visible"):  # ext.cpp:16-20, typos included
        assert callable(getattr(d._C, name))
    assert callable(d.rasterize_gaussians) and callable(d.cpu_deep_copy_tuple)
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):  # rasterize_points.cu:58-60
        d._C.rasterize_aussians_filter(torch.zeros(4, 2), torch.zeros(4, 3), torch.zeros(4, 4), 1.0, torch.empty(0),
                                       torch.eye(4), torch.eye(4), 0.5, 0.5, 16, 16, False, False)


def test_scene_generator_is_deterministic_and_matches_reference_camera_convention():
    from gscream_b200 import scenes
    a, b = scenes.make_scene(1000, 320, 180, 32, 5), scenes.make_scene(1000, 320, 180, 32, 5)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert a["colors"].shape == (1000, 32) and torch.allclose(a["rotations"].norm(dim=1), torch.ones(1000), atol=1e-6)
    cam = scenes.make_camera(320, 180)
    # full_proj = world_view @ projection^T (scene/cameras.py:66-67); identity view -> projmatrix is P^T
    P = scenes.projection_matrix(0.01, 100.0, 2 * torch.atan(torch.tensor(cam["tanfovx"])).item(),
                                 2 * torch.atan(torch.tensor(cam["tanfovy"])).item())
    assert torch.allclose(cam["projmatrix"], P.t(), atol=1e-6)
    assert P[3, 2] == 1.0 and torch.isclose(P[2, 2], torch.tensor((0.01 + 100.0) / (100.0 - 0.01)))  # graphics_utils.py:69-72
    yaw = scenes.make_camera(320, 180, yaw_deg=10.0)
    assert torch.allclose(yaw["viewmatrix"][:3, :3] @ yaw["viewmatrix"][:3, :3].t(), torch.eye(3), atol=1e-6)


def test_header_is_plain_c_and_cxx(tmp_path):
    """The drop-in boundary is a C ABI: include/gsr_b200.h must compile as C99 and as C++ with no CUDA or torch headers."""
    import shutil
    import subprocess
    inc = os.path.join(ROOT, "include")
    for compiler, std, suffix in (("gcc", "-std=c99", ".c"), ("g++", "-std=c++11", ".cpp")):
        if shutil.which(compiler) is None:
            pytest.fail("%s is part of the image" % compiler)
        src = tmp_path / ("abi_probe" + suffix)
        src.write_text('#include "gsr_b200.h"\nint probe(void) { gsr_adam_tensor t; t.numel = 0; return gsr_abi_version() + (int)sizeof(t); }\n')
        r = subprocess.run([compiler, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_ctypes_prototypes_match_header_declarations():
    """Every prototype of gscream_b200._lib.PROTOTYPES has the argument count and the argument classes (pointer / int / int64 /
    size_t / float) of its declaration in include/gsr_b200.h."""
    from gscream_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "gsr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decls = dict(re.findall(r"GSR_API\s+[\w\s\*]+?\b(gsr_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S))
    assert set(decls) == set(_lib.PROTOTYPES), set(decls) ^ set(_lib.PROTOTYPES)
    def ctype_of(param):
        param = param.strip()
        if "*" in param or param.startswith("gsr_stream_t"):
            return ctypes.c_void_p
        base = param.split()[:-1] if len(param.split()) > 1 else param.split()
        base = " ".join(b for b in base if b != "const")
        return {"int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "int64_t": ctypes.c_int64, "size_t": ctypes.c_size_t}[base]

    for name, params in decls.items():
        params = params.strip()
        plist = [] if params in ("", "void") else [p for p in params.split(",") if p.strip()]
        bound = _lib.PROTOTYPES[name][1]
        assert len(plist) == len(bound), (name, len(plist), len(bound))
        for i, (p, b) in enumerate(zip(plist, bound)):
            assert ctype_of(p) is b, "%s argument %d (%s): header says %s, ctypes binds %s" % (name, i, p.strip(), ctype_of(p).__name__, b.__name__)


def test_binning_buffer_describes_itself(lib):
    """The binning buffer's layout follows from its byte size (gsr_binning_capacity is the inverse of gsr_binning_bytes), so a
    buffer sized from an ESTIMATE of num_rendered is laid out identically by forward, backward and export."""
    for (P, W, H) in ((1000, 64, 64), (10 ** 6, 1920, 1080), (5 * 10 ** 5, 1008, 567), (7, 1, 1)):
        prev_bytes = 0
        for R in (1, 2, 4095, 4096, 4097, 10 ** 5, 7302277, 4 * 10 ** 7):
            nbytes = lib.gsr_binning_bytes(P, R, W, H)
            assert nbytes >= prev_bytes            # monotone in the instance count
            prev_bytes = nbytes
            cap = lib.gsr_binning_capacity(P, W, H, nbytes)
            assert cap >= R                        # a buffer sized for R holds R ...
            assert lib.gsr_binning_bytes(P, cap, W, H) <= nbytes        # ... cap's own layout fits ...
            assert lib.gsr_binning_bytes(P, cap + 1, W, H) > nbytes     # ... and cap is the largest such count
        assert lib.gsr_binning_capacity(P, W, H, 0) == 0 and lib.gsr_binning_capacity(P, W, H, 100) == 0


def test_capacity_estimate_policy():
    """gscream_b200._C sizes the binning buffer before num_rendered reaches the host: no history -> exact path; afterwards 25 %
    above a slowly decaying running maximum."""
    from gscream_b200 import _C
    key = ("test-device", 123, 45, 67)
    _C._R_HINT.pop(key, None)
    assert _C._capacity_guess(key) is None
    _C._capacity_update(key, 1000)
    assert _C._capacity_guess(key) == int(1000 * 1.25) + 65536
    _C._capacity_update(key, 10)                       # a much smaller view: the maximum decays by 2 % per call, not at once
    assert _C._R_HINT[key] == pytest.approx(980.0)
    _C._capacity_update(key, 5000)                     # a larger one takes over immediately
    assert _C._R_HINT[key] == 5000.0
    _C._R_HINT.pop(key, None)
