"""TEST INFRASTRUCTURE — numpy/ctypes front-end of the CPU oracle (oracle/oracle.c).

The oracle is a plain-C restatement of the reference rasterizer
(/root/reference/submodules/diff-gaussian-rasterization, CUDA-only upstream).  It exists to
check the CUDA product, never to serve it: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference legs may import this module.

Parity status: pinned against the reference's own CUDA build through tests/golden/*.npz
(made on a B200 by tests/golden/make_golden.py from oracle/_ref).

`forward()` / `backward()` mirror CudaRasterizer::Rasterizer::forward / backward
(cuda_rasterizer/rasterizer_impl.cu:199-347, 536-643) stage by stage and return every
intermediate (radii, tiles_touched, keys, sorted point_list, ranges, n_contrib, ...), so integer
artefacts can be compared bit for bit and floats within a stated tolerance.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build():
    """Compile oracle.c (gcc, a second or two).  Idempotent."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _lib(precision):
    if precision not in _LIBS:
        path = os.path.join(_HERE, "liboracle_%s.so" % precision)
        src = os.path.join(_HERE, "oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _LIBS[precision] = ctypes.CDLL(path)
    return _LIBS[precision]


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class Oracle:
    """precision: 'f32' (IEEE fp32, no FMA) or 'f64'."""

    def __init__(self, precision="f32"):
        assert precision in ("f32", "f64")
        self.precision = precision
        self.lib = _lib(precision)
        self.real = np.float32 if precision == "f32" else np.float64
        self.lib.orc_binning.restype = ctypes.c_int64
        assert self.lib.orc_real_bytes() == np.dtype(self.real).itemsize

    # -- K1 / K10 / K11 ---------------------------------------------------------------------
    def preprocess(self, means3D, scales, rotations, opacities, uncertainties, viewmatrix, projmatrix,
                   W, H, tanfovx, tanfovy, scale_modifier=1.0, cov3D_precomp=None, mode=0):
        means3D = _f32(means3D)
        P = means3D.shape[0]
        r = self.real
        out = dict(
            radii=np.zeros(P, np.int32),
            xy=np.zeros((P, 2), r), depths=np.zeros(P, r), cov3D=np.zeros((P, 6), r),
            conic_opacity=np.zeros((P, 4), r), tiles_touched=np.zeros(P, np.uint32),
            gauss_uncertainty=np.zeros(P, r), pos2d_x=np.zeros(P, r), pos2d_y=np.zeros(P, r),
        )
        scales, rotations = _f32(scales), _f32(rotations)
        cov3D_precomp = _f32(cov3D_precomp)
        opac = _f32(opacities) if opacities is not None else np.zeros(P, np.float32)
        unc = _f32(uncertainties) if uncertainties is not None else np.zeros(P, np.float32)
        view, proj = _f32(viewmatrix).reshape(-1), _f32(projmatrix).reshape(-1)
        self.lib.orc_preprocess(
            ctypes.c_int(mode), ctypes.c_int(P), _p(means3D), _p(scales), ctypes.c_float(scale_modifier),
            _p(rotations), _p(opac.reshape(-1)), _p(unc.reshape(-1)), _p(cov3D_precomp), _p(view), _p(proj),
            ctypes.c_int(W), ctypes.c_int(H), ctypes.c_float(tanfovx), ctypes.c_float(tanfovy),
            _p(out["radii"]), _p(out["xy"]), _p(out["depths"]), _p(out["cov3D"]), _p(out["conic_opacity"]),
            _p(out["tiles_touched"]), _p(out["gauss_uncertainty"]), _p(out["pos2d_x"]), _p(out["pos2d_y"]))
        if cov3D_precomp is not None:
            out["cov3D"] = cov3D_precomp.astype(r)
        return out

    def mark_visible(self, means3D, viewmatrix, projmatrix):
        means3D = _f32(means3D)
        P = means3D.shape[0]
        present = np.zeros(P, np.uint8)
        self.lib.orc_mark_visible(ctypes.c_int(P), _p(means3D), _p(_f32(viewmatrix).reshape(-1)),
                                  _p(_f32(projmatrix).reshape(-1)), _p(present))
        return present.astype(bool)

    # -- K2..K5 -----------------------------------------------------------------------------
    def binning(self, xy, depths, radii, W, H):
        """xy/depths are rounded to fp32 first: keys embed the fp32 depth bit pattern."""
        xy32 = np.ascontiguousarray(np.asarray(xy, np.float32))
        depth_bits = np.ascontiguousarray(np.asarray(depths, np.float32)).view(np.uint32)
        radii = np.ascontiguousarray(np.asarray(radii, np.int32))
        P = radii.shape[0]
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        offsets = np.zeros(P, np.uint32)
        args = (ctypes.c_int(P), _p(xy32), _p(depth_bits), _p(radii), ctypes.c_int(W), ctypes.c_int(H))
        R = int(self.lib.orc_binning(*args, _p(offsets), None, None, None, None, None))
        keys_u = np.zeros(max(R, 1), np.uint64)
        vals_u = np.zeros(max(R, 1), np.uint32)
        keys_s = np.zeros(max(R, 1), np.uint64)
        plist = np.zeros(max(R, 1), np.uint32)
        ranges = np.zeros((tiles, 2), np.uint32)
        R2 = int(self.lib.orc_binning(*args, _p(offsets), _p(keys_u), _p(vals_u), _p(keys_s), _p(plist), _p(ranges)))
        assert R2 == R
        return dict(num_rendered=R, point_offsets=offsets, keys_unsorted=keys_u[:R], values_unsorted=vals_u[:R],
                    keys_sorted=keys_s[:R], point_list=plist[:R], ranges=ranges)

    # -- K6 ---------------------------------------------------------------------------------
    def render_forward(self, W, H, ranges, point_list, xy, features, depths, unc, conic_opacity, bg):
        r = self.real
        features = np.ascontiguousarray(np.asarray(features, r))
        C = features.shape[1]
        N = W * H
        out = dict(final_T=np.zeros(N, r), n_contrib=np.zeros(N, np.uint32), color=np.zeros((C, H, W), r),
                   depth=np.zeros((1, H, W), r), uncertainty=np.zeros((1, H, W), r))
        a = lambda x: np.ascontiguousarray(np.asarray(x, r))
        self._keep = [a(xy), features, a(depths), a(unc), a(conic_opacity), a(bg),
                      np.ascontiguousarray(ranges, dtype=np.uint32), np.ascontiguousarray(point_list, dtype=np.uint32)]
        xy_, f_, d_, u_, co_, bg_, rg_, pl_ = self._keep
        self.lib.orc_render_forward(ctypes.c_int(C), ctypes.c_int(W), ctypes.c_int(H), _p(rg_), _p(pl_), _p(xy_),
                                    _p(f_), _p(d_), _p(u_), _p(co_), _p(bg_), _p(out["final_T"]), _p(out["n_contrib"]),
                                    _p(out["color"]), _p(out["depth"]), _p(out["uncertainty"]))
        return out

    # -- whole forward (Rasterizer::forward, rasterizer_impl.cu:199-347) ---------------------
    def forward(self, means3D, colors_precomp, opacities, uncertainties, scales, rotations, viewmatrix,
                projmatrix, bg, W, H, tanfovx, tanfovy, scale_modifier=1.0, cov3D_precomp=None):
        pre = self.preprocess(means3D, scales, rotations, opacities, uncertainties, viewmatrix, projmatrix,
                              W, H, tanfovx, tanfovy, scale_modifier, cov3D_precomp)
        binn = self.binning(pre["xy"], pre["depths"], pre["radii"], W, H)
        # the reference stores fp32 intermediates; the f64 oracle keeps them in double on purpose
        img = self.render_forward(W, H, binn["ranges"], binn["point_list"], pre["xy"], colors_precomp,
                                  pre["depths"], pre["gauss_uncertainty"], pre["conic_opacity"], bg)
        out = {}
        for part in (pre, binn, img):
            for k, v in part.items():
                assert k not in out, "oracle stage outputs must not shadow each other: %s" % k
                out[k] = v
        return out

    # -- K7 + K8 + K9 (Rasterizer::backward, rasterizer_impl.cu:536-643) ---------------------
    def backward(self, fwd, means3D, colors_precomp, scales, rotations, viewmatrix, projmatrix, bg, W, H,
                 tanfovx, tanfovy, dL_dcolor, dL_ddepth, dL_dunc, scale_modifier=1.0, cov3D_precomp=None):
        r = self.real
        a = lambda x: np.ascontiguousarray(np.asarray(x, r))
        means3D = _f32(means3D)
        P = means3D.shape[0]
        colors = a(colors_precomp)
        C = colors.shape[1]
        g = dict(dL_dmean2D=np.zeros((P, 3), r), dL_dconic=np.zeros((P, 4), r), dL_dopacity=np.zeros((P, 1), r),
                 dL_dcolors=np.zeros((P, C), r), dL_ddepths=np.zeros((P, 1), r), dL_duncertainty=np.zeros((P, 1), r),
                 dL_dmeans3D=np.zeros((P, 3), r), dL_dcov3D=np.zeros((P, 6), r), dL_dscales=np.zeros((P, 3), r),
                 dL_drotations=np.zeros((P, 4), r))
        rg = np.ascontiguousarray(fwd["ranges"], dtype=np.uint32)
        pl = np.ascontiguousarray(fwd["point_list"], dtype=np.uint32)
        nc = np.ascontiguousarray(fwd["n_contrib"], dtype=np.uint32)
        xy, co, dep, unc, fT, bg_ = a(fwd["xy"]), a(fwd["conic_opacity"]), a(fwd["depths"]), a(fwd["gauss_uncertainty"]), a(fwd["final_T"]), a(bg)
        gc, gd, gu = a(dL_dcolor), a(dL_ddepth), a(dL_dunc)
        self.lib.orc_render_backward(ctypes.c_int(C), ctypes.c_int(W), ctypes.c_int(H), _p(rg), _p(pl), _p(bg_), _p(xy),
                                     _p(co), _p(colors), _p(dep), _p(unc), _p(fT), _p(nc), _p(gc), _p(gd), _p(gu),
                                     _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]),
                                     _p(g["dL_ddepths"]), _p(g["dL_duncertainty"]))
        cov3D = a(fwd["cov3D"] if cov3D_precomp is None else cov3D_precomp)
        sc, ro = _f32(scales), _f32(rotations)
        radii = np.ascontiguousarray(fwd["radii"], dtype=np.int32)
        self.lib.orc_preprocess_backward(ctypes.c_int(P), _p(means3D), _p(radii), _p(sc), _p(ro),
                                         ctypes.c_float(scale_modifier), _p(cov3D), _p(_f32(viewmatrix).reshape(-1)),
                                         _p(_f32(projmatrix).reshape(-1)), ctypes.c_int(W), ctypes.c_int(H),
                                         ctypes.c_float(tanfovx), ctypes.c_float(tanfovy),
                                         _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_ddepths"].reshape(-1)),
                                         _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]))
        return g


def get_higher_msb(n):
    lib = _lib("f32")
    lib.orc_get_higher_msb.restype = ctypes.c_uint32
    return int(lib.orc_get_higher_msb(ctypes.c_uint32(n)))
