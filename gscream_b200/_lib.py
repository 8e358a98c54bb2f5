"""ctypes binding of libgsr_b200.so — the C ABI declared in include/gsr_b200.h.

There is no CPU fallback: if the CUDA library is missing this module raises, and every product
entry point raises with it.
"""
import ctypes
import os

from . import _build

_c = ctypes
_vp, _i, _f, _i64, _sz = _c.c_void_p, _c.c_int, _c.c_float, _c.c_int64, _c.c_size_t

# name -> (restype, argtypes); order and meaning exactly as include/gsr_b200.h
PROTOTYPES = {
    "gsr_abi_version": (_i, []),
    "gsr_supported_channels": (_i, [_i]),
    "gsr_error_string": (_c.c_char_p, [_i]),
    "gsr_geom_bytes": (_sz, [_i]),
    "gsr_image_bytes": (_sz, [_i, _i]),
    "gsr_binning_bytes": (_sz, [_i, _i64, _i, _i]),
    "gsr_binning_capacity": (_i64, [_i, _i, _i, _sz]),
    "gsr_forward_stage1": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp,
                                _i, _i, _f, _f, _i, _vp, _vp, _sz, _vp, _vp]),
    "gsr_forward_stage2": (_i, [_i, _i, _i64, _vp, _vp, _i, _i, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp]),
    "gsr_backward": (_i, [_i, _i, _i, _i, _i64, _vp, _i, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp,
                          _vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "gsr_visible_filter": (_i, [_i, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _i, _vp, _vp]),
    "gsr_position2d_filter": (_i, [_i, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp]),
    "gsr_mark_visible": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
    "gsr_debug_export": (_i, [_i, _i64, _i, _i, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gsr_decode_supported": (_i, [_i, _i]),
    "gsr_decode_scratch_bytes": (_sz, [_i]),
    "gsr_decode_stage1": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "gsr_decode_stage2": (_i, [_i, _i, _i, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gsr_decode_backward": (_i, [_i, _i, _i, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gsr_training_statis_scratch_bytes": (_sz, [_i, _i]),
    "gsr_training_statis": (_i, [_i, _i, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gsr_l1_ssim_forward": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "gsr_l1_ssim_backward": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "gsr_debug_plain_point_list": (_i, [_i]),
    "gsr_depth_align_l1_forward": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gsr_depth_align_l1_backward": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gsr_depth_grad_forward": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gsr_depth_grad_backward": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "gsr_adam_step": (_i, [_i, _vp, _vp]),
    "gsr_launch_count": (_i64, [_i]),
    "gsr_profile_enable": (_i, [_i]),
    "gsr_profile_read": (_i, [_i, _vp, _i]),
}

_LIB = None


class GsrError(RuntimeError):
    pass


def library_path():
    # GSR_LIB_PATH: development aid for A/B-ing differently compiled builds of the same library
    return os.environ.get("GSR_LIB_PATH") or _build.LIB


def load():
    """Load (never build implicitly on a GPU box: the .so travels with the repo)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise GsrError(
            "libgsr_b200.so is missing (%s). Build it with `python -m gscream_b200._build` "
            "(or __graft_entry__.build()); there is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    ab_variant = bool(os.environ.get("GSR_LIB_PATH"))
    for name, (res, args) in PROTOTYPES.items():
        if ab_variant and not hasattr(lib, name):
            continue  # an older A/B build of the library (development aid only): symbols added later are absent
        fn = getattr(lib, name)  # AttributeError here == ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.gsr_abi_version() != 2:
        raise GsrError("libgsr_b200.so ABI version mismatch")
    _LIB = lib
    return lib


def check(code):
    if code != 0:
        msg = load().gsr_error_string(code).decode()
        if code == -4:
            # the reference throws std::runtime_error here (rasterizer_impl.cu:246-249)
            raise RuntimeError(msg)
        raise GsrError("libgsr_b200 call failed (%d): %s" % (code, msg))
