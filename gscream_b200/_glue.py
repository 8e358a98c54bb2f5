"""Build and load the compiled torch glue (csrc/gsr_torch_glue.cpp): the pybind11 counterpart of the reference's
`diff_gaussian_rasterization._C` (ext.cpp:16-20, rasterize_points.cu:35-373) above the C ABI of libgsr_b200.so.

    build()  g++ (through torch.utils.cpp_extension, no nvcc: the file holds no kernels) -> gscream_b200/_glue/_gsr_glue.so,
             linked against ../libgsr_b200.so with an $ORIGIN-relative rpath, so the pair travels to the GPU box as it is
    load()   import that extension module (never builds implicitly)

`GSR_GLUE=cpp` makes gscream_b200._C route its five reference entry points through this module; the ctypes glue stays the default.
"""
import hashlib
import importlib.util
import os
import sys

from . import _build

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG, "csrc", "gsr_torch_glue.cpp")
OUT_DIR = os.path.join(PKG, "_glue")
SO = os.path.join(OUT_DIR, "_gsr_glue.so")
STAMP = os.path.join(OUT_DIR, "digest.txt")
_MOD = None


def _digest():
    import torch
    h = hashlib.sha256()
    for path in (SRC, os.path.join(PKG, "..", "include", "gsr_b200.h")):
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(torch.__version__.encode())
    return h.hexdigest()


def is_fresh():
    return os.path.exists(SO) and os.path.exists(STAMP) and open(STAMP).read() == _digest()


def build(force=False, verbose=False):
    if not force and is_fresh():
        return SO
    _build.build()                                    # the library the glue links against
    os.makedirs(OUT_DIR, exist_ok=True)
    from torch.utils.cpp_extension import load
    load(name="_gsr_glue", sources=[SRC], extra_cflags=["-O2", "-std=c++17"],
         # rpath: ninja turns $$ into $, the shell turns \$ into $ -> the linker sees $ORIGIN/..
         extra_ldflags=["-L" + PKG, "-l:libgsr_b200.so", "-Wl,-rpath,\\$$ORIGIN/.."],
         build_directory=OUT_DIR, with_cuda=True, is_python_module=False, verbose=verbose)
    for f in os.listdir(OUT_DIR):                     # keep only what must travel
        if f.endswith(".o") or f in ("build.ninja", ".ninja_deps", ".ninja_log"):
            os.remove(os.path.join(OUT_DIR, f))
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return SO


def load():
    global _MOD
    if _MOD is not None:
        return _MOD
    if not os.path.exists(SO):
        raise RuntimeError("compiled glue missing (%s): build it with `python -m gscream_b200._glue` or __graft_entry__.build()" % SO)
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location("_gsr_glue", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _MOD = mod
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
