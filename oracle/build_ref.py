"""TEST INFRASTRUCTURE — builds the *unmodified* reference rasterizer into oracle/_ref/.

Compiles W-Ted/GScream's `submodules/diff-gaussian-rasterization` (5 sources listed in its
setup.py:23-28) for sm_100a from the sources where they lie under /root/reference.  Nothing is
copied into git history: outputs go to oracle/_ref/ (git-ignored, but it travels to the GPU box).

Two variants are produced, each laid out like an installed `diff_gaussian_rasterization` package
(`_C.so` + the reference's own Python wrapper staged next to it, exactly what
`pip install --target` would place there):

    oracle/_ref/dgr3/diff_gaussian_rasterization/     NUM_CHANNELS 3   (stock)
    oracle/_ref/dgr32/diff_gaussian_rasterization/    NUM_CHANNELS 32  (via --pre-include
                                                       oracle/ref_config_c32.h; no source edit)

Flags: `--pre-include cstdint` works around the missing <cstdint> in
cuda_rasterizer/rasterizer_impl.h:24 on gcc 13.  No fast-math (reference setup.py:29 has none).

Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline legs may load the
result.  The product (gscream_b200/) never does.
"""
import os
import shutil
import sys

REF = os.environ.get("GSCREAM_REFERENCE", "/root/reference")
DGR = os.path.join(REF, "submodules", "diff-gaussian-rasterization")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

SOURCES = [
    "cuda_rasterizer/rasterizer_impl.cu",
    "cuda_rasterizer/forward.cu",
    "cuda_rasterizer/backward.cu",
    "rasterize_points.cu",
    "ext.cpp",
]


def variant_dir(channels):
    return os.path.join(OUT, "dgr%d" % channels, "diff_gaussian_rasterization")


def is_built(channels):
    d = variant_dir(channels)
    return os.path.exists(os.path.join(d, "_C.so")) and os.path.exists(os.path.join(d, "__init__.py"))


def build_variant(channels, verbose=False):
    if not os.path.isdir(DGR):
        raise RuntimeError("reference sources not found at %s" % DGR)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    d = variant_dir(channels)
    os.makedirs(d, exist_ok=True)
    cuda_flags = [
        "-I" + os.path.join(DGR, "third_party", "glm"),
        "--pre-include", "cstdint",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo",
    ]
    if channels != 3:
        cuda_flags += ["--pre-include", os.path.join(HERE, "ref_config_c%d.h" % channels)]
    load(
        name="_C",
        sources=[os.path.join(DGR, s) for s in SOURCES],
        extra_cuda_cflags=cuda_flags,
        extra_cflags=["-O3"],
        build_directory=d,
        with_cuda=True,
        is_python_module=False,   # build only; importing needs no GPU but is not needed here
        verbose=verbose,
    )
    # stage the reference's own Python wrapper beside its extension (installed-package layout)
    shutil.copyfile(os.path.join(DGR, "diff_gaussian_rasterization", "__init__.py"),
                    os.path.join(d, "__init__.py"))
    # keep only the artefacts that must travel
    for f in os.listdir(d):
        if f.endswith(".o") or f in ("build.ninja", ".ninja_deps", ".ninja_log"):
            os.remove(os.path.join(d, f))
    return d


def build_all(verbose=False):
    """One subprocess per variant: torch.utils.cpp_extension.load() renames a second module called `_C` built in the same
    process to `_C_v1` (its JIT version bump), which the staged package would not import."""
    import subprocess
    for c in (3, 32):
        if not is_built(c):
            d = variant_dir(c)
            if os.path.isdir(d):
                for f in os.listdir(d):
                    if f.startswith("_C_v") and f.endswith(".so"):
                        os.remove(os.path.join(d, f))
            subprocess.run([sys.executable, os.path.abspath(__file__), str(c)], check=True,
                           stdout=None if verbose else subprocess.DEVNULL, stderr=None if verbose else subprocess.DEVNULL)
            if not is_built(c):
                raise RuntimeError("reference build for NUM_CHANNELS=%d did not produce _C.so in %s" % (c, d))


if __name__ == "__main__":
    chans = [int(a) for a in sys.argv[1:]] or [3, 32]
    for c in chans:
        print("building reference rasterizer, NUM_CHANNELS=%d ->" % c, build_variant(c, verbose=True))
