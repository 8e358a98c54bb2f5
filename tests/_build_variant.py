"""Scratch: build an A/B variant of the library with extra -D flags:  python tests/_build_variant.py <name> [-DX=1 ...]
-> variants/libgsr_<name>.so (git-ignored; travels to the GPU box)."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gscream_b200 import _build as b
name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "variants", "libgsr_%s.so" % name)
obj = os.path.join(ROOT, "variants", "obj_" + name)
os.makedirs(obj, exist_ok=True)
def cc(src):
    o = os.path.join(obj, src.replace(".cu", ".o"))
    r = subprocess.run([b.NVCC] + b.FLAGS + extra + ["-c", os.path.join(b.CSRC, src), "-o", o], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return o
with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, b.SOURCES))
r = subprocess.run([b.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcudart"], capture_output=True, text=True)
assert r.returncode == 0, r.stderr
print(out)
