"""Scratch: tiny fwd+bwd for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _ref_utils as ru
from gscream_b200 import scenes, rasterizer as ours
for (P, W, H, C, seed, sm) in ((900, 83, 50, 32, 1, 4.0), (700, 64, 48, 3, 2, 5.0)):
    sc = scenes.make_scene(P, W, H, C, seed, scale_mult=sm); cam = scenes.make_camera(W, H); g = scenes.make_upstream_grads(C, W, H, seed)
    m = ru.run_impl(ours, sc, cam, g)
    print("ok", P, W, H, C, m["num_rendered"])
