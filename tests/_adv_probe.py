"""Scratch: run-to-run spread of the adversarial-parameter case (tests/test_gpu_parity.py) for both implementations and their
distance to the fp64 oracle, per gradient tensor."""
import math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _ref_utils as ru
from gscream_b200 import scenes, rasterizer as ours
import test_gpu_parity as tg
for C in (32, 3):
    P, W, H = 6000, 333, 190
    g = torch.Generator().manual_seed(900 + C)
    sc = scenes.make_scene(P, W, H, C, 900 + C, scale_mult=2.0, bg_value=0.3 if C == 3 else 0.0)
    special = torch.tensor([1 / 255, 0.99 / 255, 1.01 / 255, 0.0039, 0.004, 0.0, 0.98, 0.99, 1.0, 1.5, 0.5, 0.25])
    sc["opacities"] = special[torch.randint(0, len(special), (P, 1), generator=g)].contiguous()
    sc["scales"] = torch.exp(torch.empty(P, 3).uniform_(math.log(1e-4), math.log(3.0), generator=g)).contiguous()
    sc["rotations"] = (torch.randn(P, 4, generator=g) * torch.empty(P, 1).uniform_(0.5, 2.0, generator=g)).contiguous()
    sc["means3D"][: P // 10, :2] *= 3.0
    cam = scenes.make_camera(W, H, yaw_deg=3.0)
    grads = scenes.make_upstream_grads(C, W, H, 900 + C)
    ref_mod = ru.load_ref(C)
    rs = [ru.run_impl(ref_mod, sc, cam, grads) for _ in range(4)]
    ms = [ru.run_impl(ours, sc, cam, grads) for _ in range(6)]
    f, b = tg._oracle_run(sc, cam, grads, "f64")
    for k in tg.GRAD_KEYS:
        ref = rs[0][k]
        scale = float(np.abs(ref).max())
        sp_r = max(float(np.abs(r[k] - ref).max()) for r in rs[1:])
        sp_m = max(float(np.abs(m[k] - ms[0][k]).max()) for m in ms[1:])
        errs = [float(np.abs(m[k] - ref).max()) for m in ms]
        line = "C=%d %-16s scale %.3e spread ref %.2e ours %.2e  err ours-ref max %.2e min %.2e" % (C, k, scale, sp_r, sp_m, max(errs), min(errs))
        if k in b:
            o = b[k].reshape(ref.shape)
            line += "  | vs f64 oracle: ref %.2e ours %.2e" % (float(np.abs(ref - o).max()), float(np.abs(ms[0][k] - o).max()))
            i = np.unravel_index(np.argmax(np.abs(ms[0][k] - ref)), ref.shape)
            line += "  worst idx %s ours %.6e ref %.6e f64 %.6e" % (i, ms[0][k][i], ref[i], o[i])
        print(line, flush=True)
