"""TEST INFRASTRUCTURE — generates tests/golden/*.npz from the REFERENCE's own CUDA build.

Run on a GPU box (`gpurun -- python tests/golden/make_golden.py`): it executes the unmodified
W-Ted/GScream rasterizer compiled by oracle/build_ref.py (oracle/_ref, NUM_CHANNELS 3 and 32) on small
seeded scenes and stores inputs, every intermediate that can be parsed out of the reference's scratch
buffers (radii, means2D, depths, cov3D, conic_opacity, tiles_touched, keys, point_list, ranges, final_T,
n_contrib), the rendered planes and all gradients.  Output goes to gpurun_out/golden/, from where the
files are copied into tests/golden/ and committed.  These vectors pin the CPU oracle (tests/test_oracle_golden.py)
and are compared against the CUDA product as well (tests/test_gpu_parity.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import _ref_utils as ru  # noqa: E402
from gscream_b200 import scenes  # noqa: E402

# name -> (P, W, H, C, seed, scale_mult, bg_value, yaw)
CASES = {
    "c3_small": (1500, 160, 96, 3, 11, 2.0, 0.25, 0.0),
    "c3_ragged": (1200, 150, 83, 3, 12, 4.0, 1.0, 7.0),      # W,H not multiples of 16; big splats; yawed camera
    "c32_small": (1200, 128, 80, 32, 13, 2.0, 0.1, 0.0),
    "c32_dense": (2500, 96, 64, 32, 14, 5.0, 0.0, -5.0),     # long tile lists (> 256 per tile), saturating pixels
}


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (P, W, H, C, seed, smult, bgv, yaw) in CASES.items():
        mod = ru.load_ref(C)
        scene = scenes.make_scene(P, W, H, C, seed, scale_mult=smult, bg_value=bgv)
        cam = scenes.make_camera(W, H, yaw_deg=yaw)
        grads = scenes.make_upstream_grads(C, W, H, seed)
        r = ru.run_impl(mod, scene, cam, grads)
        R = r["num_rendered"]
        geom = ru.parse_ref_geom(r.pop("_geom").cpu().numpy(), P)
        img = ru.parse_ref_image(r.pop("_img").cpu().numpy(), W, H)
        binn = ru.parse_ref_binning(r.pop("_binning").cpu().numpy(), R)
        save = dict(P=P, W=W, H=H, C=C, seed=seed, scale_mult=smult, bg_value=bgv, yaw=yaw, num_rendered=R)
        save.update({"in_" + k: v.numpy() for k, v in scene.items()})
        save.update({"cam_" + k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in cam.items()})
        save.update({"g_color": grads[0].numpy(), "g_depth": grads[1].numpy(), "g_unc": grads[2].numpy()})
        save.update({k: v for k, v in r.items()})
        save.update({"geom_" + k: v for k, v in geom.items() if k not in ("clamped", "rgb", "internal_radii")})
        save.update({"img_" + k: v for k, v in img.items()})
        save.update({"bin_" + k: v for k, v in binn.items()})
        # second backward run: the reference's own run-to-run spread (float atomics in arbitrary order)
        r2 = ru.run_impl(mod, scene, cam, grads)
        for k in ("dL_dmeans3D", "dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_duncertainty", "dL_dscales", "dL_drotations"):
            save["rerun_" + k] = r2[k]
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **save)
        print(name, "R=%d visible=%d" % (R, int((r["radii"] > 0).sum())), "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
