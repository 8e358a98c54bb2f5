"""Scratch: tiny runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
rasterizer fwd+bwd (C = 32 ragged, C = 3, long tile lists, packed and plain point lists), fused decode fwd+bwd, fused L1+SSIM
fwd+bwd, densification statistics."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _ref_utils as ru
import _anchor_decode as ad
from gscream_b200 import _lib, decode, losses, scenes, stats, rasterizer as ours
lib = _lib.load()
for plain in (0, 1):
    lib.gsr_debug_plain_point_list(plain)
    for (P, W, H, C, seed, sm) in ((900, 83, 50, 32, 1, 4.0), (700, 64, 48, 3, 2, 5.0)):
        sc = scenes.make_scene(P, W, H, C, seed, scale_mult=sm); cam = scenes.make_camera(W, H); g = scenes.make_upstream_grads(C, W, H, seed)
        m = ru.run_impl(ours, sc, cam, g)
        print("raster ok", "plain" if plain else "packed", P, W, H, C, m["num_rendered"])
lib.gsr_debug_plain_point_list(0)
dev = torch.device("cuda")
W, H = 96, 64
cam = scenes.make_camera(W, H)
bg = torch.zeros(3, device=dev)
pc = ad.SyntheticAnchors(301, n_offsets=10, seed=3, tanfov=(cam["tanfovx"], cam["tanfovy"])).to(dev)
vis, _, _ = ad.prefilter_position2D(ours, cam, pc, bg)
class Cam: camera_center = cam["campos"].to(dev)
xyz, color, opacity, unc, scaling, rot, nop, mask = decode.generate_neural_gaussians(Cam, pc, vis, is_training=True)
ssp = torch.zeros_like(xyz, requires_grad=True)
rast = ours.GaussianRasterizer(raster_settings=ad.make_settings(ours, cam, bg, dev))
image, depth, uncer, radii = rast(means3D=xyz, means2D=ssp, shs=None, colors_precomp=color, opacities=opacity, uncertainties=unc, scales=scaling, rotations=rot, cov3D_precomp=None)
target = torch.rand(3, H, W, device=dev)
s, l1 = losses.l1_ssim(image, target, (torch.rand(1, H, W, device=dev) > 0.3).float())
(0.8 * l1 + 0.2 * (1 - s) + depth.mean() + nop.sum() * 1e-3).backward()
ad.init_statis_buffers(pc)
stats.training_statis(pc, ssp, nop, radii > 0, mask, vis)
torch.cuda.synchronize()
print("decode / loss / statistics ok", int(vis.sum()), xyz.shape[0], float(s), float(l1))
