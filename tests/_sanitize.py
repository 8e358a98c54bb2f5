"""Scratch: tiny runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
rasterizer fwd+bwd (C = 32 ragged, C = 3, long tile lists, packed and plain point lists), fused decode fwd+bwd, fused L1+SSIM
fwd+bwd, densification statistics; aligned-depth L1 + four-scale gradient loss fwd+bwd and the multi-tensor Adam step
(`python tests/_sanitize.py new` runs only that last group; `python tests/_sanitize.py decode` only the decode kernels, forward and
backward, on inputs spanning several tiles per warp and several CTA iterations at k = 10, 16 and 3)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _ref_utils as ru
import _anchor_decode as ad
from gscream_b200 import _lib, decode, losses, scenes, stats, rasterizer as ours
lib = _lib.load()
if len(sys.argv) > 1 and sys.argv[1] == "decode":
    dev = torch.device("cuda")
    class _C:
        camera_center = torch.tensor([0.05, 0.1, -0.3], device=dev)
    # 148 CTAs x 8 tiles x 16 anchors = 18 944 anchors per CTA iteration at k <= 10 (half of that above): the first case needs two
    for (A, k, frac) in ((24000, 10, 0.9), (10500, 16, None), (333, 3, 0.5)):
        pc = ad.SyntheticAnchors(A, n_offsets=k, seed=A).to(dev)
        vis = None if frac is None else (torch.rand(A, device=dev) < frac)
        outs = decode.generate_neural_gaussians(_C, pc, vis, is_training=True)
        sum((o * torch.randn_like(o)).sum() for o in outs[:7]).backward()
        torch.cuda.synchronize()
        print("decode ok", A, k, outs[0].shape[0])
    sys.exit(0)
ONLY_NEW = len(sys.argv) > 1 and sys.argv[1] == "new"
for plain in (() if ONLY_NEW else (0, 1)):
    lib.gsr_debug_plain_point_list(plain)
    for (P, W, H, C, seed, sm) in ((900, 83, 50, 32, 1, 4.0), (700, 64, 48, 3, 2, 5.0)):
        sc = scenes.make_scene(P, W, H, C, seed, scale_mult=sm); cam = scenes.make_camera(W, H); g = scenes.make_upstream_grads(C, W, H, seed)
        m = ru.run_impl(ours, sc, cam, g)
        print("raster ok", "plain" if plain else "packed", P, W, H, C, m["num_rendered"])
lib.gsr_debug_plain_point_list(0)
dev = torch.device("cuda")
def depth_and_adam():
    from gscream_b200 import optim
    for (B, H, W) in ((1, 61, 83), (2, 5, 3), (1, 1, 1)):
        d = (2 + 6 * torch.rand(B, H, W, device=dev)).requires_grad_(True)
        y = 0.7 * d.detach() + 1.3 + 0.4 * torch.randn(B, H, W, device=dev)
        m = (torch.rand(B, H, W, device=dev) > 0.3).float()
        l1, gl = losses.aligned_depth_losses(d, y, m, None, m)
        (l1 + 0.5 * gl).backward()
        p = d.detach().clone().requires_grad_(True)
        losses.multiscale_gradient_loss(p, y, None).backward()
        print("depth losses ok", B, H, W, float(l1), float(gl))
    store = torch.randn(5000, device=dev)
    ps = [torch.nn.Parameter(store[1:4098]), torch.nn.Parameter(torch.randn(33, 7, device=dev)), torch.nn.Parameter(torch.randn(1, device=dev))]
    ps += [torch.nn.Parameter(torch.randn(10 + i, device=dev)) for i in range(30)]
    opt = optim.Adam([{"params": [q], "lr": 1e-3, "name": str(i)} for i, q in enumerate(ps)], lr=0.0, eps=1e-15)
    for _ in range(2):
        for q in ps:
            q.grad = torch.randn_like(q)
        opt.step()
    torch.cuda.synchronize()
    print("adam ok", len(ps))
depth_and_adam()
if ONLY_NEW:
    sys.exit(0)
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
