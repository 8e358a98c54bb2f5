"""BENCH / TEST INFRASTRUCTURE — "train-step-equivalent" loop standing in for BASELINE.json configs[3]
(the full train.py loop on SPIN-NeRF "book", which cannot run here: dataset and six Python dependencies are absent,
SURVEY.md section 8d "Config 4").  One iteration = what train.py:497-603 does around the hot path:

    prefilter_voxel (visible_filter on the anchors)            gaussian_renderer/__init__.py:190-246
    generate_neural_gaussians + rasterize                       gaussian_renderer/__init__.py:104-179
    L1 + (1 - SSIM) on the image, scale/shift-aligned L1 depth  utils/loss_utils.py:27-28,80-110,131-164; train.py:535-560
    + the four-scale gradient-matching depth loss               train.py:232-251, 556-560
    scaling regulariser, backward, densification statistics,    train.py:575-612, scene/gaussian_model.py:729-757
    Adam step

on a synthetic anchor model (10^5 anchors x 10 offsets, 1008x567) with synthetic target image / depth.  Both arms run
this same file; they differ only in `decode` (torch restatement vs gscream_b200.decode) and `rast` (reference build vs
gscream_b200.rasterizer).  The reference arm's losses are eager torch (its own code path); this repo's arm uses the fused L1 + SSIM and aligned-depth-L1
and four-scale gradient-loss kernels (gscream_b200.losses), the fused densification statistics (gscream_b200.stats) and the fused Adam
(gscream_b200.optim); the reference arm's optimizer is torch.optim.Adam, as in the reference.
"""
import math

import torch
import torch.nn.functional as F

import _anchor_decode as ad


def _gaussian_window(size=11, sigma=1.5, channels=3, device="cpu"):
    g = torch.tensor([math.exp(-(x - size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(size)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = (g @ g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channels, 1, size, size).contiguous().to(device)


def ssim(img1, img2, window):
    """utils/loss_utils.py:131-164 (11x11 Gaussian window, C1 = 0.01^2, C2 = 0.03^2, mean)."""
    c = img1.shape[0]
    a, b = img1.unsqueeze(0), img2.unsqueeze(0)
    mu1, mu2 = F.conv2d(a, window, padding=5, groups=c), F.conv2d(b, window, padding=5, groups=c)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = F.conv2d(a * a, window, padding=5, groups=c) - mu1_sq
    s2 = F.conv2d(b * b, window, padding=5, groups=c) - mu2_sq
    s12 = F.conv2d(a * b, window, padding=5, groups=c) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


def compute_scale_and_shift(prediction, target, mask):
    """utils/loss_utils.py:80-102 (closed-form least squares per image)."""
    a_00 = torch.sum(mask * prediction * prediction, (1, 2))
    a_01 = torch.sum(mask * prediction, (1, 2))
    a_11 = torch.sum(mask, (1, 2))
    b_0 = torch.sum(mask * prediction * target, (1, 2))
    b_1 = torch.sum(mask * target, (1, 2))
    det = a_00 * a_11 - a_01 * a_01
    x_0 = torch.where(det != 0, (a_11 * b_0 - a_01 * b_1) / det, torch.zeros_like(det))
    x_1 = torch.where(det != 0, (-a_01 * b_0 + a_00 * b_1) / det, torch.zeros_like(det))
    return x_0, x_1


def gradient_loss(prediction, target, mask):
    """train.py:232-251 with reduction_image_based (:221-230): pair-masked L1 of the first differences, per-image mask mean."""
    M = torch.sum(mask, (1, 2))
    diff = mask * (prediction - target)
    grad_x = torch.abs(diff[:, :, 1:] - diff[:, :, :-1]) * (mask[:, :, 1:] * mask[:, :, :-1])
    grad_y = torch.abs(diff[:, 1:, :] - diff[:, :-1, :]) * (mask[:, 1:, :] * mask[:, :-1, :])
    image_loss = torch.sum(grad_x, (1, 2)) + torch.sum(grad_y, (1, 2))
    return torch.mean(torch.where(M != 0, image_loss / torch.where(M != 0, M, torch.ones_like(M)), image_loss))


class TrainStep:
    def __init__(self, rast_module, decode_fn, A=100000, k=10, W=1008, H=567, seed=4, device="cuda", fused_losses=False):
        from gscream_b200 import scenes
        self.mod, self.decode_fn, self.dev = rast_module, decode_fn, torch.device(device)
        self.fused_losses = fused_losses
        self.fused_statis = fused_losses
        self.cam = scenes.make_camera(W, H)
        self.pc = ad.SyntheticAnchors(A, n_offsets=k, seed=seed, tanfov=(self.cam["tanfovx"], self.cam["tanfovy"])).to(self.dev)
        g = torch.Generator().manual_seed(seed + 1)
        self.bg = torch.zeros(3, device=self.dev)
        self.target = torch.rand(3, H, W, generator=g).to(self.dev)
        self.target_depth = (2.0 + 10.0 * torch.rand(1, H, W, generator=g)).to(self.dev)
        self.valid = torch.ones(1, H, W, device=self.dev)
        self.window = _gaussian_window(device=self.dev)
        if fused_losses:
            from gscream_b200 import optim
            self.opt = optim.Adam(self.pc.parameters(), lr=1e-4, eps=1e-15)          # one launch for all tensors (gsr_adam_step)
        else:
            self.opt = torch.optim.Adam(self.pc.parameters(), lr=1e-4, eps=1e-15)   # scene/gaussian_model.py:407
        ad.init_statis_buffers(self.pc)
        self.campos = self.cam["campos"].to(self.dev)
        self.settings = ad.make_settings(self.mod, self.cam, self.bg, self.dev)
        self.last = {}

    def step(self):
        pc, mod = self.pc, self.mod
        rast = mod.GaussianRasterizer(raster_settings=self.settings)
        with torch.no_grad():
            radii = rast.visible_filter(means3D=pc._anchor, scales=pc.get_scaling[:, :3], rotations=pc.get_rotation, cov3D_precomp=None)
            vis = radii > 0
        xyz, color, opacity, unc, scaling, rot, nop, mask = self.decode_fn(self.campos, pc, vis)
        ssp = torch.zeros_like(xyz, requires_grad=True)
        image, depth, uncer, radii = rast(means3D=xyz, means2D=ssp, shs=None, colors_precomp=color, opacities=opacity, uncertainties=unc,
                                          scales=scaling, rotations=rot, cov3D_precomp=None)
        if self.fused_losses:
            from gscream_b200 import losses
            s_mean, l1 = losses.l1_ssim(image, self.target)            # one kernel each way (gsr_l1_ssim_*)
            loss = 0.8 * l1 + 0.2 * (1.0 - s_mean)
        else:
            l1 = (image - self.target).abs().mean()
            loss = 0.8 * l1 + 0.2 * (1.0 - ssim(image, self.target, self.window))
        if self.fused_losses:
            d_l1, d_gl = losses.aligned_depth_losses(depth, self.target_depth, self.valid, None, self.valid)   # gsr_depth_align_l1_* + gsr_depth_grad_*
            loss = loss + 0.1 * d_l1 + 0.5 * 0.01 * d_gl
        else:
            s, t = compute_scale_and_shift(depth, self.target_depth, self.valid)
            aligned = s.abs().view(-1, 1, 1) * depth + t.view(-1, 1, 1)
            loss = loss + 0.1 * (aligned - self.target_depth).abs().mean()
            for sc in range(4):                                                 # train.py:556-560
                step = pow(2, sc)
                loss = loss + 0.5 * 0.01 * gradient_loss(aligned[:, ::step, ::step], self.target_depth[:, ::step, ::step], self.valid[:, ::step, ::step])
        loss = loss + 0.01 * scaling.prod(dim=1).mean()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        with torch.no_grad():                                               # train.py:597-599
            if self.fused_statis:
                from gscream_b200 import stats
                stats.training_statis(pc, ssp, nop, radii > 0, mask, vis)
            else:
                ad.training_statis_eager(pc, ssp, nop, radii > 0, mask, vis)
        self.opt.step()
        self.last = {"loss": loss.detach(), "P": xyz.shape[0], "n_vis": int(mask.shape[0] // pc.n_offsets)}
        return self.last


def torch_decode(campos, pc, vis):
    return ad.generate_neural_gaussians(campos, pc, vis)


def fused_decode(campos, pc, vis):
    from gscream_b200 import decode

    class Cam:
        camera_center = campos

    return decode.generate_neural_gaussians(Cam, pc, vis, is_training=True)
