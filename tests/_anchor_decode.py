"""TEST INFRASTRUCTURE — torch restatement of the caller side of the hot path:
GScream's anchor -> neural-Gaussian decode (`generate_neural_gaussians`, gaussian_renderer/__init__.py:18-102) and the
anchor prefilter (`prefilter_position2D`, :248-302), on a synthetic anchor model with the MLP shapes of
scene/gaussian_model.py:118-144.  The reference's own GaussianModel hard-codes .cuda() and imports packages absent here,
so BASELINE.json configs[0] ("1k anchors on CPU torch, plumbing only") is exercised through this restatement.
`rasterizer_module` is either this repo's drop-in or the reference build: the glue is identical for both."""
import math

import torch
import torch.nn as nn


class SyntheticAnchors(nn.Module):
    def __init__(self, A, feat_dim=32, n_offsets=10, seed=0, extent=(2.0, 12.0), tanfov=(0.577, 0.325)):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.n_offsets, self.feat_dim = n_offsets, feat_dim
        z = torch.empty(A).uniform_(*extent, generator=g)
        x = z * tanfov[0] * torch.empty(A).uniform_(-1.05, 1.05, generator=g)
        y = z * tanfov[1] * torch.empty(A).uniform_(-1.05, 1.05, generator=g)
        self._anchor = nn.Parameter(torch.stack([x, y, z], 1))
        self._offset = nn.Parameter(torch.randn(A, n_offsets, 3, generator=g) * 0.5)
        self._anchor_feat = nn.Parameter(torch.randn(A, feat_dim, generator=g) * 0.5)
        self._scaling = nn.Parameter(torch.log(torch.full((A, 6), 0.03) * torch.exp(torch.randn(A, 6, generator=g) * 0.3)))
        self._rotation = nn.Parameter(torch.tensor([[1.0, 0, 0, 0]]).repeat(A, 1))
        torch.manual_seed(seed)
        d = feat_dim + 3 + 1
        self.mlp_opacity = nn.Sequential(nn.Linear(d, feat_dim), nn.ReLU(True), nn.Linear(feat_dim, n_offsets), nn.Tanh())
        self.mlp_uncertainty = nn.Sequential(nn.Linear(d, feat_dim), nn.ReLU(True), nn.Linear(feat_dim, n_offsets), nn.Sigmoid())
        self.mlp_cov = nn.Sequential(nn.Linear(d, feat_dim), nn.ReLU(True), nn.Linear(feat_dim, 7 * n_offsets))
        self.mlp_color = nn.Sequential(nn.Linear(d, feat_dim), nn.ReLU(True), nn.Linear(feat_dim, 3 * n_offsets), nn.Sigmoid())

    # attribute names of scene/gaussian_model.py:229-272, so the reference's own function accepts this object
    use_feat_bank = False
    rotation_activation = staticmethod(torch.nn.functional.normalize)

    @property
    def get_anchor(self):
        return self._anchor

    @property
    def get_opacity_mlp(self):
        return self.mlp_opacity

    @property
    def get_uncertainty_mlp(self):
        return self.mlp_uncertainty

    @property
    def get_cov_mlp(self):
        return self.mlp_cov

    @property
    def get_color_mlp(self):
        return self.mlp_color

    @property
    def get_scaling(self):
        return 1.0 * torch.exp(self._scaling)          # scene/gaussian_model.py:241-242

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)


def generate_neural_gaussians(camera_center, pc, visible_mask=None):
    """gaussian_renderer/__init__.py:18-102 (use_feat_bank=False, is_training=True)."""
    if visible_mask is None:
        visible_mask = torch.ones(pc._anchor.shape[0], dtype=torch.bool, device=pc._anchor.device)
    feat = pc._anchor_feat[visible_mask]
    anchor = pc._anchor[visible_mask]
    grid_offsets = pc._offset[visible_mask]
    grid_scaling = pc.get_scaling[visible_mask]
    ob_view = anchor - camera_center
    ob_dist = ob_view.norm(dim=1, keepdim=True)
    ob_view = ob_view / ob_dist
    cat_local_view = torch.cat([feat, ob_view, ob_dist], dim=1)
    neural_opacity = pc.mlp_opacity(cat_local_view).reshape([-1, 1])
    mask = (neural_opacity > 0.0).view(-1)
    opacity = neural_opacity[mask]
    n = anchor.shape[0] * pc.n_offsets
    uncertainty = pc.mlp_uncertainty(cat_local_view).reshape([n, 1])
    color = pc.mlp_color(cat_local_view).reshape([n, 3])
    scale_rot = pc.mlp_cov(cat_local_view).reshape([n, 7])
    offsets = grid_offsets.view([-1, 3])
    concatenated = torch.cat([grid_scaling, anchor], dim=-1)
    concatenated_repeated = concatenated.repeat_interleave(pc.n_offsets, dim=0)   # einops 'n c -> (n k) c'
    concatenated_all = torch.cat([concatenated_repeated, uncertainty, color, scale_rot, offsets], dim=-1)
    masked = concatenated_all[mask]
    scaling_repeat, repeat_anchor, uncertainty, color, scale_rot, offsets = masked.split([6, 3, 1, 3, 7, 3], dim=-1)
    scaling = scaling_repeat[:, 3:] * torch.sigmoid(scale_rot[:, :3])
    rot = torch.nn.functional.normalize(scale_rot[:, 3:7])
    xyz = repeat_anchor + offsets * scaling_repeat[:, :3]
    return xyz, color, opacity, uncertainty, scaling, rot, neural_opacity, mask


def init_statis_buffers(pc):
    """scene/gaussian_model.py:84-95,316-320: the four densification-statistics buffers."""
    A, k, dev = pc._anchor.shape[0], pc.n_offsets, pc._anchor.device
    pc.opacity_accum = torch.zeros((A, 1), device=dev)
    pc.anchor_demon = torch.zeros((A, 1), device=dev)
    pc.offset_gradient_accum = torch.zeros((A * k, 1), device=dev)
    pc.offset_denom = torch.zeros((A * k, 1), device=dev)


def training_statis_eager(pc, viewspace_point_tensor, opacity, update_filter, offset_selection_mask, anchor_visible_mask):
    """scene/gaussian_model.py:729-757 (GaussianModel.training_statis), the eager index code of the reference."""
    temp_opacity = opacity.clone().view(-1).detach()
    temp_opacity[temp_opacity < 0] = 0
    temp_opacity = temp_opacity.view([-1, pc.n_offsets])
    pc.opacity_accum[anchor_visible_mask] += temp_opacity.sum(dim=1, keepdim=True)
    pc.anchor_demon[anchor_visible_mask] += 1
    anchor_visible_mask = anchor_visible_mask.unsqueeze(dim=1).repeat([1, pc.n_offsets]).view(-1)
    combined_mask = torch.zeros_like(pc.offset_gradient_accum, dtype=torch.bool).squeeze(dim=1)
    combined_mask[anchor_visible_mask] = offset_selection_mask
    temp_mask = combined_mask.clone()
    combined_mask[temp_mask] = update_filter
    grad_norm = torch.norm(viewspace_point_tensor.grad[update_filter, :2], dim=-1, keepdim=True)
    pc.offset_gradient_accum[combined_mask] += grad_norm
    pc.offset_denom[combined_mask] += 1


def make_settings(mod, cam, bg, device):
    return mod.GaussianRasterizationSettings(
        image_height=cam["H"], image_width=cam["W"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=bg, scale_modifier=1.0,
        viewmatrix=cam["viewmatrix"].to(device), projmatrix=cam["projmatrix"].to(device), sh_degree=1, campos=cam["campos"].to(device),
        prefiltered=False, debug=False)


def prefilter_position2D(mod, cam, pc, bg):
    """gaussian_renderer/__init__.py:248-302: anchors, scales[:, :3] (a non-contiguous slice), rotations."""
    rast = mod.GaussianRasterizer(raster_settings=make_settings(mod, cam, bg, pc._anchor.device))
    radii, x, y = rast.position2D_filter(means3D=pc._anchor, scales=pc.get_scaling[:, :3], rotations=pc.get_rotation, cov3D_precomp=None)
    return radii > 0, x, y


def render(mod, cam, pc, bg, visible_mask):
    """gaussian_renderer/__init__.py:104-179 (training branch)."""
    dev = pc._anchor.device
    xyz, color, opacity, uncertainty, scaling, rot, neural_opacity, mask = generate_neural_gaussians(cam["campos"].to(dev), pc, visible_mask)
    screenspace_points = torch.zeros_like(xyz, requires_grad=True) + 0
    screenspace_points.retain_grad()
    rast = mod.GaussianRasterizer(raster_settings=make_settings(mod, cam, bg, dev))
    image, depth, uncer, radii = rast(means3D=xyz, means2D=screenspace_points, shs=None, colors_precomp=color, opacities=opacity,
                                      uncertainties=uncertainty, scales=scaling, rotations=rot, cov3D_precomp=None)
    return dict(render=image, render_depth=depth, uncertainty=uncer, viewspace_points=screenspace_points, visibility_filter=radii > 0,
                radii=radii, selection_mask=mask, neural_opacity=neural_opacity, scaling=scaling)
