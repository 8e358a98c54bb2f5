"""Synthetic Gaussian clouds and cameras for tests and bench.py.

Generated on the CPU from a seeded torch.Generator so every rank, both implementations and the CPU
oracle see identical bits.  Camera conventions follow the reference exactly:
world_view_transform = getWorld2View2(R, T).T and full_proj_transform = world_view @ projection.T
(scene/cameras.py:64-67, utils/graphics_utils.py:38-74 of W-Ted/GScream).
"""
import math

import torch


def projection_matrix(znear, zfar, fovx, fovy, cx=0.0, cy=0.0):
    """utils/graphics_utils.py:51-74 (note P[0,2]=cx, P[1,2]=cy, P[2,2]=(zn+zf)/(zf-zn))."""
    tan_y, tan_x = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = tan_y * znear, tan_x * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[0, 2] = cx
    P[1, 2] = cy
    P[3, 2] = 1.0
    P[2, 2] = (znear + zfar) / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(W, H, fovx_deg=60.0, yaw_deg=0.0, pitch_deg=0.0, znear=0.01, zfar=100.0):
    """Camera at the origin looking down +z, optionally yawed/pitched (used to give every rank its own view)."""
    fovx = math.radians(fovx_deg)
    tanfovx = math.tan(fovx / 2)
    tanfovy = tanfovx * H / W
    fovy = 2 * math.atan(tanfovy)
    cy_, sy_ = math.cos(math.radians(yaw_deg)), math.sin(math.radians(yaw_deg))
    cp_, sp_ = math.cos(math.radians(pitch_deg)), math.sin(math.radians(pitch_deg))
    Ry = torch.tensor([[cy_, 0, sy_], [0, 1, 0], [-sy_, 0, cy_]], dtype=torch.float64)
    Rx = torch.tensor([[1, 0, 0], [0, cp_, -sp_], [0, sp_, cp_]], dtype=torch.float64)
    w2c = torch.eye(4, dtype=torch.float64)
    w2c[:3, :3] = (Ry @ Rx).T  # world -> camera rotation
    world_view = w2c.float().transpose(0, 1).contiguous()
    proj = projection_matrix(znear, zfar, fovx, fovy).transpose(0, 1)
    full_proj = (world_view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    campos = world_view.inverse()[3, :3].contiguous()
    return dict(W=W, H=H, tanfovx=tanfovx, tanfovy=tanfovy, viewmatrix=world_view, projmatrix=full_proj, campos=campos)


def make_scene(P, W, H, C, seed, scale_mult=1.0, fovx_deg=60.0, frac_behind=0.02, bg_value=0.0):
    """SURVEY.md section 8d synthetic cloud.  Returns CPU float32 tensors."""
    g = torch.Generator().manual_seed(int(seed))
    tanfovx = math.tan(math.radians(fovx_deg) / 2)
    tanfovy = tanfovx * H / W
    f_px = W / (2 * tanfovx)
    n_behind = int(round(P * frac_behind))
    z = torch.empty(P).uniform_(2.0, 12.0, generator=g)
    if n_behind > 0:
        z[:n_behind] = torch.empty(n_behind).uniform_(-1.0, 0.2, generator=g)
    perm = torch.randperm(P, generator=g)
    z = z[perm]
    ux = torch.empty(P).uniform_(-1.05, 1.05, generator=g)
    uy = torch.empty(P).uniform_(-1.05, 1.05, generator=g)
    x = z.abs().clamp_min(0.2) * tanfovx * ux
    y = z.abs().clamp_min(0.2) * tanfovy * uy
    means3D = torch.stack([x, y, z], 1).contiguous()
    s0 = 2.0 * 7.0 / f_px * scale_mult
    scales = torch.exp(torch.randn(P, 3, generator=g) * 0.6 + math.log(s0)).contiguous()
    rot = torch.randn(P, 4, generator=g)
    rotations = (rot / rot.norm(dim=1, keepdim=True)).contiguous()
    opacities = torch.sigmoid(torch.randn(P, 1, generator=g) * 1.5).contiguous()
    uncertainties = torch.rand(P, 1, generator=g).contiguous()
    if C == 3:
        colors = torch.rand(P, 3, generator=g).contiguous()
    else:
        colors = torch.randn(P, C, generator=g).contiguous()
    bg = torch.full((C,), float(bg_value))
    return dict(means3D=means3D, scales=scales, rotations=rotations, opacities=opacities,
                uncertainties=uncertainties, colors=colors, bg=bg)


def make_upstream_grads(C, W, H, seed):
    """Upstream gradients U[-1,1]/N for every output plane (mimics mean-reduced losses)."""
    g = torch.Generator().manual_seed(int(seed) + 7919)
    N = float(W * H)
    gc = (torch.rand(C, H, W, generator=g) * 2 - 1) / N
    gd = (torch.rand(1, H, W, generator=g) * 2 - 1) / N
    gu = (torch.rand(1, H, W, generator=g) * 2 - 1) / N
    return gc.contiguous(), gd.contiguous(), gu.contiguous()


# named workloads (BASELINE.json configs 2 and 3)
CONFIGS = {
    "config2": dict(P=500_000, W=1008, H=567, C=3, seed=20240417 + 2),
    "config3": dict(P=1_000_000, W=1920, H=1080, C=32, seed=20240417 + 3),
}
