"""Scratch driver for ncu: N fwd+bwd passes of the named config through the public API."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gscream_b200 import scenes, rasterizer as ours
name = sys.argv[1] if len(sys.argv) > 1 else 'config3'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = dict(scenes.CONFIGS[name]); smult = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
P,W,H,C,seed = cfg['P'],cfg['W'],cfg['H'],cfg['C'],cfg['seed']
scene = scenes.make_scene(P,W,H,C,seed,scale_mult=smult); cam = scenes.make_camera(W,H); grads = scenes.make_upstream_grads(C,W,H,seed)
dev = torch.device('cuda')
t = {k: v.to(dev) for k, v in scene.items()}
gc, gd, gu = (x.to(dev) for x in grads)
st = ours.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam['tanfovx'], tanfovy=cam['tanfovy'], bg=t['bg'], scale_modifier=1.0, viewmatrix=cam['viewmatrix'].to(dev), projmatrix=cam['projmatrix'].to(dev), sh_degree=1, campos=cam['campos'].to(dev), prefiltered=False, debug=False)
rast = ours.GaussianRasterizer(st)
leaves = [t[k].clone().requires_grad_(True) for k in ('means3D','colors','opacities','uncertainties','scales','rotations')]
m2d = torch.zeros_like(leaves[0], requires_grad=True)
for _ in range(iters):
    c,d,u,_r = rast(means3D=leaves[0], means2D=m2d, opacities=leaves[2], uncertainties=leaves[3], shs=None, colors_precomp=leaves[1], scales=leaves[4], rotations=leaves[5], cov3D_precomp=None)
    torch.autograd.backward((c,d,u),(gc,gd,gu))
torch.cuda.synchronize()
print('done', name, 'R', c.grad_fn.num_rendered if c.grad_fn is not None else None)
