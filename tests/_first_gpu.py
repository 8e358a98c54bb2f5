import sys, os, time, numpy as np, torch
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests')
import _ref_utils as ru
from gscream_b200 import scenes, rasterizer as ours, _C
def cmp(name, a, b, exact=False):
    a = np.asarray(a); b = np.asarray(b)
    if exact:
        n = int((a != b).sum()); print(f'  {name:18s} exact mismatches {n}/{a.size}'); return n
    d = np.abs(a.astype(np.float64)-b.astype(np.float64)); rel = d/np.maximum(np.abs(b.astype(np.float64)), 1e-30)
    scale = np.abs(b).max()
    print(f'  {name:18s} maxabs {d.max():.3e} (scale {scale:.3e}) rel-to-scale {d.max()/max(scale,1e-30):.3e} p99.9 rel {np.quantile(rel,0.999):.2e}')
def run_case(P,W,H,C,seed,smult,bgv,yaw, timing=0):
    print(f'== P={P} {W}x{H} C={C} seed={seed} smult={smult}')
    scene = scenes.make_scene(P,W,H,C,seed,scale_mult=smult,bg_value=bgv); cam = scenes.make_camera(W,H,yaw_deg=yaw); grads = scenes.make_upstream_grads(C,W,H,seed)
    ref = ru.load_ref(C)
    r = ru.run_impl(ref, scene, cam, grads)
    m = ru.run_impl(ours, scene, cam, grads)
    R = r['num_rendered']; print('  R ref/ours', R, m['num_rendered'], 'visible', int((r['radii']>0).sum()))
    g = ru.parse_ref_geom(r['_geom'].cpu().numpy(), P); im = ru.parse_ref_image(r['_img'].cpu().numpy(), W, H); bn = ru.parse_ref_binning(r['_binning'].cpu().numpy(), R)
    e = _C.debug_export(P, m['num_rendered'], W, H, m['_geom'], m['_binning'], m['_img']); e = {k:v.cpu().numpy() for k,v in e.items()}
    vis = r['radii']>0
    cmp('radii', m['radii'], r['radii'], True)
    cmp('tiles_touched', e['tiles_touched'].view(np.uint32), g['tiles_touched'], True)
    cmp('xy bits', e['xy'][vis].view(np.uint32), g['means2D'][vis].view(np.uint32), True)
    cmp('depth bits', e['depths'][vis].view(np.uint32), g['depths'][vis].view(np.uint32), True)
    cmp('conic bits', e['conic_opacity'][vis].view(np.uint32), g['conic_opacity'][vis].view(np.uint32), True)
    if R == m['num_rendered']:
        cmp('point_list', e['point_list'].view(np.uint32), bn['point_list'], True)
    cmp('ranges', e['ranges'].view(np.uint32), im['ranges'], True)
    cmp('n_contrib', e['n_contrib'].view(np.uint32), im['n_contrib'], True)
    cmp('final_T', e['final_T'], im['final_T'])
    for k in ('color','depth','uncertainty','dL_dmeans3D','dL_dmeans2D','dL_dcolors','dL_dopacity','dL_duncertainty','dL_dscales','dL_drotations'):
        cmp(k, m[k], r[k])
    if timing:
        dev = torch.device('cuda')
        for name, mod in (('ref', ref), ('ours', ours)):
            t = {k: v.to(dev) for k, v in scene.items()}
            gc, gd, gu = (x.to(dev) for x in grads)
            st = mod.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam['tanfovx'], tanfovy=cam['tanfovy'], bg=t['bg'], scale_modifier=1.0, viewmatrix=cam['viewmatrix'].to(dev), projmatrix=cam['projmatrix'].to(dev), sh_degree=1, campos=cam['campos'].to(dev), prefiltered=False, debug=False)
            rast = mod.GaussianRasterizer(st)
            leaves = [t[k].clone().requires_grad_(True) for k in ('means3D','colors','opacities','uncertainties','scales','rotations')]
            m2d = torch.zeros_like(leaves[0], requires_grad=True)
            def step(bwd=True):
                c,d,u,_ = rast(means3D=leaves[0], means2D=m2d, opacities=leaves[2], uncertainties=leaves[3], shs=None, colors_precomp=leaves[1], scales=leaves[4], rotations=leaves[5], cov3D_precomp=None)
                if bwd: torch.autograd.backward((c,d,u),(gc,gd,gu))
            for _ in range(3): step()
            torch.cuda.synchronize()
            for bwd in (False, True):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                ts=[]
                for _ in range(timing):
                    e0.record(); step(bwd); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
                print(f'  TIMING {name} {"fwd+bwd" if bwd else "fwd    "}: median {np.median(ts):.3f} ms  min {min(ts):.3f}')
if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    run_case(1500,160,96,3,11,2.0,0.25,0.0)
    run_case(1200,150,83,3,12,4.0,1.0,7.0)
    run_case(1200,128,80,32,13,2.0,0.1,0.0)
    run_case(2500,96,64,32,14,5.0,0.0,-5.0)
    run_case(100000,1008,567,3,21,1.0,0.0,0.0, timing=5)
    run_case(500000,1008,567,3,20240419,1.0,0.0,0.0, timing=10)
    run_case(1000000,1920,1080,32,20240420,1.0,0.0,0.0, timing=10)
