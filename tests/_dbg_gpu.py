import sys, os, numpy as np, torch
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests')
import _ref_utils as ru
from _cases import *
from oracle.oracle import Oracle
from gscream_b200 import scenes, rasterizer as ours
g = load_golden('c3_small')
P,W,H,C = int(g['P']),int(g['W']),int(g['H']),int(g['C'])
scene = {k[3:]: torch.from_numpy(g[k]) for k in g if k.startswith('in_')}
cam = dict(W=W,H=H,tanfovx=float(g['cam_tanfovx']),tanfovy=float(g['cam_tanfovy']),viewmatrix=torch.from_numpy(g['cam_viewmatrix']),projmatrix=torch.from_numpy(g['cam_projmatrix']),campos=torch.from_numpy(g['cam_campos']))
ref = ru.load_ref(C)
o = Oracle('f64'); f = oracle_forward(o, g)
z = lambda k: np.zeros_like(g[k])
for name, sel in (('all',(1,1,1)),('color-only',(1,0,0)),('depth-only',(0,1,0)),('unc-only',(0,0,1))):
    gc = g['g_color']*sel[0]; gd = g['g_depth']*sel[1]; gu = g['g_unc']*sel[2]
    grads = tuple(torch.from_numpy(np.ascontiguousarray(x)) for x in (gc,gd,gu))
    r = ru.run_impl(ref, scene, cam, grads); m = ru.run_impl(ours, scene, cam, grads)
    gg = dict(g); gg['g_color']=gc; gg['g_depth']=gd; gg['g_unc']=gu
    b = oracle_backward(o, f, gg)
    for k,kk in (('dL_dopacity','dL_dopacity'),('dL_dmeans2D','dL_dmean2D'),('dL_dcolors','dL_dcolors'),('dL_duncertainty','dL_duncertainty')):
        ov = b[kk].reshape(r[k].shape)
        print(f'{name:11s} {k:16s} |ref-orc| {np.abs(r[k]-ov).max():.3e}  |ours-ref| {np.abs(m[k]-r[k]).max():.3e}  scale {np.abs(r[k]).max():.3e}')
