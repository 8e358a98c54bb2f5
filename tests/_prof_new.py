"""Scratch driver for ncu: a few iterations of the train-step-equivalent loop (tests/_train_step.py, this repo's arm), so that
`-k regex:adam_step|depth_grad|depth_l1|depth_fit` captures the session-3 kernels at the train-step size (567x1008, 7.5 M parameters)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _train_step as ts
from gscream_b200 import rasterizer as mod
loop = ts.TrainStep(mod, ts.fused_decode, device=torch.device("cuda"), fused_losses=True)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    loop.step()
torch.cuda.synchronize()
print("done", loop.last["P"])
