"""Scratch: per-iteration wall time of the train-step loop (where do one-off stalls sit?)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _train_step as ts
from gscream_b200 import rasterizer as mod
loop = ts.TrainStep(mod, ts.fused_decode, device=torch.device("cuda"), fused_losses=True)
times = []
for i in range(45):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loop.step()
    torch.cuda.synchronize(); times.append((time.perf_counter() - t0) * 1e3)
print("first 8:", [round(t, 1) for t in times[:8]])
top = sorted(range(8, 45), key=lambda i: -times[i])[:5]
print("slowest after 8:", [(i, round(times[i], 1)) for i in top], "median", round(sorted(times[8:])[18], 2))
