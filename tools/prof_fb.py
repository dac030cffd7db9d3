#!/usr/bin/env python
"""Minimal Farneback workload for ncu: a few 4K prepare+solve calls (variant from TFB200_FB_VARIANT)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from transflow_b200 import ops
from transflow_b200.synthetic import synthetic_clip
from oracle import flow_cv as F
h, w = int(os.environ.get("PROF_H", 2160)), int(os.environ.get("PROF_W", 3840))
clip = synthetic_clip(h, w, 2, seed=1)
a, b = (torch.from_numpy(F.gray_from_bgr(f)).cuda() for f in clip)
fb = ops.Farneback(h, w)
if os.environ.get("PROF_ROWS"):
    fb.lib.tf_farneback_tune(0, int(os.environ["PROF_ROWS"]))
out = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
for _ in range(int(os.environ.get("PROF_REPS", 3))):
    fb.prepare(0, a); fb.prepare(1, b); fb.solve(0, 1, out)
torch.cuda.synchronize()
print("done", float(out.abs().mean()))
