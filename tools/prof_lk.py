#!/usr/bin/env python
"""Dense Lucas-Kanade at 1080p, a few calls (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from transflow_b200 import ops
from transflow_b200.synthetic import synthetic_clip
from oracle import flow_cv as F
h, w = int(os.environ.get("PROF_H", 1080)), int(os.environ.get("PROF_W", 1920))
clip = synthetic_clip(h, w, 2, seed=1)
a, b = (torch.from_numpy(F.gray_from_bgr(f)).cuda() for f in clip)
lk = ops.LucasKanade(h, w, 15, 2, int(os.environ.get("PROF_STEP", 1)))
out = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
for _ in range(2): lk(a, b, out=out)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
