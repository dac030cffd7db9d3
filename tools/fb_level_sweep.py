#!/usr/bin/env python
"""Per-level timing of the Farneback iteration kernels: a single-level solve (levels=0, 3 iterations) at the sizes
of the 4K pyramid, for every fused variant and rows-per-CTA setting.  Writes gpurun_out/fb_level_sweep.txt."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import flow_cv as F
from transflow_b200 import ops, _lib
from transflow_b200.synthetic import synthetic_clip
OUT = open(os.path.join(ROOT, "gpurun_out", "fb_level_sweep.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); OUT.write(s + "\n"); OUT.flush()
lib = _lib.load()
VARIANTS = [int(v) for v in os.environ.get("SWEEP_VARIANTS", "3,4,6").split(",")]
ROWS = [int(v) for v in os.environ.get("SWEEP_ROWS", "0,28,42,56,84,112,140,168").split(",")]
for (h, w) in ((2160, 3840), (1080, 1920), (540, 960), (270, 480)):
    clip = synthetic_clip(h, w, 2, seed=1)
    a, b = (torch.from_numpy(F.gray_from_bgr(f)).cuda() for f in clip)
    out = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
    for v in VARIANTS:
        fb = ops.Farneback(h, w, levels=0, variant=v)
        fb.prepare(0, a); fb.prepare(1, b)
        res = []
        for rows in (ROWS if v >= 4 else [0]):
            if rows > h: continue
            lib.tf_farneback_tune(0, rows)
            for _ in range(3): fb.solve(0, 1, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 20
            for _ in range(n): fb.solve(0, 1, out)
            e1.record(); torch.cuda.synchronize()
            res.append(f"{rows}:{e0.elapsed_time(e1)/n/3*1e3:.1f}")
        lib.tf_farneback_tune(0, 0)
        fb.close()
        P(f"{w}x{h} variant {v} us/iter by rows  " + "  ".join(res))
