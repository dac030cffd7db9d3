#!/usr/bin/env python
"""Accuracy of the staged Farneback kernel (variant 9) vs cv2 and vs the default (variant 8), small and mid sizes,
including flows larger than the staging margin; then 4K timing of 8 vs 9."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import flow_cv as F
from transflow_b200 import ops, _lib
from transflow_b200.synthetic import synthetic_clip
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
for (h, w) in ((540, 960), (1080, 1920), (600, 1000)):
    clip = synthetic_clip(h, w, 3, seed=2)
    g = [F.gray_from_bgr(f) for f in clip]
    for (a, b, tag) in ((g[0], g[1], "1 frame apart"), (g[0], g[2], "2 frames apart"), (g[0], np.roll(g[0], (9, 13), (0, 1)), "shift 13,9")):
        want = F.farneback(a, b)
        for v in (8, 9):
            got = ops.Farneback(h, w, variant=v)(dev(a), dev(b)).cpu().numpy()
            e = np.linalg.norm(got - want, axis=-1)
            print(f"acc {w}x{h} {tag} variant {v}: mean {e.mean():.2e} max {e.max():.2e}", flush=True)
lib = _lib.load()
h, w = 2160, 3840
clip = synthetic_clip(h, w, 2, seed=1)
a, b = (dev(F.gray_from_bgr(f)) for f in clip)
out = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
import ctypes as C
for v in (8, 9, 8, 9):
    fb = ops.Farneback(h, w, variant=v)
    fb.prepare(0, a); fb.prepare(1, b)
    for rows in (0, 56, 112, 168, 224):
        lib.tf_farneback_tune(0, rows)
        for _ in range(3): fb.solve(0, 1, out)
        torch.cuda.synchronize()
        lib.tf_timer_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fb.solve(0, 1, out)
        e1.record(); torch.cuda.synchronize()
        ms, cnt = C.c_double(), C.c_uint64()
        lib.tf_timer_read(0, C.byref(ms), C.byref(cnt)); lib.tf_timer_enable(0)
        print(f"4K variant {v} rows {rows}: solve {e0.elapsed_time(e1)/10:.3f} ms; finest iter {1e3*ms.value/max(cnt.value,1):.1f} us", flush=True)
    lib.tf_farneback_tune(0, 0)
    fb.close()
