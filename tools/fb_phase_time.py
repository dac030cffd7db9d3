#!/usr/bin/env python
"""Where the packed iteration kernel's time goes: whole kernel (variant 25) vs phase A only (28) vs phases B + C only
(29) at 4K, finest-level launches (the experiment variants produce wrong flows by construction)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import flow_cv as F
from transflow_b200 import ops, _lib
from transflow_b200.synthetic import synthetic_clip
lib = _lib.load()
h, w = 2160, 3840
clip = synthetic_clip(h, w, 2, seed=1)
a, b = (torch.from_numpy(F.gray_from_bgr(f)).cuda() for f in clip)
out = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
for v in (24, 25, 28, 29, 30, 25, 28, 29, 30):
    fb = ops.Farneback(h, w, variant=v)
    fb.prepare(0, a); fb.prepare(1, b)
    for _ in range(3): fb.solve(0, 1, out)
    torch.cuda.synchronize()
    lib.tf_timer_enable(1)
    for _ in range(10): fb.solve(0, 1, out)
    torch.cuda.synchronize()
    ms, cnt = C.c_double(), C.c_uint64()
    lib.tf_timer_read(0, C.byref(ms), C.byref(cnt)); lib.tf_timer_enable(0)
    print(f"4K variant {v}: finest iter {1e3*ms.value/max(cnt.value,1):.1f} us", flush=True)
    fb.close()
