#!/usr/bin/env python
"""Farneback solve-kernel sweep: accuracy vs cv2 (small frames) and per-iteration timing at 4K for every
fused variant and a range of rows-per-CTA settings.  Writes gpurun_out/fb_sweep.txt."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import flow_cv as F
from transflow_b200 import ops, _lib
from transflow_b200.synthetic import synthetic_clip
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
OUT = open(os.path.join(ROOT, "gpurun_out", "fb_sweep.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); OUT.write(s + "\n"); OUT.flush()
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
lib = _lib.load()
VARIANTS = [int(v) for v in os.environ.get("SWEEP_VARIANTS", "3,4,5,6").split(",")]
ROWS = [int(v) for v in os.environ.get("SWEEP_ROWS", "0,56,112,168,224,252,336,448").split(",")]
try:
    for (h, w) in ((135, 201), (480, 854), (1080, 1920)):
        clip = synthetic_clip(h, w, 2, seed=2)
        g0, g1 = F.gray_from_bgr(clip[0]), F.gray_from_bgr(clip[1])
        for params in ({}, dict(winsize=9), dict(winsize=33, poly_n=7, poly_sigma=1.5), dict(pyr_scale=0.7, levels=5, winsize=21)):
            want = F.farneback(g0, g1, **params)
            for v in VARIANTS:
                got = ops.Farneback(h, w, variant=v, **params)(dev(g0), dev(g1)).cpu().numpy()
                e = np.linalg.norm(got - want, axis=-1)
                P(f"acc {w}x{h} {params} variant {v}: mean {e.mean():.2e} max {e.max():.2e}")
except Exception:
    P("ACC FAILED\n" + traceback.format_exc())
try:
    h, w = 2160, 3840
    clip = synthetic_clip(h, w, 2, seed=1)
    a, b = (dev(F.gray_from_bgr(f)) for f in clip)
    out = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
    ref = None
    for v in VARIANTS:
        fb = ops.Farneback(h, w, variant=v)
        fb.prepare(0, a); fb.prepare(1, b)
        for rows in (ROWS if v >= 4 else [0]):
            lib.tf_farneback_tune(0, rows)
            for _ in range(3): fb.solve(0, 1, out)
            torch.cuda.synchronize()
            lib.tf_timer_enable(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 10
            for _ in range(n): fb.solve(0, 1, out)
            e1.record(); torch.cuda.synchronize()
            import ctypes as C
            ms, cnt = C.c_double(), C.c_uint64()
            lib.tf_timer_read(0, C.byref(ms), C.byref(cnt))
            lib.tf_timer_enable(0)
            o = out.cpu().numpy()
            if ref is None: ref = o
            d = np.linalg.norm(o - ref, axis=-1)
            P(f"4K variant {v} rows {rows}: solve {e0.elapsed_time(e1)/n:.3f} ms; finest iter {1e3*ms.value/max(cnt.value,1):.1f} us x{cnt.value}; vs variant {VARIANTS[0]}: max {d.max():.2e}")
        lib.tf_farneback_tune(0, 0)
        fb.close()
except Exception:
    P("TIMING FAILED\n" + traceback.format_exc())
