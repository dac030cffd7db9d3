#!/usr/bin/env python
"""Packed half-buffer kernel (variants 25 / 26) against the default (8): bit-identity on small, ragged and large-motion
cases, then the 4K / 1080p / 8K per-launch timing of the finest level.  usage: tools/fb_pack_check.py [variants...]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import flow_cv as F
from transflow_b200 import ops, _lib
from transflow_b200.synthetic import synthetic_clip
VARS = [int(v) for v in sys.argv[1:]] or [25, 26]
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
ok = True
for (h, w), kw in (((540, 960), {}), ((603, 812), {}), ((720, 1282), {}), ((1000, 1284), dict(winsize=21)),
                   ((600, 1000), dict(winsize=9, levels=2)), ((1080, 1920), {})):
    clip = synthetic_clip(h, w, 2, seed=3)
    g0, g1 = (F.gray_from_bgr(f) for f in clip)
    for right, tag in ((g1, "clip"), (np.roll(g0, (-11, 17), (0, 1)), "shift")):
        want = ops.Farneback(h, w, variant=8, **kw)(dev(g0), dev(right))
        for v in VARS:
            fb = ops.Farneback(h, w, variant=v, **kw)
            same = all(torch.equal(fb(dev(g0), dev(right)), want) for _ in range(2))
            ok &= same
            if not same:
                d = (fb(dev(g0), dev(right)) - want).abs()
                print(f"MISMATCH {w}x{h} {kw} {tag} variant {v}: max {d.max().item():.3e}, {int((d > 0).sum())} values", flush=True)
    print(f"identity {w}x{h} {kw}: {'ok' if ok else 'FAILED'}", flush=True)
lib = _lib.load()
for (h, w) in ((2160, 3840), (1080, 1920), (4320, 7680)):
    clip = synthetic_clip(h, w, 2, seed=1)
    a, b = (dev(F.gray_from_bgr(f)) for f in clip)
    out = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
    for v in [24] + VARS + [24] + VARS:
        fb = ops.Farneback(h, w, variant=v)
        fb.prepare(0, a); fb.prepare(1, b)
        for _ in range(3): fb.solve(0, 1, out)
        torch.cuda.synchronize()
        lib.tf_timer_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fb.solve(0, 1, out)
        e1.record(); torch.cuda.synchronize()
        ms, cnt = C.c_double(), C.c_uint64()
        lib.tf_timer_read(0, C.byref(ms), C.byref(cnt)); lib.tf_timer_enable(0)
        print(f"{w}x{h} variant {v}: solve {e0.elapsed_time(e1)/10:.3f} ms; finest iter {1e3*ms.value/max(cnt.value,1):.1f} us", flush=True)
        fb.close()
    del a, b, out
print("ALL IDENTICAL" if ok else "IDENTITY FAILED")
