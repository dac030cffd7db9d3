#!/usr/bin/env python
"""Ring-kernel sweep: bit-identity against the default kernel (small + 4K frames, small and large motion) and
per-launch timing of the finest-level iteration at 4K / 1080p for the ring variants and rows-per-CTA settings.
Writes gpurun_out/ring_sweep.txt."""
import ctypes as C
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import flow_cv as F
from transflow_b200 import ops, _lib
from transflow_b200.synthetic import synthetic_clip
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
OUT = open(os.path.join(ROOT, "gpurun_out", os.environ.get("SWEEP_OUT", "ring_sweep.txt")), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); OUT.write(s + "\n"); OUT.flush()
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
lib = _lib.load()
BASE = int(os.environ.get("SWEEP_BASE", "8"))
VARIANTS = [int(v) for v in os.environ.get("SWEEP_VARIANTS", "18,19,20").split(",")]
ROWS = [int(v) for v in os.environ.get("SWEEP_ROWS", "0,84,112,140,168,224,252,280").split(",")]
KEY = {v: 3 for v in (18, 19, 20, 21, 22)}   # others: key 0 (half-buffer kernel rows)
try:
    for (h, w) in ((540, 960), (1080, 1920)):
        clip = synthetic_clip(h, w, 2, seed=2)
        g0, g1 = F.gray_from_bgr(clip[0]), F.gray_from_bgr(clip[1])
        for name, right in (("clip", g1), ("roll", np.roll(g0, (9, 13), (0, 1))), ("noise", np.random.default_rng(0).integers(0, 255, g0.shape, dtype=np.uint8))):
            want = ops.Farneback(h, w, variant=BASE)(dev(g0), dev(right)).cpu().numpy()
            for v in VARIANTS:
                got = ops.Farneback(h, w, variant=v)(dev(g0), dev(right)).cpu().numpy()
                d = np.abs(got - want).max()
                P(f"ident {w}x{h} {name} variant {v} vs {BASE}: max abs diff {d:.3e} {'OK' if d == 0 else 'MISMATCH'}")
except Exception:
    P("IDENT FAILED\n" + traceback.format_exc())
for (h, w) in ((2160, 3840), (1080, 1920)):
    try:
        clip = synthetic_clip(h, w, 2, seed=1)
        a, b = (dev(F.gray_from_bgr(f)) for f in clip)
        out = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
        ref = None
        for v in [BASE] + VARIANTS:
            fb = ops.Farneback(h, w, variant=v)
            fb.prepare(0, a); fb.prepare(1, b)
            for rows in (ROWS if v != BASE else [0]):
                lib.tf_farneback_tune(KEY.get(v, 0), rows)
                for _ in range(3): fb.solve(0, 1, out)
                torch.cuda.synchronize()
                lib.tf_timer_enable(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                n = 10
                for _ in range(n): fb.solve(0, 1, out)
                e1.record(); torch.cuda.synchronize()
                ms, cnt = C.c_double(), C.c_uint64()
                lib.tf_timer_read(0, C.byref(ms), C.byref(cnt))
                lib.tf_timer_enable(0)
                o = out.cpu().numpy()
                if ref is None: ref = o
                d = np.abs(o - ref).max()
                P(f"{w}x{h} variant {v} rows {rows}: solve {e0.elapsed_time(e1)/n:.3f} ms; finest iter {1e3*ms.value/max(cnt.value,1):.1f} us x{cnt.value}; vs base max {d:.2e}")
            lib.tf_farneback_tune(KEY.get(v, 0), 0)
            fb.close()
    except Exception:
        P(f"TIMING {w}x{h} FAILED\n" + traceback.format_exc())
