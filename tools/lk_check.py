#!/usr/bin/env python
"""Lucas-Kanade tracker variants: bit-identity between the column-mapped kernel (TFB200_LK_VARIANT=0) and the
linear-mapped warp kernel (=2), parity against cv2 at small sizes, timing at 4K (dense and step 4).
Writes gpurun_out/lk_check.txt."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import flow_cv as F
from transflow_b200 import ops
from transflow_b200.synthetic import synthetic_clip
OUT = open(os.path.join(ROOT, "gpurun_out", "lk_check.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); OUT.write(s + "\n"); OUT.flush()
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
def make(h, w, win, lvl, step, variant):
    os.environ["TFB200_LK_VARIANT"] = str(variant)
    return ops.LucasKanade(h, w, win, lvl, step)
for (h, w, win, lvl, step) in ((96, 128, 15, 2, 1), (120, 160, 15, 2, 4), (67, 93, 9, 3, 2), (270, 480, 15, 2, 1), (40, 52, 15, 2, 1)):
    clip = synthetic_clip(h, w, 2, seed=7)
    g0, g1 = F.gray_from_bgr(clip[0]), F.gray_from_bgr(clip[1])
    want = F.lucas_kanade(g0, g1, win, lvl, step)
    outs = {}
    for v in (0, 2):
        outs[v] = make(h, w, win, lvl, step, v)(dev(g0), dev(g1)).cpu().numpy()
    e = np.linalg.norm(outs[0] - want, axis=-1)
    P(f"{w}x{h} win {win} lvl {lvl} step {step}: new vs old max abs {np.abs(outs[0]-outs[2]).max():.3e}; vs cv2 mean {e.mean():.2e} max {e.max():.2e} frac>1e-3 {(e>1e-3).mean():.2e}")
h, w = 2160, 3840
clip = synthetic_clip(h, w, 2, seed=24)
g0, g1 = (dev(F.gray_from_bgr(f)) for f in clip)
for step in (1, 4):
    ref = None
    for v in (2, 3, 0):
        lk = make(h, w, 15, 2, step, v)
        out = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
        for _ in range(2): lk(g1, g0, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n): lk(g1, g0, out)
        e1.record(); torch.cuda.synchronize()
        o = out.cpu().numpy()
        if ref is None: ref = o
        P(f"4K step {step} variant {v}: {e0.elapsed_time(e1)/n:.2f} ms per pair; vs variant 2 max abs {np.abs(o-ref).max():.3e}")
