#!/usr/bin/env python
"""Timings of the secondary configs (C2 Horn-Schunck + sum @1080p, C4 Lucas-Kanade + static/moveref @4K)."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import flow_cv as F
from transflow_b200 import ops
from transflow_b200.compositor import Compositor
from transflow_b200.compositor.pixmap_source_interface import PixmapSourceInterface, StillQueue
from transflow_b200.config import LayerConfig
from transflow_b200.synthetic import synthetic_clip, cnoise_pixmap
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
OUT = open(os.path.join(ROOT, "gpurun_out", "diag2.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); OUT.write(s + "\n"); OUT.flush()
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
try:
    h, w = 1080, 1920
    clip = synthetic_clip(h, w, 2, seed=1)
    g0, g1 = F.gray_from_bgr(clip[0]), F.gray_from_bgr(clip[1])
    a, b = dev(g0), dev(g1)
    hs = ops.HornSchunck(h, w)
    out = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
    t = timeit(lambda: hs(a, b, None, 1, 3, 0, 1, out=out))
    t_nodelta = timeit(lambda: hs(a, b, None, 1, 3, 0, None, out=out))
    t0 = time.perf_counter(); ref = F.horn_schunck(g0, g1); t1 = time.perf_counter()
    e = np.linalg.norm(out.cpu().numpy() - ref, axis=-1)
    P(f"HS 1080p 3 sweeps: {t:.3f} ms (delta=None: {t_nodelta:.3f} ms) sweeps={hs.last_sweeps}; CPU {t1-t0:.2f} s; err mean {e.mean():.2e} max {e.max():.2e}; 98 B/px -> {98*h*w/1e6/t:.0f} GB/s")
    comp = Compositor.from_args(h, w, [LayerConfig(0, "sum")])
    comp.set_sources({0: [PixmapSourceInterface(StillQueue(dev(cnoise_pixmap(h, w, 1))), np.ones((h, w), bool))]})
    rgb = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    fl = ops.PostProcess(h, w, False)(out.clone())
    t = timeit(lambda: comp.step(fl, rgb))
    P(f"sum layer 1080p fused step: {t:.3f} ms -> {30*h*w/1e6/t:.0f} GB/s (30 B/px)")
except Exception:
    P("HS FAILED\n" + traceback.format_exc())
for (h, w) in ((1080, 1920), (2160, 3840)):
    try:
        clip = synthetic_clip(h, w, 2, seed=1)
        g0, g1 = F.gray_from_bgr(clip[0]), F.gray_from_bgr(clip[1])
        a, b = dev(g0), dev(g1)
        out = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
        for step in (1, 4):
            lk = ops.LucasKanade(h, w, 15, 2, step)
            t = timeit(lambda: lk(a, b, out=out), n=3, warm=1)
            msg = f"LK {w}x{h} step {step}: {t:.2f} ms"
            if h == 1080:
                t0 = time.perf_counter(); ref = F.lucas_kanade(g0, g1, 15, 2, step); t1 = time.perf_counter()
                e = np.linalg.norm(out.cpu().numpy() - ref, axis=-1)
                msg += f"; CPU {t1-t0:.2f} s; err mean {e.mean():.2e} max {e.max():.2e}, mismatching px {(e>1e-3).mean():.2e}"
            P(msg)
    except Exception:
        P("LK FAILED\n" + traceback.format_exc())
try:
    h, w = 2160, 3840
    rgba = np.dstack([cnoise_pixmap(h, w, 2), np.full((h, w), 255, np.uint8)])
    comp = Compositor.from_args(h, w, [LayerConfig(0, "static"), LayerConfig(1, "moveref", moving_pixels_leave_empty_spot=True)])
    vid = dev(cnoise_pixmap(h, w, 3))
    comp.set_sources({0: [PixmapSourceInterface(StillQueue(vid), np.ones((h, w), bool))],
                      1: [PixmapSourceInterface(StillQueue(dev(rgba)), np.ones((h, w), bool))]})
    rgb = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    fl = ops.PostProcess(h, w, False)(out.clone())
    t = timeit(lambda: comp.step(fl, rgb))
    P(f"C4 compositor (static + moveref -e) 4K fused step: {t:.3f} ms")
except Exception:
    P("C4 compositor FAILED\n" + traceback.format_exc())
