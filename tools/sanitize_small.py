#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family (for compute-sanitizer memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from transflow_b200 import ops
from transflow_b200.compositor import Compositor
from transflow_b200.compositor.pixmap_source_interface import PixmapSourceInterface, StillQueue
from transflow_b200.config import LayerConfig
from transflow_b200.synthetic import synthetic_clip, cnoise_pixmap
from oracle import flow_cv as F
for (h, w) in ((135, 201), (96, 128), (270, 484)):
    clip = synthetic_clip(h, w, 3, seed=1)
    g = [ops.gray_from_bgr(torch.from_numpy(f).cuda()) for f in clip]
    for variant in (3, 1, 0):
        fb = ops.Farneback(h, w, variant=variant)
        fb.prepare(0, g[0])
        fl = fb.step(1, g[1], 0, 1)
        fl2 = fb.step(0, g[2], 1, 0)
    fbw = ops.Farneback(h, w, winsize=33, poly_n=7, poly_sigma=1.5)
    fbw(g[0], g[1])
    hs = ops.HornSchunck(h, w); hs(g[0], g[1], None, 1, 4, 0, 1); hs(g[0], g[1], fl, 1, 4, 0.9, 0.01)
    for step in (1, 4):
        ops.LucasKanade(h, w, 15, 2, step)(g[0], g[1])
    for fwd in (False, True):
        flow = ops.PostProcess(h, w, fwd)(fl.clone())
        pix3, pix4 = cnoise_pixmap(h, w, 1), np.dstack([cnoise_pixmap(h, w, 2), np.full((h, w), 200, np.uint8)])
        for cfgs, srcs in (([LayerConfig(0, "moveref", reset_mode="random", reset_random_factor=0.5)], {0: [pix3]}),
                           ([LayerConfig(0, "static"), LayerConfig(1, "moveref", moving_pixels_leave_empty_spot=True)], {0: [pix3], 1: [pix4]}),
                           ([LayerConfig(0, "sum", reset_mode="linear")], {0: [pix4]}),
                           ([LayerConfig(0, "introduction", moving_pixels_leave_empty_spot=True)], {0: [pix3]}),
                           ([LayerConfig(0, "moveref", reset_mode="constant", mask_alpha="ones")], {0: [pix3, pix4]})):
            comp = Compositor.from_args(h, w, cfgs)
            comp.set_sources({li: [PixmapSourceInterface(StillQueue(torch.from_numpy(np.ascontiguousarray(p)).cuda()), np.ones((h, w), bool)) for p in ps] for li, ps in srcs.items()})
            for _ in range(2):
                comp.step(flow)
                comp.update(flow); comp.render()
            for layer in comp.layers:
                layer.check_indices()
torch.cuda.synchronize()
print("sanitize exercise ok")
