#!/usr/bin/env python
"""Round-2 kernels under compute-sanitizer (memcheck / racecheck): the ring kernels (TMA + mbarriers), the default
half-buffer configuration, the column-mapped Lucas-Kanade tracker and the sum fast path, at sizes with ragged tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from transflow_b200 import ops
from transflow_b200.compositor import Compositor
from transflow_b200.compositor.pixmap_source_interface import PixmapSourceInterface, StillQueue
from transflow_b200.config import LayerConfig
from transflow_b200.synthetic import synthetic_clip, cnoise_pixmap
for (h, w) in ((540, 960), (603, 812)):
    clip = synthetic_clip(h, w, 2, seed=1)
    g = [ops.gray_from_bgr(torch.from_numpy(f).cuda()) for f in clip]
    rolled = torch.roll(g[0], (9, 13), (0, 1))
    for variant in (8, 24, 23, 19, 21):
        fb = ops.Farneback(h, w, variant=variant)
        fb(g[0], g[1]); fb(g[0], rolled)
    for step in (1, 4):
        ops.LucasKanade(h, w, 15, 2, step)(g[0], g[1])
    ops.LucasKanade(h, w, 9, 3, 2)(g[0], g[1])
    flow = ops.PostProcess(h, w, False)(ops.Farneback(h, w)(g[0], g[1]))
    for kind in ("sum", "moveref"):
        comp = Compositor.from_args(h, w, [LayerConfig(0, kind, reset_mode="random", reset_random_factor=0.5)])
        comp.set_sources({0: [PixmapSourceInterface(StillQueue(torch.from_numpy(cnoise_pixmap(h, w, 1)).cuda()), np.ones((h, w), bool))]})
        for _ in range(2):
            comp.step(flow)
torch.cuda.synchronize()
print("sanitize round-2 exercise ok")
