#!/usr/bin/env python
"""A few steps of the C3 workload at 4K (device-resident), for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from transflow_b200 import ops
from transflow_b200.compositor import Compositor
from transflow_b200.compositor.pixmap_source_interface import PixmapSourceInterface, StillQueue
from transflow_b200.config import LayerConfig
H, W = int(os.environ.get("PROF_H", 2160)), int(os.environ.get("PROF_W", 3840))
clip, mask, pixmap = B.build_workload(H, W, 4, seed=0)
frames = torch.from_numpy(clip).cuda()
fb = ops.Farneback(H, W)
post = ops.PostProcess(H, W, True)
comp = Compositor.from_args(H, W, [LayerConfig(0, "moveref", reset_mode="random", reset_random_factor=0.5,
                                               reset_mask=B.write_mask_png(mask, "prof"))], background_color=B.BG)
comp.set_sources({0: [PixmapSourceInterface(StillQueue(torch.from_numpy(pixmap).cuda()), np.ones((H, W), bool))]})
rgb = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
gray = torch.empty((H, W), dtype=torch.uint8, device="cuda")
flow = torch.empty((H, W, 2), dtype=torch.float32, device="cuda")
fb.prepare(0, ops.gray_from_bgr(frames[0], gray))
slot = 0
for t in range(int(os.environ.get("PROF_STEPS", 3))):
    cur = slot ^ 1
    ops.gray_from_bgr(frames[B.frame_order(t + 1, 4)], gray)
    fb.step(cur, gray, slot, cur, flow); post(flow); comp.step(flow, rgb)
    slot = cur
torch.cuda.synchronize()
print("ok", int(rgb.sum()))
