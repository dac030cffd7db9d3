#!/usr/bin/env python
"""A few frames of a bench config (default C3 at 4K, device-resident, one pair at a time), for ncu captures."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from transflow_b200 import ops
ns = argparse.Namespace(config=os.environ.get("PROF_CONFIG", "C3"), height=int(os.environ.get("PROF_H", 0)),
                        width=int(os.environ.get("PROF_W", 0)), frames_per_step=0,
                        lk_step=int(os.environ.get("PROF_LK_STEP", 1)))
cfg = B.resolve_config(ns)
H, W = cfg["height"], cfg["width"]
clip, mask, pixmaps = B.build_workload(cfg, 4)
frames = torch.from_numpy(clip).cuda()
video_rgb = frames.flip(-1).contiguous()
est = B.Estimator(cfg, frames, 1)
post = ops.PostProcess(H, W, cfg["direction"] == "forward")
comp = B.make_compositor(cfg, B.write_mask_png(mask, "prof"), pixmaps, video_rgb)
rgb = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
flow = torch.empty((H, W, 2), dtype=torch.float32, device="cuda")
est.begin(0)
for t in range(int(os.environ.get("PROF_STEPS", 3))):
    est.pair(t + 1, flow)
    post(flow)
    comp.step(flow, rgb)
torch.cuda.synchronize()
print("ok", int(rgb.sum()))
