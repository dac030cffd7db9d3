#!/usr/bin/env python
"""Time the fused compositor kernel of a bench config alone (device-resident flow, CUDA events)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from transflow_b200 import ops
ns = argparse.Namespace(config=os.environ.get("PROF_CONFIG", "C3"), height=0, width=0, frames_per_step=0, lk_step=1)
cfg = B.resolve_config(ns)
H, W = cfg["height"], cfg["width"]
clip, mask, pixmaps = B.build_workload(cfg, 4)
frames = torch.from_numpy(clip).cuda()
est = B.Estimator(cfg, frames, 1)
post = ops.PostProcess(H, W, cfg["direction"] == "forward")
comp = B.make_compositor(cfg, B.write_mask_png(mask, "ct"), pixmaps, frames.flip(-1).contiguous())
rgb = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
est.begin(0)
flows = []
for t in range(3):
    f = torch.empty((H, W, 2), dtype=torch.float32, device="cuda")
    est.pair(t + 1, f); post(f); flows.append(f)
for i in range(20): comp.step(flows[i % 3], rgb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 200
e0.record()
for i in range(n): comp.step(flows[i % 3], rgb)
e1.record(); torch.cuda.synchronize()
print(f"{cfg['name']} compositor step: {1e3 * e0.elapsed_time(e1) / n:.1f} us per frame; checksum {int(rgb.long().sum())}")
