#!/usr/bin/env python
"""4K forward post-process (scatter + gather) and fused compositor step timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from transflow_b200 import ops
h, w = 2160, 3840
rng = np.random.default_rng(0)
yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
flow = np.stack([3 * np.sin(yy / 97) + 2, 2 * np.cos(xx / 131) + 1], -1).astype(np.float32)
src = torch.from_numpy(flow).cuda()
buf = torch.empty_like(src)
post = ops.PostProcess(h, w, True)
def run():
    buf.copy_(src); post(buf)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
t_all = e0.elapsed_time(e1) / 20
e0.record()
for _ in range(20): buf.copy_(src)
e1.record(); torch.cuda.synchronize()
t_copy = e0.elapsed_time(e1) / 20
print(f"forward post-process 4K: {1e3*(t_all - t_copy):.1f} us (copy {1e3*t_copy:.1f} us)")
