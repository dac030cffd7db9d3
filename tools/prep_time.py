#!/usr/bin/env python
"""4K Farneback prepare() (pyramid + polynomial expansion) timing: fused blur+resize vs the two passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from transflow_b200 import ops, _lib
from transflow_b200.synthetic import synthetic_clip
from oracle import flow_cv as F
h, w = 2160, 3840
a = torch.from_numpy(F.gray_from_bgr(synthetic_clip(h, w, 1, seed=1)[0])).cuda()
fb = ops.Farneback(h, w)
lib = _lib.load()
for two in (1, 0, 1, 0):
    lib.tf_farneback_tune(1, two)
    for _ in range(3): fb.prepare(0, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.tf_timer_enable(1)
    e0.record()
    for _ in range(20): fb.prepare(0, a)
    e1.record(); torch.cuda.synchronize()
    print(f"two_pass={two}: prepare {e0.elapsed_time(e1)/20*1e3:.1f} us; polyexp finest {_lib.timer_read('fb_polyexp_finest')}")
    lib.tf_timer_enable(0)
