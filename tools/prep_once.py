#!/usr/bin/env python
"""A few 4K Farneback prepare() calls (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transflow_b200 import ops, _lib
from transflow_b200.synthetic import synthetic_clip
from oracle import flow_cv as F
h, w = 2160, 3840
a = torch.from_numpy(F.gray_from_bgr(synthetic_clip(h, w, 1, seed=1)[0])).cuda()
fb = ops.Farneback(h, w)
_lib.load().tf_farneback_tune(1, int(os.environ.get("TWO_PASS", "0")))
for _ in range(3): fb.prepare(0, a)
torch.cuda.synchronize()
print("ok")
