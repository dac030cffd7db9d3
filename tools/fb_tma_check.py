import sys, numpy as np, torch
sys.path.insert(0, ".")
from transflow_b200 import ops
from transflow_b200.synthetic import synthetic_clip
from oracle import flow_cv as F
h, w = 540, 960
clip = synthetic_clip(h, w, 2, seed=6)
g0, g1 = (torch.from_numpy(F.gray_from_bgr(f)).cuda() for f in clip)
want = ops.Farneback(h, w, variant=8)(g0, g1)
torch.cuda.synchronize()
print("v8 ok", flush=True)
got = ops.Farneback(h, w, variant=12)(g0, g1)
torch.cuda.synchronize()
print("v12 ran; equal:", bool(torch.equal(want, got)), float((want - got).abs().max()), flush=True)
