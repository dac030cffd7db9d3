#!/usr/bin/env python
"""cProfile of the end-to-end loop (FlowSource + Compositor.step + D2H) of a bench config: where the host time goes."""
import argparse, cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
ns = argparse.Namespace(config=os.environ.get("PROF_CONFIG", "C1"), height=0, width=0, frames_per_step=int(os.environ.get("PROF_FPS", 200)),
                        lk_step=1, steps=3, warmup=1, gpus=1)
cfg = B.resolve_config(ns)
clip, mask, pixmaps = B.build_workload(cfg, B.N_DISTINCT)
mask_png = B.write_mask_png(mask, "prof")
B.run_e2e(ns, cfg, clip, pixmaps, mask_png)          # warm
pr = cProfile.Profile()
pr.enable()
out = B.run_e2e(ns, cfg, clip, pixmaps, mask_png)
pr.disable()
print(out)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
