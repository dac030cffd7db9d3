#!/usr/bin/env python
"""One-shot GPU diagnostics: per-stage errors vs the oracle and coarse kernel timings.
Writes gpurun_out/diag.txt (gpurun brings it back)."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
OUT = open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    OUT.write(s + "\n")
    OUT.flush()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    from oracle import farneback_np as FB, flow_cv as F
    from transflow_b200 import ops
    from transflow_b200.synthetic import synthetic_clip
    P("device", torch.cuda.get_device_name(0))
    h, w = 270, 480
    clip = synthetic_clip(h, w, 2, seed=2)
    g0, g1 = F.gray_from_bgr(clip[0]), F.gray_from_bgr(clip[1])
    trace = {}
    ref_np = FB.farneback(g0, g1, trace=trace)
    ref_cv = F.farneback(g0, g1)
    plan = FB.level_plan(w, h, 0.5, 3)
    for variant in (1, 3, 0):
        try:
            fb = ops.Farneback(h, w, variant=variant, debug=True)
            fb.prepare(0, dev(g0)); fb.prepare(1, dev(g1))
            for li, lvl in enumerate(plan):
                k = lvl["k"]
                img = fb.debug_read(1, li, 0).cpu().numpy()
                R0 = fb.debug_read(0, li, 1).cpu().numpy().transpose(1, 2, 0)
                P(f"v{variant} L{li} k={k} img(slot1) err {np.abs(img - trace[k]['I1']).max():.3g}  R0 err {np.abs(R0 - trace[k]['R0']).max():.3g} (|R| max {np.abs(trace[k]['R0']).max():.3g})")
            flow = fb.solve(0, 1).cpu().numpy()
            for li, lvl in enumerate(plan[:-1]):
                fl = fb.debug_read(0, li, 2).cpu().numpy()
                e = np.linalg.norm(fl - trace[lvl['k']]['flows'][-1], axis=-1)
                P(f"v{variant} L{li} flow err mean {e.mean():.3g} max {e.max():.3g}")
            e = np.linalg.norm(flow - ref_cv, axis=-1)
            P(f"v{variant} FINAL vs cv2: mean {e.mean():.3g} max {e.max():.3g}; vs numpy restatement max {np.linalg.norm(flow - ref_np, axis=-1).max():.3g}")
        except Exception:
            P(f"v{variant} FAILED\n" + traceback.format_exc())
    # timings at 4K
    for (hh, ww) in ((1080, 1920), (2160, 3840)):
        clip = synthetic_clip(hh, ww, 2, seed=1)
        a, b = dev(F.gray_from_bgr(clip[0])), dev(F.gray_from_bgr(clip[1]))
        for variant in (1, 3, 0):
            for fp16 in (False, True):
                try:
                    fb = ops.Farneback(hh, ww, variant=variant, r_fp16=fp16)
                    out = torch.empty((hh, ww, 2), dtype=torch.float32, device="cuda")
                    t_prep = timeit(lambda: fb.prepare(1, b))
                    fb.prepare(0, a)
                    t_solve = timeit(lambda: fb.solve(0, 1, out))
                    P(f"{ww}x{hh} variant {variant} fp16={fp16}: prepare {t_prep:.3f} ms  solve {t_solve:.3f} ms  -> {1000/(t_prep+t_solve):.1f} pairs/s; alg bytes {fb.algorithmic_bytes()/1e9:.3f} GB -> {fb.algorithmic_bytes()/1e6/(t_prep+t_solve):.1f} GB/s")
                    fb.close()
                except Exception:
                    P(f"timing {ww}x{hh} v{variant} fp16={fp16} FAILED\n" + traceback.format_exc())
        try:
            t0 = time.perf_counter(); ref = F.farneback(F.gray_from_bgr(clip[0]), F.gray_from_bgr(clip[1])); t1 = time.perf_counter()
            fb = ops.Farneback(hh, ww)
            e = np.linalg.norm(fb(a, b).cpu().numpy() - ref, axis=-1)
            P(f"{ww}x{hh} cv2 farneback {t1-t0:.2f} s; device vs cv2 mean {e.mean():.3g} max {e.max():.3g}")
        except Exception:
            P("cv2 compare FAILED\n" + traceback.format_exc())
        # compositor timing
        try:
            from transflow_b200.compositor import Compositor
            from transflow_b200.compositor.pixmap_source_interface import PixmapSourceInterface, StillQueue
            from transflow_b200.config import LayerConfig
            from transflow_b200.synthetic import cnoise_pixmap
            pix = dev(cnoise_pixmap(hh, ww, 1))
            comp = Compositor.from_args(hh, ww, [LayerConfig(0, "moveref", reset_mode="random", reset_random_factor=0.5)])
            comp.set_sources({0: [PixmapSourceInterface(StillQueue(pix), np.ones((hh, ww), bool))]})
            flow = ops.PostProcess(hh, ww, False)(fb(a, b))
            rgb = torch.empty((hh, ww, 3), dtype=torch.uint8, device="cuda")
            t = timeit(lambda: comp.step(flow, rgb))
            P(f"{ww}x{hh} compositor moveref+random(device rng) fused step {t:.3f} ms -> {50*hh*ww/1e6/t:.1f} GB/s (50 B/px)")
            ppf = ops.PostProcess(hh, ww, True)
            f2 = flow.clone()
            t = timeit(lambda: ppf(f2))
            P(f"{ww}x{hh} postprocess forward {t:.3f} ms")
        except Exception:
            P("compositor timing FAILED\n" + traceback.format_exc())


if __name__ == "__main__":
    try:
        main()
    except Exception:
        P("DIAG FAILED\n" + traceback.format_exc())
        sys.exit(1)
