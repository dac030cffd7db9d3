"""Experiment: aggregate Farneback pairs/s with 1, 2 and 3 independent handles stepping concurrently on
their own streams (does filling the tails / the latency-bound coarse levels with a second pair pay?).

    python tools/fb_concurrency.py [H W [steps]]
"""
import sys

import torch

from transflow_b200 import ops
from transflow_b200.synthetic import synthetic_clip


def run(n_lanes, frames, H, W, steps):
    fbs = [ops.Farneback(H, W) for _ in range(n_lanes)]
    streams = [torch.cuda.Stream() for _ in range(n_lanes)]
    flows = [torch.empty((H, W, 2), dtype=torch.float32, device="cuda") for _ in range(n_lanes)]
    grays = [ops.gray_from_bgr(f) for f in frames]
    slots = [0] * n_lanes
    for fb in fbs:
        fb.prepare(0, grays[0])
    torch.cuda.synchronize()

    def loop(n):
        for t in range(n):
            for i in range(n_lanes):
                with torch.cuda.stream(streams[i]):
                    cur = slots[i] ^ 1
                    fbs[i].step(cur, grays[(t + i + 1) % len(grays)], slots[i], cur, flows[i])
                    slots[i] = cur

    loop(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    loop(steps)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return n_lanes * steps / (ms / 1e3), ms / steps


def run_lanes(lanes, frames, H, W, steps):
    """ONE handle, tf_farneback_step_lane: frame t in slot t % 3, pair t on lane t % 2 (every frame prepared once)."""
    fb = ops.Farneback(H, W)
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    flows = [torch.empty((H, W, 2), dtype=torch.float32, device="cuda") for _ in range(lanes)]
    grays = [ops.gray_from_bgr(f) for f in frames]
    ns = 3 if lanes > 1 else 2
    fb.prepare(0, grays[0])
    torch.cuda.synchronize()
    t0 = [0]

    def loop(n):
        for t in range(t0[0], t0[0] + n):
            lane = t % lanes
            with torch.cuda.stream(streams[lane]):
                fb.step((t + 1) % ns, grays[(t + 1) % len(grays)], t % ns, (t + 1) % ns, flows[lane], lane=lane)
        t0[0] += n

    loop(6)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    loop(steps)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return steps / (e0.elapsed_time(e1) / 1e3)


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 2160
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 3840
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    clip = synthetic_clip(H, W, 6, seed=0)
    frames = [torch.from_numpy(f).cuda() for f in clip]
    if len(sys.argv) > 4 and sys.argv[4] == "handles":
        for n in (1, 2, 3, 1, 2):
            pps, ms = run(n, frames, H, W, steps)
            print(f"{W}x{H} handles {n}: {pps:8.1f} pairs/s aggregate, {ms:.3f} ms per round", flush=True)
        return
    from transflow_b200 import _lib
    import os
    for rows in ([int(v) for v in os.environ["CONC_ROWS"].split(",")] if "CONC_ROWS" in os.environ
                 else (0, 84, 112, 168, 224, 280, 0)):
        _lib.check(_lib.load().tf_farneback_tune(0, rows))
        _lib.check(_lib.load().tf_farneback_tune(2, int(os.environ.get("CONC_MIN_PX", "0"))))
        line = f"{W}x{H} rows/CTA {rows:3d}:"
        for lanes in (1, 2):
            line += f"  lanes {lanes}: {run_lanes(lanes, frames, H, W, steps):7.1f} pairs/s"
        print(line, flush=True)


if __name__ == "__main__":
    main()
