#!/usr/bin/env python
"""bench.py -- frames/sec at 4K for flow + accumulate + remap (BASELINE.json metric).

Workload (config C3 of SURVEY.md 8d, the configuration the metric is quoted on): synthetic
3840x2160 clip, Farneback defaults, FORWARD direction, one `moveref` layer with random reset
0.5 and a radial reset mask, seeded colour-noise pixmap.  One step = one frame pair through
BGR->gray, pyramid + polynomial expansion of the new frame, the coarse-to-fine displacement
solve, the forward post-process, and the fused move/reset/remap/composite kernel.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, C ABI)
  python bench.py --impl reference [...]                          reference CPU arm (oracle: cv2 + NumPy)

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes
through the public plugin API (FlowSource + Compositor) with pinned HOST frames in and a HOST
RGB frame out every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H4K, W4K = 2160, 3840
N_DISTINCT = 12            # distinct frames cycled (ping-pong order keeps |flow| small)
BG = "#204060"


def frame_order(t: int, n: int) -> int:
    """0, 1, ..., n-1, n-2, ..., 1, 0, 1, ... (consecutive frames always neighbours in time)."""
    period = 2 * (n - 1)
    k = t % period
    return k if k < n else period - k


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling of SM clocks and throttle reasons DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []
        self.begin = 0          # first line that belongs to the timed region (mark_begin)
        self.end = None         # one past its last line (mark_end); None = everything up to stop()
        self.window = "timed region"

    def mark_begin(self):
        self.begin = len(self.lines)

    def mark_end(self):
        self.end = len(self.lines)

    def samples_in_window(self) -> int:
        return (len(self.lines) if self.end is None else self.end) - self.begin

    def extend(self, keep_busy, want=2, limit_s=4.0):
        """The timed region was shorter than nvidia-smi's period: keep the GPU under the SAME load (untimed steps)
        until `want` samples have arrived, and say so in the line."""
        t0 = time.perf_counter()
        self.end = None
        self.begin = len(self.lines)
        while len(self.lines) - self.begin < want and time.perf_counter() - t0 < limit_s and self.proc is not None:
            keep_busy()
        self.window = "same load, right after the timed region (it was shorter than the sampling period)"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines[self.begin:self.end]:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "window": self.window}


def build_workload(height, width, n_frames, seed=0):
    from transflow_b200.synthetic import cnoise_pixmap, radial_mask, synthetic_clip
    clip = synthetic_clip(height, width, n_frames, seed=seed)       # (T, H, W, 3) BGR uint8
    return clip, radial_mask(height, width), cnoise_pixmap(height, width, seed + 1)


def write_mask_png(mask, tag):
    import PIL.Image
    import tempfile
    path = os.path.join(tempfile.gettempdir(), f"tfb200_mask_{tag}_{os.getpid()}.png")
    PIL.Image.fromarray(np.rint(mask * 255).astype(np.uint8)).save(path)
    return path


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (cv2 Farneback call site + NumPy compositor), oracle port
# ------------------------------------------------------------------------------------------------
def reference_step_fn(clip, mask, pixmap):
    """Returns step(t) running one frame of the reference CPU pipeline on the given frames."""
    import cv2
    from oracle import compositor_np as CN
    from oracle import flow_cv as F
    h, w = clip.shape[1:3]
    layer = CN.LayerOracle(CN.LayerSpec(classname="moveref", reset_mode="random", reset_random_factor=0.5), h, w,
                           intro_masks=[np.ones((h, w), bool)], reset_mask=mask)
    bg = np.empty((h, w, 3), np.uint8)
    bg[:, :] = (0x20, 0x40, 0x60)
    state = {"prev": cv2.cvtColor(clip[0], cv2.COLOR_BGR2GRAY)}

    def step(t):
        gray = cv2.cvtColor(clip[frame_order(t + 1, len(clip))], cv2.COLOR_BGR2GRAY)
        flow = F.farneback(state["prev"], gray)                       # forward: (prev, cur)
        flow = F.post_process(flow, True)
        layer.update(flow, [pixmap])
        out = CN.composite(bg, [layer.render()])
        state["prev"] = gray
        return out
    return step


def cpu_threads():
    import cv2
    return max(int(cv2.getNumThreads()), 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    total = args.steps + args.warmup
    # bounded sample: a full 4K frame costs ~6 s on the CPU path; shrink the band of rows so the
    # whole run stays within ~150 s, and scale the result back to 4K-frame equivalents
    per_step_budget = 150.0 / max(total, 1)
    frac = min(1.0, per_step_budget / 6.0)
    hs = int(min(args.height, max(136, round(args.height * frac / 8) * 8)))
    clip, mask, pixmap = build_workload(hs, args.width, 4, seed=0)
    step = reference_step_fn(clip, mask, pixmap)
    for t in range(args.warmup):
        step(t)
    t0 = time.perf_counter()
    for t in range(args.warmup, total):
        step(t)
    dt = time.perf_counter() - t0
    scale = hs / args.height
    fps = args.steps / dt * scale
    sample = (f"{args.steps} frames of a {args.width}x{hs} band ({scale:.3f} of a {args.width}x{args.height} frame), "
              f"value scaled by that fraction; cv2.calcOpticalFlowFarneback + NumPy post-process + NumPy moveref")
    line = {
        "impl": "reference", "metric": "frames/sec at 4K (flow+accumulate+remap)", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps / scale,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cpu_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args):
    return {"workload": f"C3: synthetic {args.width}x{args.height} clip, Farneback defaults (pyr 0.5, 3 levels, win 15, "
                        "3 iters, poly 5/1.2), forward direction, moveref layer with random reset 0.5 + radial reset "
                        "mask, cnoise pixmap",
            "frames_cycled": N_DISTINCT,
            "cache": "working set per frame (R pyramids 2x220 MB, data 2x133 MB, flow 66 MB at 4K) exceeds the 126 MB L2; "
                     "no explicit flush",
            "reset_rng": "device Philox (throughput mode)",
            "farneback_variant": os.environ.get("TFB200_FB_VARIANT", "default"),
            "pairs_in_flight": max(1, min(2, int(os.environ.get("TFB200_FB_LANES", "2"))))}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def cpu_baseline_sample(args):
    """Oracle timed on this box's host cores on a bounded sample (about 10-30 s of CPU work)."""
    hs = min(args.height, 1080)
    clip, mask, pixmap = build_workload(hs, args.width, 3, seed=0)
    step = reference_step_fn(clip, mask, pixmap)
    step(0)
    t0 = time.perf_counter()
    n = 2
    for t in range(1, 1 + n):
        step(t)
    dt = time.perf_counter() - t0
    scale = hs / args.height
    return {"value": n / dt * scale, "unit": "frames/s", "cores": cpu_threads(), "kind": "port",
            "sample": f"{n} frames of a {args.width}x{hs} band ({scale:.3f} of a frame) through cv2 Farneback + NumPy "
                      "post-process + NumPy moveref oracle, scaled to full frames; cv2 Farneback is effectively "
                      "single-threaded"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from transflow_b200 import _lib, ops
    from transflow_b200.compositor import Compositor
    from transflow_b200.compositor.pixmap_source_interface import PixmapSourceInterface, StillQueue
    from transflow_b200.config import LayerConfig
    from transflow_b200.flow import FlowSource
    from transflow_b200.flow.sources.cv import ArrayCapture, CvFlowConfig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION; stdout carries one JSON line only
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1:
        from transflow_b200 import distributed as tfd
        return tfd.bench_sharded(args, rank, world, local)

    H, W = args.height, args.width
    clip, mask, pixmap = build_workload(H, W, N_DISTINCT, seed=0)
    mask_png = write_mask_png(mask, "bench")
    variant = int(os.environ.get("TFB200_FB_VARIANT", "-1"))
    layer_cfg = dict(classname="moveref", reset_mode="random", reset_random_factor=0.5, reset_mask=mask_png)

    # ---- device-resident throughput (`value`) -------------------------------------------------
    frames_dev = torch.from_numpy(clip).cuda()
    pix_dev = torch.from_numpy(pixmap).cuda()
    fb = ops.Farneback(H, W) if variant < 0 else ops.Farneback(H, W, variant=variant)
    post = ops.PostProcess(H, W, forward=True)
    comp = Compositor.from_args(H, W, [LayerConfig(0, **layer_cfg)], background_color=BG)
    comp.set_sources({0: [PixmapSourceInterface(StillQueue(pix_dev), np.ones((H, W), bool))]})
    rgb = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    # Streams, like the reference's processes (flow SourceProcess -> queue -> compositor main loop,
    # pipeline.py:326-327): the caller's stream post-processes and composites pair t while the flow streams estimate
    # the next pairs.  With two LANES (default) pairs t+1 and t+2 are in flight on two flow streams (frame t in
    # handle slot t % 3, pair t on lane t % 2; every frame is still prepared once): the second pair fills the SMs
    # the first leaves idle.  A ring of flow buffers decouples the lanes from the compositor; events order them.
    lanes = max(1, min(2, int(os.environ.get("TFB200_FB_LANES", "2"))))
    pipelined = os.environ.get("TFB200_BENCH_STREAMS", "2") != "1"
    if not pipelined:
        lanes = 1
    NB = 2 * lanes
    flows = [torch.empty((H, W, 2), dtype=torch.float32, device="cuda") for _ in range(NB)]
    grays = [torch.empty((H, W), dtype=torch.uint8, device="cuda") for _ in range(lanes)]
    flow_streams = [torch.cuda.Stream() if pipelined else torch.cuda.current_stream() for _ in range(lanes)]
    ev_ready = [torch.cuda.Event() for _ in range(NB)]
    ev_free = [torch.cuda.Event() for _ in range(NB)]
    fb.prepare(0, ops.gray_from_bgr(frames_dev[0], grays[0]))
    torch.cuda.synchronize()
    nslots = 3 if lanes > 1 else 2

    def step(t, serial=False):
        lane = t % lanes
        old, new = t % nslots, (t + 1) % nslots
        k = t % NB
        main = torch.cuda.current_stream()
        with torch.cuda.stream(flow_streams[lane]):
            if pipelined:
                flow_streams[lane].wait_event(ev_free[k])  # the compositor is done with this buffer (NB steps ago)
            ops.gray_from_bgr(frames_dev[frame_order(t + 1, N_DISTINCT)], grays[lane])
            # prepare(new) overlapped with solve(old, new): forward direction
            fb.step(new, grays[lane], old, new, flows[k], lane=lane)
            if pipelined:
                ev_ready[k].record(flow_streams[lane])
        if pipelined:
            main.wait_event(ev_ready[k])
        post(flows[k])
        comp.step(flows[k], rgb)
        if pipelined:
            ev_free[k].record(main)
            if serial:      # kernel-alone pass: nothing of step t+1 may overlap step t
                for fs in flow_streams:
                    fs.wait_event(ev_free[k])

    sampler = ClockSampler(local)
    sampler.start()             # running before the warm-up, so that its first samples are not lost to start-up
    for t in range(args.warmup):
        step(t)
    torch.cuda.synchronize()
    sampler.mark_begin()
    _lib.timer_enable(True)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(args.warmup, args.warmup + args.steps):
        step(t)
    e1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    kernel_ms = {tag: _lib.timer_read(tag) for tag in _lib.KERNEL_TAGS}
    _lib.timer_enable(False)
    if sampler.samples_in_window() < 2:
        t_extra = [args.warmup + args.steps]

        def keep_busy():
            for _ in range(20):
                step(t_extra[0])
                t_extra[0] += 1
            torch.cuda.synchronize()
        sampler.extend(keep_busy)
        t_after = t_extra[0]
    else:
        t_after = args.warmup + args.steps
    clocks = sampler.stop()
    fps = args.steps / (ms / 1000.0)
    # With two lanes a kernel's event-bracketed duration in the timed region includes the kernels of the other
    # pair it shares the SMs with.  The roofline of the kernel ITSELF is taken from a second, serialised pass of
    # the same steps (every step waits for the previous one: the ncu launch list's condition); both are reported.
    kernel_ms_region = kernel_ms
    if lanes > 1:
        n_serial = max(20, min(args.steps, 100))
        torch.cuda.synchronize()
        _lib.timer_enable(True)
        t_next = t_after
        for t in range(t_next, t_next + n_serial):
            step(t, serial=True)
        torch.cuda.synchronize()
        kernel_ms = {tag: _lib.timer_read(tag) for tag in _lib.KERNEL_TAGS}
        _lib.timer_enable(False)

    # ---- roofline of the dominant kernel --------------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    n_px = H * W
    # algorithmic bytes per launch: fused iteration = 56 B/px; unfused kernels: their own traffic
    alg = {"fb_iter_finest": 56.0 * n_px, "fb_um_finest": 68.0 * n_px, "fb_boxv_finest": 60.0 * n_px,
           "fb_boxh_finest": 48.0 * n_px}
    cand = {k: v for k, v in kernel_ms.items() if k in alg and v[1] > 0}
    dom = max(cand, key=lambda k: cand[k][0]) if cand else None
    roofline = None
    if dom:
        tot_ms, n = cand[dom]
        achieved = alg[dom] / (tot_ms / n / 1000.0) / 1e9
        # DRAM traffic of the same kernel from the committed ncu --set full capture
        # (profiles/r01_ncu_full_k_fb_iter_half_c4.txt: 453.52 MB read + 46.27 MB written per 4K launch of the
        # default kernel; the rolling-tile kernel, TFB200_FB_VARIANT=3, moved 428.36 + 58.62 MB)
        default_kernel = os.environ.get("TFB200_FB_VARIANT", "8") == "8"
        traffic = None
        if dom == "fb_iter_finest" and (H, W) == (H4K, W4K):
            traffic = 499.80e6 if default_kernel else (486.99e6 if os.environ.get("TFB200_FB_VARIANT") == "3" else None)
        reg_ms, reg_n = kernel_ms_region[dom]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "avg_launch_ms": tot_ms / n, "launches": n, "share_of_step": reg_ms / ms,
                    "algorithmic_bytes_per_launch": alg[dom],
                    "timing": ("CUDA events around each launch on its stream; serialised pass after the timed region "
                               "(one pair at a time), because in the timed region two pairs share the SMs"
                               if lanes > 1 else "CUDA events around each launch on its stream, timed region"),
                    "in_timed_region": {"avg_launch_ms": reg_ms / max(reg_n, 1), "launches": reg_n,
                                        "achieved": alg[dom] / (reg_ms / max(reg_n, 1) / 1000.0) / 1e9,
                                        "frac": alg[dom] / (reg_ms / max(reg_n, 1) / 1000.0) / 1e9 / peak}}
    frame_bytes = fb.algorithmic_bytes(True) + (24.0 + 50.0) * n_px + 3.0 * n_px + n_px
    pipeline_frac = frame_bytes * fps / 1e9 / peak

    # ---- end to end through the public plugin API with HOST frames --------------------------------
    e2e = run_e2e(args, clip, pixmap, layer_cfg)

    cpu = cpu_baseline_sample(args)
    line = {
        "metric": "frames/sec at 4K (flow+accumulate+remap)", "value": fps, "unit": "frames/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "pipeline_hbm_frac": pipeline_frac, "pipeline_algorithmic_bytes_per_frame": frame_bytes,
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kernel_ms_region.items() if v[1] > 0},
    }
    print(json.dumps(line), flush=True)
    return 0


class CyclingCapture:
    """cv2.VideoCapture-like view of the pinned host clip, cycling in ping-pong order."""

    def __init__(self, frames_pinned, total, fps=25.0):
        self.frames, self.total, self.fps, self.pos = frames_pinned, total, fps, 0

    def read(self):
        if self.pos >= self.total:
            return False, None
        f = self.frames[frame_order(self.pos, len(self.frames))]
        self.pos += 1
        return True, f

    def get(self, prop):
        import cv2
        return {cv2.CAP_PROP_FRAME_WIDTH: self.frames.shape[2], cv2.CAP_PROP_FRAME_HEIGHT: self.frames.shape[1],
                cv2.CAP_PROP_FPS: self.fps, cv2.CAP_PROP_FRAME_COUNT: self.total}.get(prop, 0)

    def set(self, prop, value):
        import cv2
        if prop == cv2.CAP_PROP_POS_MSEC:
            self.pos = int(round(value / 1000.0 * self.fps))
        return True

    def release(self):
        pass


def run_e2e(args, clip, pixmap, layer_cfg):
    """The call a user makes: FlowSource.from_args(...) iterated, Compositor.step(flow), frame read
    back to the host.  Every step copies that step's BGR frame H2D (from pinned memory) and the RGB
    result D2H."""
    import torch
    from transflow_b200.compositor import Compositor
    from transflow_b200.compositor.pixmap_source_interface import PixmapSourceInterface, StillQueue
    from transflow_b200.config import LayerConfig
    from transflow_b200.flow import FlowSource
    from transflow_b200.flow.sources.cv import CvFlowConfig
    H, W = args.height, args.width
    frames_pinned = torch.from_numpy(clip).pin_memory()
    total = args.warmup + args.steps + 1
    cap = CyclingCapture(frames_pinned, total)
    comp = Compositor.from_args(H, W, [LayerConfig(0, **layer_cfg)], background_color=BG)
    comp.set_sources({0: [PixmapSourceInterface(StillQueue(torch.from_numpy(pixmap).cuda()), np.ones((H, W), bool))]})
    out_ring = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    dev_ring = [torch.empty((H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
    down = torch.cuda.Stream()
    done = [None, None]
    checksum = 0
    with FlowSource.from_args(cap, cv_config=CvFlowConfig(), direction="forward") as src:
        src.output = "device"
        t0 = None
        for i, flow in enumerate(src):
            if i == args.warmup:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            k = i & 1
            if done[k] is not None:
                done[k].synchronize()
                checksum += int(out_ring[k][0, 0, 0])          # the host consumes the frame
            frame = comp.step(flow, dev_ring[k])
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(down):
                down.wait_event(ready)
                out_ring[k].copy_(frame, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(down)
            done[k] = ev
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return {"value": args.steps / dt, "unit": "frames/s", "h2d_bytes_per_step": int(H * W * 3),
            "d2h_bytes_per_step": int(H * W * 3), "timing": "wall clock between device synchronisations",
            "api": "FlowSource.from_args(capture).__next__ + Compositor.step + pinned D2H"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--height", type=int, default=H4K)
    ap.add_argument("--width", type=int, default=W4K)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
