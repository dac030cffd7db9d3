#!/usr/bin/env python
"""bench.py -- frames/sec for flow + accumulate + remap (BASELINE.json metric), any BASELINE config.

  python bench.py [--config C1..C5] [--gpus N] [--steps K] [--warmup W]     our arm (CUDA, C ABI)
  python bench.py --impl reference [...]                                     the reference's own CPU classes

Default = config C3, the one the metric is quoted on: synthetic 3840x2160 clip, Farneback defaults, FORWARD
direction, one `moveref` layer with random reset 0.5 and a radial reset mask.  The other configs of BASELINE.json
 are selectable with --config:

  C1  854x480   Farneback, backward, moveref                       (README basic transfer; River.mp4 stand-in)
  C2  1920x1080 Horn-Schunck (alpha 1, 3 sweeps, decay 0), sum layer
  C3  3840x2160 Farneback, forward, moveref + random reset 0.5 + reset mask
  C4  3840x2160 pyramidal Lucas-Kanade (win 15, 3 levels, dense), static layer fed by the video + moveref -e
                fed by an RGBA still (README sticky texture); --lk-step 4 = assets/configs/lukas-kanade.json
  C5  7680x4320 Farneback, backward, moveref (frame pairs sharded over the ranks at N > 1)

One STEP is a fixed batch of frames (`config.frames_per_step`), sized so that the driver's `--steps 20` window lasts
about a second; `value` = steps * frames_per_step / device time, frames resident in HBM.  `e2e` runs the same number
of frames through the public plugin API (FlowSource.from_args(...) iterated + Compositor.step) with pinned HOST
frames in and a HOST RGB frame out every frame.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DISTINCT = 12            # distinct frames cycled (ping-pong order keeps |flow| small)
BG = "#204060"
METRIC = "frames/sec at 4K (flow+accumulate+remap)"

CONFIGS = {
    "C1": dict(height=480, width=854, method="farneback", direction="backward", frames_per_step=512,
               layers=[dict(classname="moveref")],
               what="README basic transfer stand-in: synthetic 854x480 clip, Farneback defaults, backward, moveref, "
                    "cnoise pixmap"),
    "C2": dict(height=1080, width=1920, method="horn-schunck", direction="backward", frames_per_step=256,
               layers=[dict(classname="sum")],
               what="synthetic 1080p clip, Horn-Schunck (alpha 1, 3 sweeps, decay 0, delta 1), backward, sum layer, "
                    "cnoise pixmap"),
    "C3": dict(height=2160, width=3840, method="farneback", direction="forward", frames_per_step=64,
               layers=[dict(classname="moveref", reset_mode="random", reset_random_factor=0.5, reset_mask="@radial")],
               what="synthetic 3840x2160 clip, Farneback defaults (pyr 0.5, 3 levels, win 15, 3 iters, poly 5/1.2), "
                    "forward direction, moveref layer with random reset 0.5 + radial reset mask, cnoise pixmap"),
    "C4": dict(height=2160, width=3840, method="lukas-kanade", direction="backward", frames_per_step=8,
               layers=[dict(classname="static"), dict(classname="moveref", moving_pixels_leave_empty_spot=True)],
               what="synthetic 3840x2160 clip, pyramidal Lucas-Kanade (win 15, max level 2), backward, layer 0 static "
                    "fed by the video itself, layer 1 moveref with moving_pixels_leave_empty_spot fed by an RGBA still"),
    "C5": dict(height=4320, width=7680, method="farneback", direction="backward", frames_per_step=16,
               layers=[dict(classname="moveref")],
               what="synthetic 7680x4320 stream, Farneback defaults, backward, moveref, cnoise pixmap"),
}


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


def profiled_traffic(kernel_tag: str, shape):
    """DRAM bytes per launch of a kernel from its committed `ncu --set full` capture (profiles/traffic.json), or
    (None, None): this run does not measure DRAM traffic, it quotes the capture and says which."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(path):
        return None, None
    with open(path) as f:
        table = json.load(f)
    entry = table.get(f"{kernel_tag}@{shape[1]}x{shape[0]}")
    if not entry:
        return None, None
    variant = os.environ.get("TFB200_FB_VARIANT")
    if entry.get("variant") is not None and variant is not None and str(entry["variant"]) != variant:
        return None, None
    return float(entry["dram_bytes"]), entry["source"]


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


# ------------------------------------------------------------------------------------------------
# workload (shared by both arms)
# ------------------------------------------------------------------------------------------------
def resolve_config(args):
    cfg = dict(CONFIGS[args.config])
    cfg["name"] = args.config
    if args.height:
        cfg["height"] = args.height
    if args.width:
        cfg["width"] = args.width
    if args.frames_per_step:
        cfg["frames_per_step"] = args.frames_per_step
    cfg["lk_step"] = args.lk_step
    return cfg


def cv_params(cfg) -> dict:
    """Fields of the reference's CvFlowConfig JSON for this config (flow/sources/cv.py:273-305 defaults)."""
    if cfg["method"] == "lukas-kanade":
        return dict(method="lukas-kanade", lk_window_size=15, lk_max_level=2, lk_step=cfg["lk_step"])
    if cfg["method"] == "horn-schunck":
        return dict(method="horn-schunck", hs_alpha=1, hs_iterations=3, hs_decay=0, hs_delta=1)
    return dict(method="farneback")


def build_workload(cfg, n_frames, height=None, seed=0):
    """-> clip (T, H, W, 3) BGR u8, reset mask f32 (H, W), pixmaps per layer (u8 stills; None = the video itself)."""
    from transflow_b200.synthetic import cnoise_pixmap, radial_mask, synthetic_clip
    h, w = height or cfg["height"], cfg["width"]
    clip = synthetic_clip(h, w, n_frames, seed=seed)
    mask = radial_mask(h, w)
    pixmaps = []
    for li, layer in enumerate(cfg["layers"]):
        if layer["classname"] == "static":
            pixmaps.append(None)
        elif cfg["name"] == "C4":       # RGBA still with a transparent centre (stand-in for assets/Frame.png)
            rgba = np.concatenate([cnoise_pixmap(h, w, seed + 1 + li), np.zeros((h, w, 1), np.uint8)], axis=2)
            rgba[..., 3] = np.where(mask > 0.45, 255, 0)
            pixmaps.append(rgba)
        else:
            pixmaps.append(cnoise_pixmap(h, w, seed + 1 + li))
    return clip, mask, pixmaps


def write_mask_png(mask, tag):
    import PIL.Image
    path = os.path.join(tempfile.gettempdir(), f"tfb200_mask_{tag}_{os.getpid()}.png")
    PIL.Image.fromarray(np.rint(mask * 255).astype(np.uint8)).save(path)
    return path


def layer_kwargs(cfg, mask_png):
    out = []
    for layer in cfg["layers"]:
        kw = dict(layer)
        if kw.get("reset_mask") == "@radial":
            kw["reset_mask"] = mask_png
        out.append(kw)
    return out


def workload_config(cfg, extra=None):
    h, w = cfg["height"], cfg["width"]
    d = {"workload": f"{cfg['name']}: {cfg['what']}" + (f" [size overridden: {w}x{h}]"
                                                        if (h, w) != (CONFIGS[cfg['name']]['height'],
                                                                      CONFIGS[cfg['name']]['width']) else ""),
         "frames_per_step": cfg["frames_per_step"], "frames_cycled": N_DISTINCT,
         "cache": "inputs larger than L2: the per-frame working set (R pyramids, data ping-pong, flows; > 1 GB at 4K) "
                  "exceeds the 126 MB L2 and 12 distinct frames are cycled; no explicit flush",
         "reset_rng": "device Philox (throughput mode)"}
    if cfg["method"] == "farneback":
        d["farneback_variant"] = os.environ.get("TFB200_FB_VARIANT", "default")
        d["pairs_in_flight"] = max(1, min(2, int(os.environ.get("TFB200_FB_LANES", "2"))))
    if cfg["method"] == "lukas-kanade":
        d["lk_step"] = cfg["lk_step"]
    if cfg["direction"] == "forward":
        d["forward_flow"] = ("the scatter pass's claim plane goes straight to the compositor layer (no gather pass, no flow "
                             "tensor) when a single move-reference layer is the only consumer")
    if extra:
        d.update(extra)
    return d


def algorithmic_bytes_per_frame(cfg, fb=None) -> float:
    """SURVEY.md 8(d) / DESIGN.md 4: dependency-respecting minimum HBM bytes of one frame of this config."""
    n = float(cfg["height"] * cfg["width"])
    forward = cfg["direction"] == "forward"
    if cfg["method"] == "farneback":
        flow = fb.algorithmic_bytes(True) + n * 3.0 + n      # + BGR frame read, gray write
    elif cfg["method"] == "horn-schunck":
        flow = 98.0 * n + 4.0 * n
    else:
        flow = 2.0 * 1.3125 * n + 16.0 * 3.0 * n / (cfg["lk_step"] ** 2) + 4.0 * n
    post = 24.0 * n if forward else 16.0 * n                 # backward: clip pass reads + writes the flow
    layers = cfg["layers"]
    if (forward and len(layers) == 1 and layers[0]["classname"] == "moveref"
            and os.environ.get("TFB200_FORWARD_CLAIMS", "1") != "0"):
        # claim-plane path: scatter pass only (flow read 8 + claim write 4); the layer then reads the claim
        # (4) and clears it (4) where the flow-fed layer reads the flow (8): its 22 / 46 bytes below stay as they are
        post = 12.0 * n
    comp = 0.0
    for layer in cfg["layers"]:
        kind = layer["classname"]
        # moveref: flow 8 + record read and write + pixmap 3 + frame 3; the records are 4 B packed (13/13/1/5 bits)
        # up to 8192 x 8192, the reference's int32 x 4 (16 B) beyond
        packed = max(cfg["height"], cfg["width"]) <= 8192
        comp += {"moveref": 22.0 if packed else 46.0, "sum": 30.0, "static": 6.0}[kind] * n
        if layer.get("reset_mask"):
            comp += 4.0 * n
    return flow + post + comp


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's OWN classes (baseline/_ref, installed from /root/reference) on the host cores
# ------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "transflow", "pipeline.py"))


class _FakeQueue:
    """Stands in for the multiprocessing.Queue a pixmap SourceProcess feeds (one copy per get, like unpickling)."""

    def __init__(self, frames, cycle=False):
        self.frames, self.cycle, self.i = frames, cycle, 0

    def get(self, timeout=None):
        i = frame_order(self.i, len(self.frames)) if self.cycle else min(self.i, len(self.frames) - 1)
        self.i += 1
        return np.array(self.frames[i])


def reference_step_fn(cfg, hs, n_frames, tmp):
    """-> (step(), kind): one frame through the reference's CPU path on a band of `hs` rows.

    kind "reference": the unmodified reference classes -- `FlowSource.from_args(<FFV1 clip>, cv_config=<json>,
    direction=...)` iterated (CvFlowSource: decode, resize, BGR2GRAY, cv2 / NumPy flow, post_process) then
    `Compositor.update(flow)` + `Compositor.render()` with in-process stand-ins for the pixmap queues, single process,
    as SURVEY.md 8(d) specifies.  kind "port": baseline/_ref is absent -> the oracle restatement of the same calls."""
    import cv2
    band = dict(cfg, height=hs)
    clip, mask, pixmaps = build_workload(band, min(n_frames, N_DISTINCT), height=hs)
    mask_png = write_mask_png(mask, "ref")
    w = cfg["width"]
    if not reference_available():
        return _port_step_fn(cfg, clip, mask, pixmaps), "port"
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_DIR)
    from transflow.compositor.compositor import Compositor
    from transflow.compositor.pixmap_source_interface import PixmapSourceInterface
    from transflow.config import LayerConfig
    from transflow.flow.sources.source import FlowSource
    avi = os.path.join(tmp, "clip.avi")
    vw = cv2.VideoWriter(avi, cv2.VideoWriter_fourcc(*"FFV1"), 25, (w, hs))
    if not vw.isOpened():
        raise RuntimeError("cv2 cannot write FFV1")
    for t in range(n_frames):
        vw.write(clip[frame_order(t, len(clip))])
    vw.release()
    cfg_json = os.path.join(tmp, "cv.json")
    with open(cfg_json, "w") as fp:
        json.dump(cv_params(cfg), fp)
    builder = FlowSource.from_args(avi, cv_config=cfg_json, direction=cfg["direction"])
    source = builder.__enter__()          # (includes the one-off FlowSource.__init__ index arrays: not timed)
    layers = [LayerConfig(i, **kw) for i, kw in enumerate(layer_kwargs(cfg, mask_png))]
    comp = Compositor.from_args(hs, w, layers, background_color=BG)
    everywhere = np.ones((hs, w), bool)
    rgb_video = [np.ascontiguousarray(f[..., ::-1]) for f in clip]
    comp.set_sources({i: [PixmapSourceInterface(_FakeQueue(rgb_video, cycle=True) if p is None else _FakeQueue([p]),
                                                 everywhere)] for i, p in enumerate(pixmaps)})
    it = iter(source)

    def step():
        flow = next(it)
        comp.update(flow)
        return comp.render()
    return step, "reference"


def _port_step_fn(cfg, clip, mask, pixmaps):
    import cv2
    from oracle import compositor_np as CN
    from oracle import flow_cv as F
    h, w = clip.shape[1:3]
    forward = cfg["direction"] == "forward"
    everywhere = np.ones((h, w), bool)
    layers = []
    for kw in cfg["layers"]:
        spec = {k: v for k, v in kw.items() if k != "reset_mask"}
        layers.append(CN.LayerOracle(CN.LayerSpec(**spec), h, w, intro_masks=[everywhere],
                                     reset_mask=mask if kw.get("reset_mask") else None))
    bg = np.empty((h, w, 3), np.uint8)
    bg[:, :] = (0x20, 0x40, 0x60)
    state = {"prev": cv2.cvtColor(clip[0], cv2.COLOR_BGR2GRAY), "t": 0}
    p = cv_params(cfg)

    def step():
        t = state["t"]
        frame = clip[frame_order(t + 1, len(clip))]
        gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        left, right = (state["prev"], gray) if forward else (gray, state["prev"])
        if cfg["method"] == "farneback":
            flow = F.farneback(left, right)
        elif cfg["method"] == "horn-schunck":
            flow = F.horn_schunck(left, right, None, p["hs_alpha"], p["hs_iterations"], p["hs_decay"], p["hs_delta"])
        else:
            flow = F.lukas_kanade(left, right, p["lk_window_size"], p["lk_max_level"], p["lk_step"])
        flow = F.post_process(flow, forward)
        for layer, pm in zip(layers, pixmaps):
            layer.update(flow, [np.ascontiguousarray(frame[..., ::-1]) if pm is None else pm])
        out = CN.composite(bg, [layer.render() for layer in layers])
        state["prev"], state["t"] = gray, t + 1
        return out
    return step


def cpu_threads():
    import cv2
    return max(int(cv2.getNumThreads()), 1)


#: rough seconds of reference CPU work per full frame, only used to size the bounded sample
_REF_SECONDS_PER_FRAME = {"C1": 0.45, "C2": 2.5, "C3": 10.0, "C4": 7.0, "C5": 40.0}


def _band_rows(cfg, seconds_per_step):
    frac = min(1.0, seconds_per_step / _REF_SECONDS_PER_FRAME[cfg["name"]])
    return int(min(cfg["height"], max(136, round(cfg["height"] * frac / 8) * 8)))


def _sample_text(cfg, n, hs, kind):
    w, h = cfg["width"], cfg["height"]
    who = ("the unmodified reference classes from baseline/_ref (CvFlowSource via FlowSource.from_args on an FFV1 clip, "
           "Compositor.update + render)" if kind == "reference" else "the oracle port of the reference's calls")
    return (f"{n} frames of a {w}x{hs} band ({hs / h:.3f} of a {w}x{h} frame) through {who}, single process, value "
            f"scaled by that fraction; cv2 threads = {cpu_threads()} (cv2 Farneback itself runs on ~1 core)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = resolve_config(args)
    total = args.steps + args.warmup
    hs = _band_rows(cfg, 150.0 / max(total, 1))
    with tempfile.TemporaryDirectory() as tmp:
        step, kind = reference_step_fn(cfg, hs, total + 2, tmp)
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    scale = hs / cfg["height"]
    fps = args.steps / dt * scale
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps / scale,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cpu_threads(), "kind": kind,
                         "sample": _sample_text(cfg, args.steps, hs, kind)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_sample(cfg):
    """The reference CPU path timed on this box's host cores on a bounded sample (about 10-30 s of CPU work)."""
    if os.environ.get("TFB200_BENCH_SKIP_CPU") == "1":      # kernel A/B runs only; the contract's line always has it
        return None
    n = 2
    hs = _band_rows(cfg, 6.0)
    with tempfile.TemporaryDirectory() as tmp:
        step, kind = reference_step_fn(cfg, hs, n + 3, tmp)
        step()
        t0 = time.perf_counter()
        for _ in range(n):
            step()
        dt = time.perf_counter() - t0
    scale = hs / cfg["height"]
    return {"value": n / dt * scale, "unit": "frames/s", "cores": cpu_threads(), "kind": kind,
            "sample": _sample_text(cfg, n, hs, kind)}


# ------------------------------------------------------------------------------------------------
# our arm: device-resident workload
# ------------------------------------------------------------------------------------------------
class _CyclingQueue:
    """Pixmap queue of a video source that plays the clip itself: frame t + 1 at step t (cycled like the flow)."""

    def __init__(self, frames, first=1):
        self.frames, self.i = frames, first

    def get(self, timeout=None):
        f = self.frames[frame_order(self.i, len(self.frames))]
        self.i += 1
        return f


def make_compositor(cfg, mask_png, pixmaps, video_frames_rgb, to_device=True):
    """Compositor of the config with its pixmap interfaces; `video_frames_rgb` feeds the layers whose pixmap is the
    video itself (C4's static layer)."""
    import torch
    from transflow_b200.compositor import Compositor
    from transflow_b200.compositor.pixmap_source_interface import PixmapSourceInterface, StillQueue
    from transflow_b200.config import LayerConfig
    h, w = cfg["height"], cfg["width"]
    comp = Compositor.from_args(h, w, [LayerConfig(i, **kw) for i, kw in enumerate(layer_kwargs(cfg, mask_png))],
                                background_color=BG, seed=0)
    everywhere = np.ones((h, w), bool)
    sources = {}
    for i, p in enumerate(pixmaps):
        if p is None:
            q = _CyclingQueue(video_frames_rgb)
        else:
            q = StillQueue(torch.from_numpy(p).cuda() if to_device else p)
        sources[i] = [PixmapSourceInterface(q, everywhere)]
    comp.set_sources(sources)
    return comp


class Estimator:
    """Flow of pair (t, t + 1) of the device-resident clip for one config: gray conversion + the method's engine.
    Farneback keeps two pairs in flight on the handle's lanes; frame t lives in slot t % 3."""

    def __init__(self, cfg, frames_dev, lanes):
        import torch
        from transflow_b200 import ops
        self.cfg, self.frames, self.ops = cfg, frames_dev, ops
        h, w = cfg["height"], cfg["width"]
        self.forward = cfg["direction"] == "forward"
        self.method = cfg["method"]
        self.lanes = lanes if self.method == "farneback" else 1
        self.nslots = 3 if self.lanes > 1 else 2
        variant = int(os.environ.get("TFB200_FB_VARIANT", "-1"))
        if self.method == "farneback":
            self.engine = (ops.Farneback(h, w, lanes=self.lanes) if variant < 0
                           else ops.Farneback(h, w, variant=variant, lanes=self.lanes))
            self.grays = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(self.lanes)]
        else:
            p = cv_params(cfg)
            self.engine = (ops.HornSchunck(h, w) if self.method == "horn-schunck"
                           else ops.LucasKanade(h, w, p["lk_window_size"], p["lk_max_level"], p["lk_step"]))
            self.params = p
            self.grays = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(2)]
        self.n = 0      # pairs submitted since begin()

    def frame(self, idx):
        return self.frames[frame_order(idx, len(self.frames))]

    def begin(self, first_frame, frame_fn=None):
        """(Re)start at `first_frame`: its gray image (+ pyramid / expansion for Farneback) is built here."""
        self.frame_fn = frame_fn or self.frame
        self.n = 0
        g = self.ops.gray_from_bgr(self.frame_fn(first_frame), self.grays[0])
        if self.method == "farneback":
            self.engine.prepare(0, g)

    def pair(self, next_frame, out, lane=0):
        """Flow of (previous frame, `next_frame`) into `out`, on the current stream."""
        n = self.n
        self.n = n + 1
        if self.method == "farneback":
            old, new = n % self.nslots, (n + 1) % self.nslots
            g = self.ops.gray_from_bgr(self.frame_fn(next_frame), self.grays[lane])
            left, right = (old, new) if self.forward else (new, old)
            return self.engine.step(new, g, left, right, out, lane=lane)
        prev, cur = self.grays[n & 1], self.grays[(n + 1) & 1]
        self.ops.gray_from_bgr(self.frame_fn(next_frame), cur)
        left, right = (prev, cur) if self.forward else (cur, prev)
        if self.method == "horn-schunck":
            p = self.params
            return self.engine(left, right, None, p["hs_alpha"], p["hs_iterations"], p["hs_decay"], p["hs_delta"],
                               out=out)
        return self.engine(left, right, out=out)


ROOFLINE_KERNELS = {
    # method -> (timer tag, algorithmic bytes per launch as a function of (n_px, lk_step), note)
    "farneback": ("fb_iter_finest", lambda n, s: 56.0 * n, None),
    "horn-schunck": ("hs_sweep", lambda n, s: 28.0 * n, None),
    "lukas-kanade": ("lk_track_finest", lambda n, s: 6.0 * n + 16.0 * n / (s * s),
                     "the tracker is bound by integer ALU / shared-memory throughput, not HBM (SURVEY.md 8d): the HBM "
                     "fraction is reported because the contract asks for it, not as the kernel's ceiling"),
}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from transflow_b200 import _lib, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cfg = resolve_config(args)
    if world > 1:
        # NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION; stdout carries one JSON line only
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return run_sharded(args, cfg, rank, world, local)

    H, W = cfg["height"], cfg["width"]
    FPS = cfg["frames_per_step"]
    clip, mask, pixmaps = build_workload(cfg, N_DISTINCT)
    mask_png = write_mask_png(mask, "bench")
    forward = cfg["direction"] == "forward"

    # ---- device-resident throughput (`value`) -------------------------------------------------
    frames_dev = torch.from_numpy(clip).cuda()
    video_rgb = frames_dev.flip(-1).contiguous() if any(p is None for p in pixmaps) else None
    lanes = max(1, min(2, int(os.environ.get("TFB200_FB_LANES", "2"))))
    pipelined = os.environ.get("TFB200_BENCH_STREAMS", "2") != "1"
    if not pipelined:
        lanes = 1
    est = Estimator(cfg, frames_dev, lanes)
    lanes = est.lanes
    post = ops.PostProcess(H, W, forward=forward)
    comp = make_compositor(cfg, mask_png, pixmaps, video_rgb)
    rgb = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    # Streams, like the reference's processes (flow SourceProcess -> queue -> compositor main loop,
    # pipeline.py:326-327): the caller's stream post-processes and composites pair t while the flow stream(s) estimate
    # the next pair(s).  Farneback keeps two pairs in flight (handle lanes, one stream each); a ring of flow buffers
    # decouples the estimation from the compositor; events order them.
    NB = 2 * lanes
    flows = [torch.empty((H, W, 2), dtype=torch.float32, device="cuda") for _ in range(NB)]
    flow_streams = [torch.cuda.Stream() if pipelined else torch.cuda.current_stream() for _ in range(lanes)]
    ev_ready = [torch.cuda.Event() for _ in range(NB)]
    ev_free = [torch.cuda.Event() for _ in range(NB)]
    est.begin(0)
    torch.cuda.synchronize()

    def step(t, serial=False):
        lane = t % lanes
        k = t % NB
        main = torch.cuda.current_stream()
        with torch.cuda.stream(flow_streams[lane]):
            if pipelined:
                flow_streams[lane].wait_event(ev_free[k])  # the compositor is done with this buffer (NB frames ago)
            est.pair(t + 1, flows[k], lane=lane)
            if pipelined:
                ev_ready[k].record(flow_streams[lane])
        if pipelined:
            main.wait_event(ev_ready[k])
        if forward:
            comp.step(post.claims(flows[k]), rgb)      # scatter pass -> claim plane -> compositor (no gather pass)
        else:
            post(flows[k])
            comp.step(flows[k], rgb)
        if pipelined:
            ev_free[k].record(main)
            if serial:      # kernel-alone pass: nothing of frame t+1 may overlap frame t
                for fs in flow_streams:
                    fs.wait_event(ev_free[k])

    sampler = ClockSampler(local)
    sampler.start()             # running before the warm-up, so that its first samples are not lost to start-up
    t = 0
    for _ in range(args.warmup * FPS):
        step(t)
        t += 1
    torch.cuda.synchronize()
    sampler.mark_begin()
    _lib.timer_enable(True)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps * FPS):
        step(t)
        t += 1
    e1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    n_frames = args.steps * FPS
    launches = _lib.launch_count() - launches0
    kernel_ms = {tag: _lib.timer_read(tag) for tag in _lib.KERNEL_TAGS}
    _lib.timer_enable(False)
    if sampler.samples_in_window() < 2:
        def keep_busy():
            nonlocal t
            for _ in range(max(4, min(FPS, 32))):
                step(t)
                t += 1
            torch.cuda.synchronize()
        sampler.extend(keep_busy)
    clocks = sampler.stop()
    fps = n_frames / (ms / 1000.0)
    # With two lanes a kernel's event-bracketed duration in the timed region includes the kernels of the other
    # pair it shares the SMs with.  The roofline of the kernel ITSELF is taken from a second, serialised pass of
    # the same frames (every frame waits for the previous one: the ncu launch list's condition); both are reported.
    kernel_ms_region = kernel_ms
    if pipelined:
        n_serial = max(8, min(n_frames, 64))
        torch.cuda.synchronize()
        _lib.timer_enable(True)
        for _ in range(n_serial):
            step(t, serial=True)
            t += 1
        torch.cuda.synchronize()
        kernel_ms = {tag: _lib.timer_read(tag) for tag in _lib.KERNEL_TAGS}
        _lib.timer_enable(False)

    # ---- roofline of the dominant kernel --------------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    n_px = H * W
    tag, alg_fn, note = ROOFLINE_KERNELS[cfg["method"]]
    alg = alg_fn(float(n_px), cfg["lk_step"])
    roofline = None
    if kernel_ms.get(tag, (0, 0))[1] > 0:
        tot_ms, n = kernel_ms[tag]
        reg_ms, reg_n = kernel_ms_region[tag]
        achieved = alg / (tot_ms / n / 1000.0) / 1e9
        traffic, traffic_src = profiled_traffic(tag, (H, W))
        roofline = {"bound": "hbm", "kernel": tag, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src, "avg_launch_ms": tot_ms / n, "launches": n,
                    "share_of_step": reg_ms / ms, "algorithmic_bytes_per_launch": alg,
                    "timing": ("CUDA events around each launch on its stream; serialised pass after the timed region "
                               "(one frame at a time), because in the timed region neighbouring frames share the SMs"
                               if pipelined else "CUDA events around each launch on its stream, timed region"),
                    "in_timed_region": {"avg_launch_ms": reg_ms / max(reg_n, 1), "launches": reg_n,
                                        "achieved": alg / (reg_ms / max(reg_n, 1) / 1000.0) / 1e9,
                                        "frac": alg / (reg_ms / max(reg_n, 1) / 1000.0) / 1e9 / peak}}
        if note:
            roofline["note"] = note
    frame_bytes = algorithmic_bytes_per_frame(cfg, est.engine if cfg["method"] == "farneback" else None)
    pipeline_frac = frame_bytes * fps / 1e9 / peak

    # ---- end to end through the public plugin API with HOST frames --------------------------------
    del est, comp, flows, frames_dev
    torch.cuda.empty_cache()
    e2e = run_e2e(args, cfg, clip, pixmaps, mask_png)

    cpu = cpu_baseline_sample(cfg)
    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(cfg),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "frames_timed": n_frames, "ms_per_frame": ms / n_frames,
        "pipeline_hbm_frac": pipeline_frac, "pipeline_algorithmic_bytes_per_frame": frame_bytes,
        "kernel_ms_per_frame": {k: v[0] / n_frames for k, v in kernel_ms_region.items() if v[1] > 0},
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


E2E_LOOKAHEAD = 4      # frames the source may still run ahead when the clock stops


def run_e2e(args, cfg, clip, pixmaps, mask_png):
    """The call a user makes: FlowSource.from_args(...) iterated, Compositor.step(flow), frame read back to the
    host.  Every frame copies that frame's BGR image H2D (from pinned memory, into the source's preallocated device
    ring) and the RGB result D2H.  The clock runs between two device synchronisations over exactly
    steps * frames_per_step frames in the steady state: the source is in its look-ahead regime at BOTH ends (the
    capture holds E2E_LOOKAHEAD more frames than are timed), so no frame whose flow was computed before the clock
    started is counted without an equal amount of look-ahead work inside the window."""
    import torch
    from transflow_b200.flow import FlowSource
    from transflow_b200.flow.sources.cv import CvFlowConfig
    H, W = cfg["height"], cfg["width"]
    FPS = cfg["frames_per_step"]
    n_warm, n_timed = max(args.warmup * FPS, 8), args.steps * FPS
    frames_pinned = torch.from_numpy(clip).pin_memory()
    video_rgb = None
    if any(p is None for p in pixmaps):
        video_rgb = torch.from_numpy(np.ascontiguousarray(clip[..., ::-1])).pin_memory()
    cap = CyclingCapture(frames_pinned, n_warm + n_timed + 1 + E2E_LOOKAHEAD)
    comp = make_compositor(cfg, mask_png, pixmaps, video_rgb)
    out_ring = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    dev_ring = [torch.empty((H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
    down = torch.cuda.Stream()
    done = [None, None]
    checksum = 0
    p = cv_params(cfg)
    with FlowSource.from_args(cap, cv_config=CvFlowConfig(**p), direction=cfg["direction"]) as src:
        src.output = "claims"       # device flows; forward flows reach Compositor.step as the scatter's claim plane
        t0 = dt = None
        for i, flow in enumerate(src):
            if i == n_warm:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            if i == n_warm + n_timed:
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                break
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
            done[k] = ev          # (dev_ring[k] is rewritten two frames later, after done[k].synchronize() above)
        if dt is None:
            raise RuntimeError("the capture ended before the timed frames were done")
        torch.cuda.synchronize()
    h2d = H * W * 3 * (2 if video_rgb is not None else 1)
    return {"value": n_timed / dt, "unit": "frames/s", "h2d_bytes_per_step": int(h2d * FPS),
            "d2h_bytes_per_step": int(H * W * 3 * FPS), "frames_timed": n_timed,
            "timing": "wall clock between device synchronisations, steady state at both ends",
            "api": "FlowSource.from_args(capture).__next__ + Compositor.step + pinned D2H"}


# ------------------------------------------------------------------------------------------------
# our arm, N > 1: frame pairs sharded over the ranks, accumulate + remap on rank 0
# ------------------------------------------------------------------------------------------------
class HostFrameFeeder:
    """Pinned host clip -> device frames through a side stream with one frame of look-ahead, into a preallocated
    ring of device frames (a slot is rewritten once its consumer's event has passed)."""

    def __init__(self, frames_pinned, order_fn, depth=4):
        import torch
        self.torch = torch
        self.frames, self.order = frames_pinned, order_fn
        self.stream = torch.cuda.Stream()
        self.ring = [torch.empty(tuple(frames_pinned.shape[1:]), dtype=torch.uint8, device="cuda") for _ in range(depth)]
        self.used = [None] * depth
        self.next_slot = 0
        self.cache = {}

    def _start(self, idx):
        if idx in self.cache:
            return
        torch = self.torch
        slot = self.next_slot
        self.next_slot = (slot + 1) % len(self.ring)
        with torch.cuda.stream(self.stream):
            if self.used[slot] is not None:
                self.stream.wait_event(self.used[slot])
            self.ring[slot].copy_(self.frames[self.order(idx)], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.cache[idx] = (slot, ev)

    def get(self, idx):
        """Device frame `idx`, valid for the kernels queued on the current stream until `release` is called."""
        self._start(idx)
        slot, ev = self.cache.pop(idx)
        for stale in list(self.cache):
            if stale != idx + 1:
                del self.cache[stale]
        self._start(idx + 1)
        self.torch.cuda.current_stream().wait_event(ev)
        self._last = slot
        return self.ring[slot]

    def release(self):
        """The consumer of the last frame handed out has been queued on the current stream."""
        ev = self.torch.cuda.Event()
        ev.record()
        self.used[self._last] = ev


class _QueuedChunk(list):
    """Flows of a chunk queued with join=False: ready() orders the current stream after that chunk only."""

    def __init__(self, flows):
        super().__init__(flows)
        self.events = []

    def ready(self):
        import torch
        cur = torch.cuda.current_stream()
        for ev in self.events:
            cur.wait_event(ev)


def run_sharded(args, cfg, rank, world, local):
    import torch
    import torch.distributed as dist
    from transflow_b200 import _lib, ops
    from transflow_b200.distributed import ShardedFlowStream, plan_rank0_pairs, plan_round

    H, W = cfg["height"], cfg["width"]
    K, Q = 8, 4
    clip, mask, pixmaps = build_workload(cfg, N_DISTINCT)
    forward = cfg["direction"] == "forward"
    frames_dev = torch.from_numpy(clip).cuda()
    feeder = HostFrameFeeder(torch.from_numpy(clip).pin_memory(), lambda i: frame_order(i, N_DISTINCT))
    io = {"host": False}

    def frame(idx):
        return feeder.get(idx) if io["host"] else frames_dev[frame_order(idx, N_DISTINCT)]

    lanes_req = max(1, min(2, int(os.environ.get("TFB200_FB_LANES", "2"))))
    est = Estimator(cfg, frames_dev, lanes_req)
    lanes = est.lanes
    # two pairs of a chunk in flight (handle lanes): pair i on lane i % 2 / its own stream; each lane has its own
    # post-process scratch.  Chunks stay ordered on the caller's stream.
    posts = [ops.PostProcess(H, W, forward=forward) for _ in range(lanes)]
    lane_streams = [torch.cuda.Stream() for _ in range(lanes)] if lanes > 1 else None
    est_main = torch.cuda.Stream()
    # hand-off of a finished flow to rank 0: "store" = the post-process gather kernel stores straight into rank 0's
    # ring over NVLink (the kernel then runs at NVLink speed on the lane); "copy" = post-process locally, then a
    # copy-engine transfer on a side stream while the lane already solves the next pair
    handoff = os.environ.get("TFB200_HANDOFF", "copy")
    copy_streams = [torch.cuda.Stream() for _ in range(lanes)]
    peer_views = {}

    def peer_view(address):
        from transflow_b200.peer import raw_tensor
        if address not in peer_views:
            peer_views[address] = raw_tensor(address, (H, W, 2))
        return peer_views[address]
    chunk_state = {"next": None}
    # forward flows as claim planes (the scatter pass's int32 plane, 4 bytes per pixel) instead of flows: decided on rank 0
    # (its compositor must take them) and shared below; rank 0 reads them from the flow slots' first halves
    claims_mode = {"on": False}
    claim_planes = [[torch.zeros((H, W), dtype=torch.int32, device="cuda") for _ in range(2)] for _ in range(lanes)]
    claim_clear = [[None, None] for _ in range(lanes)]
    claim_turn = [0] * lanes
    peer_claim_views = {}

    def peer_claims_view(address):
        from transflow_b200.peer import raw_tensor
        if address not in peer_claim_views:
            peer_claim_views[address] = raw_tensor(address, (H, W), dtype=torch.int32)
        return peer_claim_views[address]

    def as_compositor_input(f):
        """What rank 0's compositor is stepped with: a flow, or (claims mode) the claim plane behind `f`."""
        if not claims_mode["on"]:
            return f
        plane = f if f.dtype == torch.int32 else f.view(torch.int32).reshape(-1)[:H * W].view(H, W)
        return ops.ForwardClaims.from_plane(posts[0], plane)

    def join_lanes():
        main = torch.cuda.current_stream()
        if lane_streams:
            for s in lane_streams:
                main.wait_stream(s)
        for s in copy_streams:
            main.wait_stream(s)

    def estimate_chunk(first_pair, n_pairs, outs=None, gate=None, join=True):
        """n_pairs consecutive pairs: that many solves + post-processes, and one extra prepare when the chunk does
        not continue the previous one (a rank's chunks are consecutive in frame order within a round: the lanes then
        run on without a new prepare and WITHOUT draining -- the lanes are never ordered after the caller's stream
        between chunks).  `outs` (raw device addresses, possibly peer memory) receive the post-processed flows;
        `gate` is queued on every stream that stores into them (the ring's "slot free" wait of a new round);
        join=False leaves the caller's stream unordered after the chunk (the caller joins before it reads)."""
        main = torch.cuda.current_stream()
        if gate is not None:
            gate()
            if lane_streams:
                for s in lane_streams:
                    with torch.cuda.stream(s):
                        gate()
        if chunk_state["next"] != first_pair:
            # the first frame's gray image + expansion are built on the estimator's own stream, ordered after the
            # lanes (they may still read the slot) but NOT after the caller's stream: on rank 0 that stream holds the
            # accumulation of a whole round
            with torch.cuda.stream(est_main if lane_streams else main):
                if lane_streams:
                    for s in lane_streams:
                        est_main.wait_stream(s)
                est.begin(first_pair, frame)
                if io["host"]:
                    feeder.release()
            if lane_streams:
                for s in lane_streams:
                    s.wait_stream(est_main)
        flows = []
        for i in range(n_pairs):
            lane = est.n % lanes
            with torch.cuda.stream(lane_streams[lane] if lane_streams else main):
                flow = est.pair(first_pair + i + 1, None, lane=lane)
                if io["host"]:
                    feeder.release()
                flow.record_stream(main)
                if claims_mode["on"] and outs is None:
                    # rank 0's own pair: scatter pass only, into a plane of its own (consumed -- and zeroed -- by the
                    # compositor up to a round later)
                    plane = torch.zeros((H, W), dtype=torch.int32, device="cuda")
                    plane.record_stream(main)
                    posts[lane].claims(flow, plane=plane)
                    flows.append(plane)
                elif claims_mode["on"]:
                    # producer: scatter into one of the lane's two planes, copy-engine transfer of the 4-byte claims into
                    # the first half of rank 0's flow slot, then the plane is cleared for its next use
                    k = claim_turn[lane] & 1
                    claim_turn[lane] += 1
                    plane = claim_planes[lane][k]
                    if claim_clear[lane][k] is not None:
                        torch.cuda.current_stream().wait_event(claim_clear[lane][k])
                    posts[lane].claims(flow, plane=plane)
                    done = torch.cuda.Event()
                    done.record()
                    cs = copy_streams[lane]
                    with torch.cuda.stream(cs):
                        cs.wait_event(done)
                        peer_claims_view(outs[i]).copy_(plane, non_blocking=True)
                        plane.zero_()
                        cleared = torch.cuda.Event()
                        cleared.record(cs)
                    claim_clear[lane][k] = cleared
                    flows.append(outs[i])
                elif outs is None or handoff == "store":
                    flows.append(posts[lane](flow, None if outs is None else outs[i]))
                else:
                    posts[lane](flow)
                    done = torch.cuda.Event()
                    done.record()
                    cs = copy_streams[lane]
                    with torch.cuda.stream(cs):
                        cs.wait_event(done)
                        peer_view(outs[i]).copy_(flow, non_blocking=True)
                        flow.record_stream(cs)
                    flows.append(outs[i])
        chunk_state["next"] = first_pair + n_pairs
        if join:
            join_lanes()
            return flows
        done = _QueuedChunk(flows)
        for s in (lane_streams or [main]):
            ev = torch.cuda.Event()
            ev.record(s)
            done.events.append(ev)
        return done

    comp = None
    mask_png = write_mask_png(mask, f"shard{rank}")
    video_rgb = frames_dev.flip(-1).contiguous() if any(p is None for p in pixmaps) else None
    if rank == 0:
        comp = make_compositor(cfg, mask_png, pixmaps, video_rgb)
        rgb = [torch.empty((H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
        rgb_host = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        down = torch.cuda.Stream()
        copied = [None, None]
        count = {"n": 0}
    fanout_box = {"f": None, "frames": 0}

    def accumulate(flow):
        k = count["n"] & 1
        count["n"] += 1
        if io["host"] and copied[k] is not None:
            torch.cuda.current_stream().wait_event(copied[k])   # the D2H out of this buffer is done
        comp.step(as_compositor_input(flow), rgb[k])
        fan = fanout_box["f"]
        if io["host"] and fan is not None:
            target = fanout_box["frames"] % world            # frame i leaves through rank i % N's PCIe link
            fanout_box["frames"] += 1
            if target != 0:
                fan.send(rgb[k], target)
                done = torch.cuda.Event()
                done.record()
                copied[k] = done                              # rgb[k] may be rewritten once the peer copy ran
                return
        if io["host"]:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(down):
                down.wait_event(ready)
                rgb_host[k].copy_(rgb[k], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(down)
            copied[k] = ev

    # calibrate F (flow per pair) and A (accumulate per frame) on rank 0, share the plan
    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def fresh_chunk(n):
        chunk_state["next"] = None
        return estimate_chunk(0, n)
    want_claims = torch.zeros(1, dtype=torch.int32, device="cuda")
    if (rank == 0 and forward and handoff == "copy" and os.environ.get("TFB200_TRANSPORT", "p2p") == "p2p"
            and os.environ.get("TFB200_FORWARD_CLAIMS", "1") != "0"
            and len(comp.layers) == 1 and comp.layers[0].takes_claims()):
        want_claims += 1
    dist.broadcast(want_claims, src=0)
    claims_mode["on"] = bool(int(want_claims))
    fresh_chunk(K)
    # F = cost of a pair INSIDE a producer's run of consecutive chunks (no new prepare, lanes not drained): time 4 chunks
    # that continue one another, as the stream's chunks do
    cal = {"next": K}

    def continuing_chunk():
        estimate_chunk(cal["next"], K)
        cal["next"] += K
    f_ms = timed(continuing_chunk, 4) / K
    plan = torch.zeros(2, dtype=torch.float64, device="cuda")
    if rank == 0:
        fl = fresh_chunk(1)[0]
        torch.cuda.current_stream().synchronize()
        copies = [fl.clone() for _ in range(5)]       # (claims mode: the compositor clears the plane it consumes)
        accumulate(copies.pop())
        a_ms = timed(lambda: accumulate(copies.pop()), 4)
        plan[0], plan[1] = f_ms, a_ms
        # calibration frames went through the compositor: start the measured stream from a fresh state
        comp = make_compositor(cfg, mask_png, pixmaps, video_rgb)
        count["n"] = 0
    dist.broadcast(plan, src=0)
    f_ms, a_ms = float(plan[0]), float(plan[1])
    counts = plan_round(world, Q, f_ms, a_ms)
    # rank 0's share counted in pairs and queued beside its accumulation (TFB200_RANK0_PAIRS=chunks: whole chunks first)
    # (rank 0's estimation shares the SMs with its accumulation and the two do not add up perfectly: keep one pair of slack)
    p0 = (max(0, plan_rank0_pairs(world, Q, K, f_ms, a_ms) - 1)
          if os.environ.get("TFB200_RANK0_PAIRS", "pairs") == "pairs" else None)
    transport = os.environ.get("TFB200_TRANSPORT", "p2p")
    frames_per_round = sum(counts) * K if p0 is None else sum(counts[1:]) * K + p0
    fan = None
    if os.environ.get("TFB200_FANOUT", "1") == "1":
        from transflow_b200.peer import PeerFrameFanout
        fan = PeerFrameFanout(rank, world, (H, W, 3))
        fanout_box["f"] = fan
    e2e_first_round = {"j": None}

    def round_hook(j):
        # in the end-to-end pass, every non-zero rank queues the D2H of the frames it will be handed
        if io["host"] and fan is not None and rank != 0:
            if e2e_first_round["j"] is None:
                e2e_first_round["j"] = j
            lo = (j - e2e_first_round["j"]) * frames_per_round
            fan.expect(sum(1 for i in range(lo, lo + frames_per_round) if i % world == rank))
    stream = ShardedFlowStream(rank, world, K, counts, estimate_chunk, accumulate, (H, W, 2), "cuda",
                               transport=transport, round_hook=round_hook, rank0_pairs=p0)
    chunk_state["next"] = None
    # rounds per step so that one step lasts about as long as a single-GPU step
    rounds_per_step = max(1, int(round(cfg["frames_per_step"] * world / frames_per_round)))

    def timed_rounds(first, warm, steps):
        stream.run(first, warm)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stream.run(first + warm, steps)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    n_warm, n_steps = args.warmup * rounds_per_step, args.steps * rounds_per_step
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.timer_enable(True)
    launches0 = _lib.launch_count()
    if rank == 0:
        sampler.mark_begin()
    ms = timed_rounds(0, n_warm, n_steps)
    launches_done = _lib.launch_count() - launches0
    kernel_ms = {tag: _lib.timer_read(tag) for tag in _lib.KERNEL_TAGS}
    _lib.timer_enable(False)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the sharded stream must leave rank 0's compositor in the state a single rank reaches on the same frames
    check = {"frames": (n_warm + n_steps) * frames_per_round}
    if rank == 0:
        def state_sum(c):
            total = 0
            for layer in c.layers:
                d = None if getattr(layer, "KIND", None) == "static" else layer.data   # (a static layer has no map)
                if d is not None:
                    total += int(torch.from_numpy(np.ascontiguousarray(d)).long().sum())
                total += int(torch.from_numpy(np.ascontiguousarray(layer.rgba)).long().sum())
            return total
        check["sharded"] = state_sum(comp)
        replay_frames = min(check["frames"], int(os.environ.get("TFB200_REPLAY_FRAMES", "100000")))
        if replay_frames == check["frames"]:
            solo = make_compositor(cfg, mask_png, pixmaps, video_rgb)
            sharded_comp, comp = comp, solo
            chunk_state["next"] = None
            done = 0
            while done < replay_frames:
                n = min(K, replay_frames - done)
                for f in estimate_chunk(done, n):
                    comp.step(as_compositor_input(f), rgb[0])
                done += n
            torch.cuda.synchronize()
            check["single_rank_replay"] = state_sum(solo)
            check["equal"] = check["single_rank_replay"] == check["sharded"]
            comp = sharded_comp
            chunk_state["next"] = None
            if not check["equal"]:
                raise RuntimeError(f"sharded compositor state differs from the single-rank replay: {check}")
    dist.barrier()

    # end to end: every frame enters from pinned host memory, every RGB frame returns to it
    io["host"] = True
    chunk_state["next"] = None
    ms_e2e = timed_rounds(n_warm + n_steps, n_warm, n_steps)
    io["host"] = False
    launches = torch.tensor([launches_done], dtype=torch.float64, device="cuda")
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    if rank == 0:
        frames = n_steps * stream.frames_per_round
        fps = frames / (float(ms) / 1000.0)
        peak, peak_src = measured_peak_gbs()
        n_px = H * W
        frame_bytes = algorithmic_bytes_per_frame(cfg, est.engine if cfg["method"] == "farneback" else None)
        tag, alg_fn, note = ROOFLINE_KERNELS[cfg["method"]]
        alg = alg_fn(float(n_px), cfg["lk_step"])
        roofline = None
        if kernel_ms.get(tag, (0, 0))[1] > 0:
            tot_ms, n = kernel_ms[tag]
            achieved = alg / (tot_ms / n / 1000.0) / 1e9
            traffic, traffic_src = profiled_traffic(tag, (H, W))
            roofline = {"bound": "hbm", "kernel": tag, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peak_src, "avg_launch_ms": tot_ms / n, "launches": n,
                        "algorithmic_bytes_per_launch": alg,
                        "timing": "rank 0's launches, CUDA events around each launch on its stream inside the timed "
                                  "region (two pairs share the SMs there, and the compositor runs beside them)"}
            if note:
                roofline["note"] = note
        ceiling = 1000.0 / a_ms if a_ms > 0 else None
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, dict(
                sharding=(f"chunks of {K} pairs; per round {counts} chunks per rank" if p0 is None else
                          f"chunks of {K} pairs; per round {counts[1:]} chunks per producer rank and {p0} pairs on rank 0, "
                          "queued one round ahead beside its accumulation") +
                         f" (rank 0 also runs the sequential accumulate+remap); {rounds_per_step} rounds per step; "
                         f"transport {transport}: "
                         + (("the producer's last post-process kernel stores the flow into rank 0's ring over NVLink "
                             "peer memory" if handoff == "store" else
                             ("a copy-engine transfer moves the forward scatter's claim plane (int32, 4 bytes per pixel: the "
                              "compositor reads the claimant directly, no flow is formed) into rank 0's ring over NVLink "
                              "peer memory while the lane solves the next pair" if claims_mode["on"] else
                              "a copy-engine transfer moves the post-processed flow into rank 0's ring over NVLink peer "
                              "memory while the lane solves the next pair")) +
                            ", counters + cuStreamWaitValue32 order it" if transport == "p2p" else
                            "batched NCCL send/recv, receives posted one round ahead"),
                frames_per_step=stream.frames_per_round * rounds_per_step, state_check=check,
                calibrated_ms={"flow_per_pair": f_ms, "accumulate_per_frame": a_ms},
                sequential_tail_ceiling_fps=ceiling)),
            "roofline": roofline,
            "e2e": {"value": frames / (float(ms_e2e) / 1000.0), "unit": "frames/s",
                    "h2d_bytes_per_step": int(stream.frames_per_round * rounds_per_step * n_px * 3 * (K + 1) / K),
                    "d2h_bytes_per_step": int(stream.frames_per_round * rounds_per_step * n_px * 3),
                    "api": "sharded stream: pinned BGR frames H2D on every rank; RGB frame i leaves through rank i % N's "
                           "PCIe link (rank 0 hands it over NVLink)" if fan is not None else
                           "sharded stream: pinned BGR frames H2D on every rank, RGB frames D2H on rank 0"},
            "gpu_launches": int(launches), "clocks": clocks, "frames_timed": frames,
            "pipeline_hbm_frac": frame_bytes * fps / 1e9 / (peak * world),
            "exchange_bytes_per_step": int(sum(counts[1:]) * K * rounds_per_step * n_px * (4 if claims_mode["on"] else 8)),
            "scaling_ceiling": {"producers": world - 1, "flow_ms_per_pair": f_ms, "accumulate_ms_per_frame": a_ms,
                                "fps_if_producers_never_wait": (world - 1) * 1000.0 / f_ms
                                + ((p0 or 0) / (Q * K)) * 1000.0 / f_ms,
                                "sequential_tail_fps": ceiling},
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--height", type=int, default=0, help="override the config's frame height")
    ap.add_argument("--width", type=int, default=0, help="override the config's frame width")
    ap.add_argument("--frames-per-step", type=int, default=0, help="override the config's batch of frames per step")
    ap.add_argument("--lk-step", type=int, default=1, help="lk_step of the Lucas-Kanade config (1 = dense)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries ONE JSON line: whatever a library writes to file descriptor 1 meanwhile (NCCL's version banner at
    # communicator creation, for one) goes to stderr instead, and the line is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
