#!/usr/bin/env python
"""Generate golden vectors by running the REAL reference (``/root/reference``) in the build
container.  ``/root/reference`` does not exist on the GPU box, so the outputs are committed
as small ``.npz`` fixtures next to this script:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Compositor cases drive ``transflow.compositor.layers.*`` directly with in-process fake
pixmap queues; flow cases drive the reference's own ``CvFlowSource`` through
its ``Builder`` on a lossless FFV1 ``.avi`` of the seeded synthetic clip, for all three
methods and both directions (this includes ``FlowSource.post_process``).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(1, REPO)

from transflow.compositor.compositor import Compositor  # noqa: E402
from transflow.compositor.layers.layer import Layer  # noqa: E402
from transflow.compositor.pixmap_source_interface import PixmapSourceInterface  # noqa: E402
from transflow.config import LayerConfig  # noqa: E402

H, W, FRAMES = 20, 28, 6


class FakeQueue:
    """Stands in for the multiprocessing.Queue a pixmap SourceProcess feeds."""

    def __init__(self, frames):
        self.frames = list(frames)
        self.i = 0

    def get(self, timeout=None):
        f = self.frames[min(self.i, len(self.frames) - 1)]
        self.i += 1
        return f.copy()


def clipped_flows(rng, n, h, w, mag, fractional=True):
    """Random flows passed through the final clip of post_process (source.py:361-362)."""
    f = rng.uniform(-mag, mag, size=(n, h, w, 2)).astype(np.float32)
    if not fractional:
        f = np.rint(f).astype(np.float32)
    # make a good share of pixels static and some exactly at .5 to exercise half-even rounding
    f[rng.random((n, h, w)) < 0.3] = 0
    half = rng.random((n, h, w)) < 0.1
    f[half] = np.floor(f[half]) + 0.5
    xs = np.arange(w, dtype=np.float32)[None, None, :]
    ys = np.arange(h, dtype=np.float32)[None, :, None]
    f[..., 0] = np.clip(f[..., 0], -xs, w - 1 - xs)
    f[..., 1] = np.clip(f[..., 1], -ys, h - 1 - ys)
    return f


def save_mask_png(arr_u8, path):
    import PIL.Image
    PIL.Image.fromarray(arr_u8).save(path)


def compositor_cases(tmp):
    rng = np.random.default_rng(7)
    grad = np.clip(np.rint(np.linspace(0, 255, W)[None, :] * np.ones((H, 1))), 0, 255).astype(np.uint8)
    blob = (rng.random((H, W)) < 0.6).astype(np.uint8) * 255
    blob2 = (rng.random((H, W)) < 0.5).astype(np.uint8) * 255
    p_grad, p_blob, p_blob2 = (os.path.join(tmp, n) for n in ("grad.png", "blob.png", "blob2.png"))
    save_mask_png(grad, p_grad)
    save_mask_png(blob, p_blob)
    save_mask_png(blob2, p_blob2)
    rgb = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(FRAMES)]
    rgb_b = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(FRAMES)]
    rgba = []
    for _ in range(FRAMES):
        a = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
        a[..., 3] = np.where(rng.random((H, W)) < 0.4, 0, a[..., 3])
        rgba.append(a)
    all_on = np.ones((H, W), bool)
    left = np.zeros((H, W), bool)
    left[:, : W // 2] = True
    ragged = rng.random((H, W)) < 0.7

    # name -> (list of LayerConfig kwargs, {layer: [(pixmap frames, intro mask)]}, flow magnitude)
    cases = {
        "moveref_default": ([dict(classname="moveref")], {0: [(rgb, all_on)]}, 3.0),
        "moveref_random_mask": ([dict(classname="moveref", reset_mode="random", reset_random_factor=0.5,
                                      reset_mask=p_grad)], {0: [(rgb[:1], all_on)]}, 3.0),
        "moveref_random_full": ([dict(classname="moveref", reset_mode="random", reset_random_factor=1)],
                                {0: [(rgb[:1], all_on)]}, 2.0),
        "moveref_constant": ([dict(classname="moveref", reset_mode="constant", reset_constant_step=1.5,
                                   reset_mask=p_grad)], {0: [(rgb[:1], all_on)]}, 4.0),
        "moveref_linear": ([dict(classname="moveref", reset_mode="linear", reset_linear_factor=0.3,
                                 reset_mask=p_grad)], {0: [(rgb[:1], all_on)]}, 4.0),
        "moveref_leave_empty_rgba": ([dict(classname="moveref", moving_pixels_leave_empty_spot=True)],
                                     {0: [(rgba[:1], all_on)]}, 3.0),
        "moveref_flags": ([dict(classname="moveref", moving_pixels_leave_empty_spot=True,
                                pixels_can_move_to_filled_spot=False, mask_src=p_blob, mask_dst=p_blob2,
                                mask_alpha=p_grad)], {0: [(rgb, all_on)]}, 3.0),
        "moveref_transparent": ([dict(classname="moveref", moving_pixels_leave_empty_spot=True,
                                      transparent_pixels_can_move=True, pixels_can_move_to_empty_spot=False)],
                                {0: [(rgb[:1], all_on)]}, 3.0),
        "moveref_two_sources_reset_source": (
            [dict(classname="moveref", reset_mode="random", reset_random_factor=0.3, reset_source=True)],
            {0: [(rgb, left), (rgba, ~left | ragged)]}, 3.0),
        "moveref_two_rgb_sources": ([dict(classname="moveref")], {0: [(rgb, left), (rgb_b, ~left)]}, 3.0),
        "sum_default": ([dict(classname="sum")], {0: [(rgb, all_on)]}, 2.5),
        "sum_linear": ([dict(classname="sum", reset_mode="linear", reset_linear_factor=0.2)],
                       {0: [(rgba, all_on)]}, 2.5),
        "static_then_moveref": ([dict(classname="static"),
                                 dict(classname="moveref", moving_pixels_leave_empty_spot=True)],
                                {0: [(rgb, all_on)], 1: [(rgba[:1], all_on)]}, 3.0),
        "static_ragged": ([dict(classname="static")], {0: [(rgb, ragged), (rgba, left)]}, 1.0),
        "introduction_default": ([dict(classname="introduction")], {0: [(rgb, ragged)]}, 3.0),
        "introduction_rgba_once": ([dict(classname="introduction", introduce_once=True,
                                         moving_pixels_leave_empty_spot=True)],
                                   {0: [(rgba, all_on)]}, 3.0),
        "introduction_flags": ([dict(classname="introduction", introduce_pixels_on_filled_spots=False,
                                     introduce_moving_pixels=False, mask_alpha=p_grad)],
                               {0: [(rgb, ragged), (rgba, left)]}, 3.0),
        "introduction_all_filled": ([dict(classname="introduction", introduce_on_all_filled_spots=True,
                                          introduce_unmoving_pixels=False,
                                          introduce_pixels_on_empty_spots=False)],
                                    {0: [(rgb, all_on)]}, 3.0),
    }
    out = {}
    for ci, (name, (layer_kw, sources, mag)) in enumerate(cases.items()):
        flows = clipped_flows(np.random.default_rng(100 + ci), FRAMES, H, W, mag)
        cfgs = [LayerConfig(i, **kw) for i, kw in enumerate(layer_kw)]
        comp = Compositor.from_args(H, W, cfgs, background_color="#204060")
        comp.set_sources({li: [PixmapSourceInterface(FakeQueue(fr), m) for fr, m in lst]
                          for li, lst in sources.items()})
        np.random.seed(1234 + ci)
        datas, renders = [], []
        for t in range(FRAMES):
            comp.update(flows[t].copy())
            frame = comp.render()
            renders.append(frame)
            for li, layer in enumerate(comp.layers):
                if getattr(layer, "data", None) is not None:
                    datas.append((li, t, np.array(layer.data)))
        # the random fields the reference consumed (one per frame per random-reset layer, in order)
        np.random.seed(1234 + ci)
        n_random_layers = sum(1 for kw in layer_kw if kw.get("reset_mode") == "random")
        randoms = np.stack([np.random.random(size=(H, W)) for _ in range(FRAMES * n_random_layers)]) \
            if n_random_layers else np.zeros((0, H, W))
        out[f"{name}/flows"] = flows
        out[f"{name}/randoms"] = randoms
        out[f"{name}/render"] = np.stack(renders)
        for li, layer in enumerate(comp.layers):
            if getattr(layer, "data", None) is not None:
                out[f"{name}/data{li}"] = np.stack([d for (l2, _, d) in datas if l2 == li])
            out[f"{name}/mask_alpha{li}"] = layer.mask_alpha
            if hasattr(layer, "mask_src"):
                out[f"{name}/mask_src{li}"] = layer.mask_src
                out[f"{name}/mask_dst{li}"] = layer.mask_dst
            if hasattr(layer, "reset_mask"):
                out[f"{name}/reset_mask{li}"] = layer.reset_mask
        for li, lst in sources.items():
            for si, (fr, m) in enumerate(lst):
                out[f"{name}/pixmap{li}_{si}"] = np.stack(fr)
                out[f"{name}/intro{li}_{si}"] = m
        out[f"{name}/layers"] = np.array(repr(layer_kw))
    np.savez_compressed(os.path.join(HERE, "compositor_golden.npz"), **out)
    print("compositor cases:", len(cases))


def flow_cases(tmp):
    """Reference CvFlowSource on a synthetic FFV1 clip: raw method output + post-processed flow."""
    import cv2
    import json
    from transflow.flow.sources.source import FlowSource
    from transflow_b200.synthetic import synthetic_clip

    fh, fw, n = 96, 128, 4
    clip = synthetic_clip(fh, fw, n, seed=3)
    avi = os.path.join(tmp, "clip.avi")
    vw = cv2.VideoWriter(avi, cv2.VideoWriter_fourcc(*"FFV1"), 25, (fw, fh))
    for f in clip:
        vw.write(f)
    vw.release()
    cap = cv2.VideoCapture(avi)
    for f in clip:
        ok, g = cap.read()
        assert ok and np.array_equal(g, f), "FFV1 round trip is not lossless"
    cap.release()

    configs = {
        "farneback": dict(method="farneback"),
        "farneback_blurred": dict(method="farneback", fb_winsize=19, fb_poly_n=7, fb_poly_sigma=1.5),
        "horn_schunck": dict(method="horn-schunck", hs_alpha=1, hs_iterations=3, hs_decay=0, hs_delta=1),
        "horn_schunck_decay": dict(method="horn-schunck", hs_alpha=10.0, hs_iterations=2, hs_decay=0.9,
                                   hs_delta=1.0),
        "lukas_kanade": dict(method="lukas-kanade", lk_window_size=15, lk_max_level=2, lk_step=1),
        "lukas_kanade_step4": dict(method="lukas-kanade", lk_window_size=15, lk_max_level=2, lk_step=4),
    }
    out = {"clip": clip}
    for name, cfg in configs.items():
        path = os.path.join(tmp, name + ".json")
        with open(path, "w") as fp:
            json.dump(cfg, fp)
        for direction in ("forward", "backward"):
            with FlowSource.from_args(avi, cv_config=path, direction=direction) as src:
                flows = np.stack([np.array(f) for f in src])
            assert flows.shape == (n - 1, fh, fw, 2) and flows.dtype == np.float32
            out[f"{name}/{direction}"] = flows
        out[f"{name}/config"] = np.array(json.dumps(cfg))
    np.savez_compressed(os.path.join(HERE, "flow_golden.npz"), **out)
    print("flow cases:", len(configs))


def kat_cases():
    """The three known-answer tests of the reference (tests/test_compositor.py:29-54), re-run
    through the reference so the expected values are recorded, not retyped."""
    from transflow.compositor.layers.move_reference import MoveReferenceLayer
    flow = np.array([[[0, 1], [0, 1], [0, 0]], [[0, 0], [0, 0], [0, 0]]]).astype(np.float32)
    out = {"flow": flow}
    for name, kw in (("moveref", {}),
                     ("moveref_reset", dict(reset_mode="random", reset_random_factor=1)),
                     ("moveref_reset_mask", dict(reset_mode="random", reset_random_factor=1,
                                                 reset_mask="border-left:1"))):
        layer = MoveReferenceLayer(LayerConfig(0, **kw), 2, 3, [])
        np.random.seed(5)
        layer.update(flow.copy())
        out[name] = np.array(layer.data)
    np.savez_compressed(os.path.join(HERE, "kat_golden.npz"), **out)


def postprocess_cases(tmp):
    """Reference ``FlowSource.post_process`` with flow filters, a mask and a convolution kernel
    (``flow/filters.py``, ``flow/sources/source.py:337-363``), driven through ``FlowSource.from_args`` on the
    same FFV1 clip as ``flow_cases`` (the raw flow is cv2 Farneback with the default parameters)."""
    import cv2
    import json
    from transflow.flow.sources.source import FlowSource
    from transflow_b200.synthetic import synthetic_clip

    fh, fw, n = 96, 128, 4
    clip = synthetic_clip(fh, fw, n, seed=3)
    avi = os.path.join(tmp, "clip_pp.avi")
    vw = cv2.VideoWriter(avi, cv2.VideoWriter_fourcc(*"FFV1"), 25, (fw, fh))
    for f in clip:
        vw.write(f)
    vw.release()
    cfg_path = os.path.join(tmp, "fb.json")
    with open(cfg_path, "w") as fp:
        json.dump(dict(method="farneback"), fp)
    rng = np.random.default_rng(11)
    grad = np.clip(np.add.outer(np.linspace(0, 200, fh), np.linspace(0, 55, fw)), 0, 255).astype(np.uint8)
    p_mask = os.path.join(tmp, "pp_mask.png")
    save_mask_png(grad, p_mask)
    kernels = {"box3": np.full((3, 3), 1 / 9), "rand45": rng.normal(size=(4, 5)), "row7": rng.normal(size=(1, 7)) * 0.5}
    kpaths = {}
    for kname, k in kernels.items():
        kpaths[kname] = os.path.join(tmp, kname + ".npy")
        np.save(kpaths[kname], k)
    cases = {
        "scale": dict(flow_filters="scale=2+t", direction="backward"),
        "threshold_fw": dict(flow_filters="threshold=0.5", direction="forward"),
        "clip": dict(flow_filters="clip=1.5*(1+t)", direction="backward"),
        "chain_mask_fw": dict(flow_filters="scale=0.5;clip=1;threshold=0.2", mask_path=p_mask, direction="forward"),
        "strong": dict(flow_filters="scale=numpy.float64(1.5)+t;clip=numpy.sqrt(2.0);threshold=numpy.float64(0.3)",
                       direction="backward"),
        "polar": dict(flow_filters="polar=r*2:a+t", direction="backward"),
        "polar_mid": dict(flow_filters="scale=3;polar=r+1:-a;clip=2", direction="forward"),
        "kernel_box3": dict(kernel_path=kpaths["box3"], direction="backward"),
        "kernel_box3_fw": dict(kernel_path=kpaths["box3"], direction="forward"),
        "kernel_rand45_mask_fw": dict(kernel_path=kpaths["rand45"], mask_path=p_mask, flow_filters="scale=4",
                                      direction="forward"),
        "kernel_row7": dict(kernel_path=kpaths["row7"], flow_filters="clip=2", direction="backward"),
    }
    from transflow.utils import load_float_mask
    out = {"clip": clip, "mask": load_float_mask(p_mask), "framerate": np.array(25.0)}
    for kname, k in kernels.items():
        out["kernel/" + kname] = k
    for name, kw in cases.items():
        with FlowSource.from_args(avi, cv_config=cfg_path, **kw) as src:
            flows = np.stack([np.array(f) for f in src])
        assert flows.shape == (n - 1, fh, fw, 2)
        out[f"{name}/flows"] = flows          # float32, float64 after a convolution kernel
        out[f"{name}/args"] = np.array(json.dumps({k: (os.path.basename(v) if k.endswith("_path") else v)
                                                   for k, v in kw.items()}))
    np.savez_compressed(os.path.join(HERE, "postprocess_golden.npz"), **out)
    print("post-process cases:", len(cases))


def merge_cases(tmp):
    """The reference's flow merging functions (``pipeline.py:149-158``), ``utils.upscale_array`` and a ``.flow.zip``
    written by the reference's own ``NumpyOutput`` (``output/numpy.py``) -- plus a check, here where the reference
    is importable, that its ``ArchiveFlowSource`` reads an archive written by OUR ``NumpyOutput``."""
    import shutil
    from transflow.pipeline import Pipeline
    from transflow.utils import upscale_array
    from transflow.output.numpy import NumpyOutput as RefNumpyOutput
    from transflow.flow.sources.source import FlowSource
    rng = np.random.default_rng(21)
    h, w = 10, 14
    flows = [rng.uniform(-1, 1, (h, w, 2)).astype(np.float32) * s for s in (3, 0.4, 1)]
    flows[1][rng.random((h, w)) < 0.3] = 0
    out = {f"flow{i}": f for i, f in enumerate(flows)}
    for mode, fn in Pipeline.FLOW_MERGING_FUNCTIONS.items():
        for n in ((2,) if mode == "absmax" else (1, 2, 3)):
            out[f"{mode}/{n}"] = np.array(fn([f.copy() for f in flows[:n]]))
    out["upscale/3x2"] = upscale_array(flows[0].copy(), 3, 2)
    np.savez_compressed(os.path.join(HERE, "merge_golden.npz"), **out)
    meta = {"path": "clip.avi", "width": w, "height": h, "framerate": 25.0, "direction": 1, "seek_time": 0}
    ref_zip = os.path.join(tmp, "ref.flow.zip")
    o = RefNumpyOutput(ref_zip, replace=True)
    o.write_meta(meta)
    for f in flows:
        o.write_array(f)
    o.close()
    shutil.copy(ref_zip, os.path.join(HERE, "ref_archive.flow.zip"))
    from transflow_b200.output import NumpyOutput
    ours = os.path.join(tmp, "ours.flow.zip")
    o = NumpyOutput(ours, replace=True)
    o.write_meta(meta)
    for f in flows:
        o.write_array(f)
    o.close()
    with FlowSource.from_args(ours) as src:      # the reference reads our archive
        assert (src.width, src.height, src.framerate, src.direction.value) == (w, h, 25.0, 1)
        # the reference's archive Builder.build never resolves the frame range (archive.py:25-36 does not chain to
        # FlowSource.Builder.build), so its iterator has no end: read the three frames explicitly
        back = [np.array(next(src)) for _ in range(3)]
    from oracle import flow_cv as F
    for got, f in zip(back, flows):
        assert np.array_equal(got, F.post_process(f.copy(), False))     # backward: the final clip only
    print("merge cases:", len(out), "- reference read", len(back), "flows of our archive")


RENDER_CASES = [
    ("2d/default", dict(kind="2d", scale=1, colors=None)),
    ("2d/scaled", dict(kind="2d", scale=0.37, colors=("#ff8000", "#0080ff", "rgb(12, 200, 77)", "#101010"))),
    ("2d/strong", dict(kind="2d", scale=3, colors=None)),
    ("1d/default", dict(kind="1d", scale=1, colors=None, binary=False)),
    ("1d/scaled", dict(kind="1d", scale=0.21, colors=("#203040", "#f0e0d1"), binary=False)),
    ("1d/binary", dict(kind="1d", scale=0.5, colors=("#ff0000", "#00ffff"), binary=True)),
]


def render_cases():
    """The reference's flow visualisers ``render2d`` / ``render1d`` (``output/render.py``), the latter on the
    magnitude exactly as ``Pipeline._update_output`` computes it (``pipeline.py:512-516``)."""
    from transflow.output.render import render1d, render2d
    rng = np.random.default_rng(23)
    h, w = 37, 53
    flow = (rng.standard_normal((h, w, 2)) * 2.5).astype(np.float32)
    flow[0, :8] = np.array([[0, 0], [1, -1], [0.5, 0.5], [-0.5, 1.5], [2.5, -2.5], [1e-3, 0], [100, -100], [3, 4]],
                           dtype=np.float32)
    out = {"flow": flow}
    for name, c in RENDER_CASES:
        if c["kind"] == "2d":
            out[name] = render2d(flow, c["scale"], c["colors"])
        else:
            mag = np.sqrt(np.sum(np.power(flow, 2), axis=2))
            out[name] = render1d(mag, c["scale"], c["colors"], c["binary"])
    np.savez_compressed(os.path.join(HERE, "render_golden.npz"), **out)
    print("render cases:", len(RENDER_CASES))


def lock_cases(tmp):
    """Reference flow sources whose ``prev_flow`` is read again -- a locked flow (stay / skip mode) and Horn-Schunck's
    decay -- together with a mask or a convolution kernel: ``post_process`` edits ``prev_flow`` in place only up to
    the filters there (``numpy.multiply`` / ``numpy.stack`` copy, ``source.py:337-348``), and fully in place without
    them (quirk Q5).  Also the documented ``math`` / ``random`` / ``numpy`` names inside filter expressions."""
    import cv2
    import json
    from transflow.flow.sources.source import FlowSource
    from transflow_b200.synthetic import synthetic_clip

    fh, fw, n = 64, 96, 7
    clip = synthetic_clip(fh, fw, n, seed=5)
    avi = os.path.join(tmp, "clip_lock.avi")
    vw = cv2.VideoWriter(avi, cv2.VideoWriter_fourcc(*"FFV1"), 25, (fw, fh))
    for f in clip:
        vw.write(f)
    vw.release()
    grad = np.clip(np.add.outer(np.linspace(40, 220, fh), np.linspace(0, 35, fw)), 0, 255).astype(np.uint8)
    p_mask = os.path.join(tmp, "lock_mask.png")
    save_mask_png(grad, p_mask)
    p_kernel = os.path.join(tmp, "box3.npy")
    np.save(p_kernel, np.full((3, 3), 1 / 9))
    cfgs = {"fb": dict(method="farneback"),
            "hs": dict(method="horn-schunck", hs_alpha=10.0, hs_iterations=2, hs_decay=0.9, hs_delta=1.0)}
    for k, c in cfgs.items():
        with open(os.path.join(tmp, k + ".json"), "w") as fp:
            json.dump(c, fp)
    cases = {
        "stay_mask_fw": dict(cfg="fb", lock_expr="(0.04,0.12),(100,1)", lock_mode="stay", mask_path=p_mask,
                             flow_filters="scale=1.5", direction="forward"),
        "stay_kernel_bw": dict(cfg="fb", lock_expr="(0.04,0.12),(100,1)", lock_mode="stay", kernel_path=p_kernel,
                               direction="backward"),
        "stay_plain_fw": dict(cfg="fb", lock_expr="(0.04,0.12),(100,1)", lock_mode="stay", flow_filters="scale=0.9",
                              direction="forward"),
        "skip_mask_bw": dict(cfg="fb", lock_expr="0.07<t<0.17", lock_mode="skip", mask_path=p_mask,
                             flow_filters="scale=2", direction="backward"),
        "hs_decay_mask_fw": dict(cfg="hs", mask_path=p_mask, direction="forward"),
        "hs_decay_kernel_bw": dict(cfg="hs", kernel_path=p_kernel, flow_filters="scale=1.25", direction="backward"),
        "usage_math": dict(cfg="fb", flow_filters="scale=1-math.exp(-.5*(t+1));clip=2+math.sin(t)",
                           direction="backward"),
        "usage_polar_numpy": dict(cfg="fb", flow_filters="polar=r*(1+numpy.abs(numpy.sin(a))):a+numpy.pi/8*t",
                                  direction="backward"),
    }
    from transflow.utils import load_float_mask
    out = {"clip": clip, "mask": load_float_mask(p_mask), "kernel/box3": np.load(p_kernel), "framerate": np.array(25.0)}
    for k, c in cfgs.items():
        out["config/" + k] = np.array(json.dumps(c))
    for name, kw in cases.items():
        kw = dict(kw)
        cfg = kw.pop("cfg")
        with FlowSource.from_args(avi, cv_config=os.path.join(tmp, cfg + ".json"), **kw) as src:
            import itertools
            flows = np.stack([np.array(f) for f in itertools.islice(src, 9)])   # (the stay tuples extend the length)
        out[f"{name}/flows"] = flows.astype(np.float32)
        out[f"{name}/args"] = np.array(json.dumps(dict({k: (os.path.basename(v) if k.endswith("_path") else v)
                                                         for k, v in kw.items()}, cfg=cfg)))
        print(name, flows.shape)
    np.savez_compressed(os.path.join(HERE, "lock_golden.npz"), **out)
    print("lock cases:", len(cases))


if __name__ == "__main__":
    only = sys.argv[1:]
    if not only or "render" in only:
        render_cases()
    with tempfile.TemporaryDirectory() as tmp:
        if not only or "compositor" in only:
            compositor_cases(tmp)
            kat_cases()
        if not only or "flow" in only:
            flow_cases(tmp)
        if not only or "postprocess" in only:
            postprocess_cases(tmp)
        if not only or "merge" in only:
            merge_cases(tmp)
        if not only or "lock" in only:
            lock_cases(tmp)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
