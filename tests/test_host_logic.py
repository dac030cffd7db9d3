"""Host-side logic that needs no GPU: the video pixmap source (seek / repeat / length bookkeeping of
``transflow/pixmap/cv.py``) checked against the REAL reference class where it is importable (the build container),
and against the expected frame order everywhere."""
import json
import os
import sys

import numpy as np
import pytest


class HostCapture:
    """cv2.VideoCapture stand-in over a (T, H, W, 3) BGR array (no torch / CUDA import)."""

    def __init__(self, frames, fps=25.0):
        self.frames, self.fps, self.pos = frames, float(fps), 0

    def read(self):
        if self.pos >= len(self.frames):
            return False, None
        self.pos += 1
        return True, self.frames[self.pos - 1]

    def get(self, prop):
        import cv2
        return {cv2.CAP_PROP_FRAME_WIDTH: self.frames.shape[2], cv2.CAP_PROP_FRAME_HEIGHT: self.frames.shape[1],
                cv2.CAP_PROP_FPS: self.fps, cv2.CAP_PROP_FRAME_COUNT: len(self.frames)}.get(prop, 0)

    def set(self, prop, value):
        import cv2
        if prop == cv2.CAP_PROP_POS_MSEC:
            self.pos = int(round(value / 1000.0 * self.fps))
        return True

    def isOpened(self):
        return True

    def release(self):
        pass


def make_frames(n=5, h=6, w=8):
    rng = np.random.default_rng(4)
    return rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("kw", [dict(), dict(seek=2), dict(repeat=3), dict(seek=1, repeat=2), dict(seek_time=0.08, repeat=2)])
def test_video_pixmap_source_order_and_length(kw):
    from transflow_b200.pixmap.cv import CvPixmapSource
    frames = make_frames()
    with CvPixmapSource(HostCapture(frames), **kw) as src:
        assert (src.width, src.height, src.framerate) == (8, 6, 25)
        got = list(src)
        length = src.length
    seek = kw.get("seek", int(kw["seek_time"] * 25) if "seek_time" in kw else 0)
    one_pass = [f[:, :, ::-1] for f in frames[seek:]]
    want = one_pass * kw.get("repeat", 1)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    # the announced length only accounts for a TIME seek (pixmap/cv.py:39-44 of the reference)
    assert length == (len(frames) - (seek if "seek_time" in kw else 0)) * kw.get("repeat", 1)


@pytest.mark.skipif(not os.path.isdir("/root/reference/transflow"), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("kw", [dict(), dict(seek=2, repeat=2), dict(seek_time=0.04, repeat=3)])
def test_video_pixmap_source_matches_reference_class(kw, tmp_path):
    """Same FFV1 clip through the reference's CvPixmapSource and ours: identical frames and metadata."""
    import cv2
    sys.path.insert(0, "/root/reference")
    sys.dont_write_bytecode = True
    try:
        from transflow.pixmap.cv import CvPixmapSource as Ref
    finally:
        sys.path.remove("/root/reference")
    from transflow_b200.pixmap.cv import CvPixmapSource
    frames = make_frames(6, 32, 48)
    path = str(tmp_path / "clip.avi")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 25, (48, 32))
    for f in frames:
        vw.write(f)
    vw.release()
    with Ref(path, **kw) as ref:
        want = [np.array(f) for f in ref]
        meta = (ref.width, ref.height, ref.framerate, ref.length)
    with CvPixmapSource(path, **kw) as src:
        got = list(src)
        assert (src.width, src.height, src.framerate, src.length) == meta
    assert len(got) == len(want)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


def test_bench_clock_sampler_counts_only_the_timed_region():
    """bench.ClockSampler: samples that arrive before mark_begin / after mark_end do not count; when the timed region
    holds fewer than two samples the same load is kept until two have arrived and the line says so."""
    import bench

    class FakeProc:
        def terminate(self):
            pass

    def line(mhz, power_cap="Not Active"):
        return f"{mhz}, 1965, 700.0, Not Active, Not Active, Not Active, {power_cap}"

    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    s.lines = [line(300), line(400)]                 # start-up, before the timed region
    s.mark_begin()
    s.lines += [line(1950), line(1965, "Active"), line(1965)]
    s.mark_end()
    s.lines += [line(210)]                           # after it
    assert s.samples_in_window() == 3
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["window"] == "timed region"

    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    s.lines = [line(300)]
    s.mark_begin()
    s.mark_end()                                     # a timed region shorter than the sampling period
    assert s.samples_in_window() == 0
    calls = []

    def keep_busy():
        calls.append(1)
        s.lines.append(line(1900 + len(calls)))
    s.extend(keep_busy, want=2, limit_s=5.0)
    out = s.stop()
    assert len(calls) == 2 and out["samples"] == 2 and out["sm_mhz"] == 1901.5
    assert out["window"].startswith("same load")


def test_pipeline_config_round_trip_and_secondary_paths():
    """``Config.todict`` / ``fromdict`` (checkpoint meta.json, reference config.py:258-323) and the checkpoint /
    archive path rule ``get_secondary_output_path`` (config.py:325-341)."""
    from transflow_b200.config import LayerConfig, PixmapSourceConfig
    from transflow_b200.flow import FlowSource
    from transflow_b200.flow.sources.cv import CvFlowConfig
    from transflow_b200.pipeline import Config
    cfg = Config("clips/a.mp4", mask_path="m.png", cv_config=CvFlowConfig(method="horn-schunck", hs_decay=0.5),
                 flow_filters="scale=2;clip=1+t", direction="backward", seek_time="0:01.5", duration_time=3,
                 repeat=2, lock_expr="(1,2)", lock_mode="stay",
                 pixmap_sources=[PixmapSourceConfig("p.jpg", layers=[0, 1], introduction_path="border:1")],
                 layers=[LayerConfig(0, "moveref", reset_mode="random", reset_random_factor=0.5), LayerConfig(1, "sum")],
                 compositor_background="#102030", output_path="out/x-%d.png", size=(640, 360), seed=5,
                 extra_flow_paths=["b.mp4"], flows_merging_function="sum", view_flow=True, render_scale=0.5,
                 render_colors="#000,#fff")
    d = json.loads(json.dumps(cfg.todict()))          # what lands in meta.json
    back = Config.fromdict(d)
    assert back.todict() == cfg.todict()
    assert back.seek_time == 1.5 and back.duration_time == 3 and back.size == (640, 360) and back.seed == 5
    assert FlowSource.Direction.from_arg(back.direction) == FlowSource.Direction.BACKWARD
    assert isinstance(back.cv_config, CvFlowConfig) and back.cv_config.hs_decay == 0.5
    assert back.layers[0].reset_random_factor == 0.5 and back.pixmap_sources[0].layers == [0, 1]
    assert back.extra_flow_paths == ["b.mp4"] and back.flows_merging_function == "sum" and back.view_flow
    assert cfg.get_secondary_output_path("_00005.ckpt.zip") == "out/x-%d_00005.ckpt.zip"
    assert Config("v/clip.flow.zip").get_secondary_output_path(".x") == "v/clip.x"
    assert Config("v/clip_00012.ckpt.zip").get_secondary_output_path(".y") == "v/clip_00012.y"
    assert Config("v/clip.mp4", output_path="o/r.001.mp4").get_secondary_output_path(".flow.zip") == "o/r.flow.zip"
    ref_root = "/root/reference"
    if os.path.isdir(ref_root):                        # same rule as the reference's own Config
        import importlib
        sys.path.insert(0, ref_root)
        try:
            RC = importlib.import_module("transflow.config").Config
            for flow, out in (("v/clip.mp4", "o/r.001.mp4"), ("v/clip.flow.zip", None), ("a/b.mp4", "out/x-%d.png")):
                assert (Config(flow, output_path=out).get_secondary_output_path("_s.zip")
                        == RC(flow, output_path=out).get_secondary_output_path("_s.zip"))
        finally:
            sys.path.remove(ref_root)


def test_polar_expressions_run_numpy_functions_on_tensors():
    """Inside a ``polar`` filter ``numpy.f(a)`` with a tensor argument runs as ``torch.f`` (the reference hands the
    expressions NumPy arrays, filters.py:79-84); constants and scalar calls stay NumPy; a function torch does not have
    raises unless host evaluation of the user's lambda is switched on explicitly."""
    import torch
    from transflow_b200.flow.filters import PolarFlowFilter, _DeviceNumpy
    from transflow_b200.utils import parse_lambda_expression
    a = torch.linspace(-3, 3, 7)
    npm = _DeviceNumpy()
    assert torch.allclose(npm.sin(a), torch.sin(a)) and torch.allclose(npm.arctan2(a, a + 1), torch.atan2(a, a + 1))
    assert torch.allclose(npm.maximum(a, 0.5), torch.clamp(a, min=0.5)) and npm.pi == np.pi and npm.sqrt(4.0) == 2.0
    f = parse_lambda_expression("r*(1+numpy.abs(numpy.sin(a))) + numpy.pi*t", ("t", "r", "a"), numpy_module=npm)
    assert torch.allclose(f(0.5, a, a), a * (1 + torch.abs(torch.sin(a))) + np.pi * 0.5)
    # a function torch does not have: the filter evaluates on host copies, like the reference
    flow = torch.stack([torch.linspace(-2, 2, 12).reshape(3, 4), torch.linspace(1, 3, 12).reshape(3, 4)], dim=-1).clone()
    want = flow.clone().numpy()
    r, th = np.linalg.norm(want.reshape(-1, 2), axis=1).reshape(3, 4), np.arctan2(want[..., 1], want[..., 0])
    nr = np.unwrap(th) * 0 + r * 2
    want[..., 1], want[..., 0] = nr * np.sin(th + 0.25), nr * np.cos(th + 0.25)
    flt = PolarFlowFilter(("numpy.unwrap(a)*0 + r*2", "a+t"))
    with pytest.raises(NotImplementedError):          # no silent host path
        flt.apply(flow.clone(), 0.25)
    flt.allow_host_expressions = True                  # explicit opt-in: only the user's lambda runs on the host
    flt.apply(flow, 0.25)
    np.testing.assert_allclose(flow.numpy(), want, atol=1e-5)


def test_gradient_pixmap_tree_draws_match_reference():
    """``GradientPixmapSource.generate`` is written as its own grammar; a seeded run must consume ``random`` exactly
    as the reference does (still.py:94-118) so that a given seed paints the same picture."""
    import random
    ref_root = "/root/reference"
    if not os.path.isdir(ref_root):
        pytest.skip("the reference tree is only present in the build container")
    import importlib
    from transflow_b200.pixmap.still import GradientPixmapSource as G
    sys.path.insert(0, ref_root)
    try:
        R = importlib.import_module("transflow.pixmap.still").GradientPixmapSource
        for seed in range(12):
            for depth in (0, 1, 2, 4):
                for kind in (G.NODE_MIX, G.NODE_TRIPLE, G.NODE_Z, G.NODE_B):
                    random.seed(seed)
                    mine = G.generate(G.__new__(G), kind, depth)
                    random.seed(seed)
                    theirs = R.generate(R.__new__(R), kind, depth)
                    assert mine == theirs, (seed, depth, kind)
    finally:
        sys.path.remove(ref_root)
    with pytest.raises(ValueError):
        G.generate(G.__new__(G), 99, 2)
