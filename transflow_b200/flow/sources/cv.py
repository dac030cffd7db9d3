"""``CvFlowSource`` on device (drop-in for ``transflow/flow/sources/cv.py:271-524``).

Same ``CvFlowConfig`` fields (``method, fb_*, hs_*, lk_*``) and JSON files, same
``Builder`` / ``rewind`` / ``next`` contract.  Frames still arrive from ``cv2.VideoCapture``
(host decode is out of scope) or from an in-memory ``ArrayCapture``; every
arithmetic step after the decode -- BGR->gray, Farneback / Horn-Schunck / Lucas-Kanade, the
post-process -- runs in CUDA.  A frame's Farneback pyramid + polynomial expansion is computed
once and reused as the other side of the next pair.
"""
import collections
import enum
import json
import os
import re

import numpy as np
import torch

from .source import FlowSource
from ... import ops


class CvFlowConfig:

    _DEFAULTS = dict(fb_pyr_scale=0.5, fb_levels=3, fb_winsize=15, fb_iterations=3, fb_poly_n=5,
                     fb_poly_sigma=1.2, fb_flags=0, hs_alpha=1, hs_iterations=3, hs_decay=0, hs_delta=1,
                     lk_window_size=15, lk_max_level=2, lk_step=1)

    def __init__(self, method="farneback", show_window=False, **params):
        unknown = set(params) - set(self._DEFAULTS)
        if unknown:
            raise TypeError(f"Unexpected CvFlowConfig fields: {sorted(unknown)}")
        self.method = CvFlowSource.Method.from_string(method) if isinstance(method, str) else method
        for name, default in self._DEFAULTS.items():
            setattr(self, name, params.get(name, default))
        self.show_window = show_window  # the Qt tuning window is out of scope (PySide6 UI)
        self.window = None

    def start(self):
        if self.show_window:
            raise NotImplementedError("the Qt configuration window is not part of the accelerated path")

    def update(self, attrname, value):
        if attrname == "method" and isinstance(value, str):
            value = CvFlowSource.Method.from_string(value)
        setattr(self, attrname, value)

    def reset(self):
        self.method = CvFlowSource.Method.FARNEBACK
        for name, default in self._DEFAULTS.items():
            setattr(self, name, default)
        self.hs_decay = 0.95  # the reference's reset() differs from its constructor here (cv.py:321)

    def to_dict(self):
        d = {"method": CvFlowSource.Method.to_string(self.method)}
        d.update({name: getattr(self, name) for name in self._DEFAULTS})
        return d

    def to_file(self, path: str):
        with open(path, "w", encoding="utf8") as file:
            json.dump(self.to_dict(), file, indent=4)

    @classmethod
    def from_file(cls, path: str):
        with open(path, "r", encoding="utf8") as file:
            return cls(**json.load(file))


class ArrayCapture:
    """In-memory stand-in for ``cv2.VideoCapture`` over a ``(T, H, W, 3) uint8`` BGR array
    (NumPy, or a CUDA tensor to skip the per-frame upload)."""

    def __init__(self, frames, fps: float = 25.0):
        self.frames = frames
        self.fps = float(fps)
        self.pos = 0

    def read(self):
        if self.pos >= len(self.frames):
            return False, None
        frame = self.frames[self.pos]
        self.pos += 1
        return True, frame

    def get(self, prop):
        import cv2
        if prop == cv2.CAP_PROP_FRAME_WIDTH:
            return self.frames.shape[2]
        if prop == cv2.CAP_PROP_FRAME_HEIGHT:
            return self.frames.shape[1]
        if prop == cv2.CAP_PROP_FPS:
            return self.fps
        if prop == cv2.CAP_PROP_FRAME_COUNT:
            return len(self.frames)
        return 0

    def set(self, prop, value):
        import cv2
        if prop == cv2.CAP_PROP_POS_MSEC:
            self.pos = int(round(value / 1000.0 * self.fps))
        return True

    def release(self):
        pass


class _Uploaded:
    """A frame on its way into a slot of the device ring."""
    __slots__ = ("frame", "ready", "slot")

    def __init__(self, frame, ready, slot):
        self.frame, self.ready, self.slot = frame, ready, slot


class CvFlowSource(FlowSource):

    @enum.unique
    class Method(enum.Enum):
        FARNEBACK = 0
        HORN_SCHUNCK = 1
        LUKAS_KANADE = 2
        LITEFLOWNET = 3

        @classmethod
        def from_string(cls, string: str):
            for m, s in _METHOD_NAMES.items():
                if s == string:
                    return cls[m]
            raise ValueError(f"Invalid Flow Method: {string}")

        @staticmethod
        def to_string(method):
            if method.name in _METHOD_NAMES:
                return _METHOD_NAMES[method.name]
            raise ValueError(f"Unknown flow method {method}")

    class Builder(FlowSource.Builder):

        def __init__(self, file, config: CvFlowConfig, size=None, **kwargs):
            super().__init__(**kwargs)
            self.file = file
            self.config = config
            self.size = size
            self.capture = None

        @property
        def cls(self):
            return CvFlowSource

        def build(self):
            import cv2
            if hasattr(self.file, "read") and hasattr(self.file, "get"):
                self.capture = self.file          # an already-open capture (e.g. ArrayCapture)
            elif re.match(r"\d+", self.file):
                self.capture = cv2.VideoCapture(int(self.file))  # webcam index
            else:
                self.capture = cv2.VideoCapture(self.file)
            if self.size is not None:
                self.capture.set(cv2.CAP_PROP_FRAME_WIDTH, self.size[0])
                self.capture.set(cv2.CAP_PROP_FRAME_HEIGHT, self.size[1])
            self.width = int(self.capture.get(cv2.CAP_PROP_FRAME_WIDTH))
            self.height = int(self.capture.get(cv2.CAP_PROP_FRAME_HEIGHT))
            self.framerate = float(self.capture.get(cv2.CAP_PROP_FPS))
            self.base_length = int(self.capture.get(cv2.CAP_PROP_FRAME_COUNT)) - 1
            super().build()

        def args(self):
            return [self.capture, self.config, *FlowSource.Builder.args(self)]

    def __init__(self, capture, config: CvFlowConfig, *args, **kwargs):
        self.config = config
        self.capture = capture
        self.prev_gray = None
        self._slot = 0          # Farneback slot holding prev_gray's pyramid + expansion
        self._engine = None
        self._engine_key = None
        self._stage = None      # pinned host staging for uploads
        self._copy_stream = None
        self._dev_ring = None   # device frames the uploads land in
        self._lookahead = None
        #: read one frame ahead so its H2D copy overlaps the current frame's kernels
        self.prefetch = True
        #: Farneback pairs kept in flight (each on its own stream and handle lane).  The reference's flow process
        #: also runs ahead of its consumer through a queue (pipeline.py:326); pair t does not depend on pair t-1
        #: (fb_flags == 0), so the second pair fills the SMs the first leaves idle.  1 = strictly one pair at a time.
        self.pairs_in_flight = int(os.environ.get("TFB200_FB_LANES", "2"))
        self._inflight = collections.deque()
        self._lane_streams = None
        self._submitted = 0     # pairs submitted since the last (re)prepare; the previous frame is in slot _submitted % 3
        self._eos = False
        self.config.start()
        FlowSource.__init__(self, *args, **kwargs)

    def validate(self):
        super().validate()
        self.assert_type("config", CvFlowConfig)
        if not (hasattr(self.capture, "read") and hasattr(self.capture, "release")):
            raise ValueError("Attribute capture has incorrect type")

    # -- frame prep (cv.py:461-466): resize NEAREST on the host if needed, upload, gray on device ------
    def _upload(self, frame):
        """Start the H2D copy of a decoded BGR frame on the side stream into the next slot of a preallocated ring of
        device frames -> ``_Uploaded``.  CUDA tensors pass through, pinned CPU tensors are copied directly, NumPy
        frames go through a pinned staging buffer."""
        if isinstance(frame, _Uploaded) or (isinstance(frame, torch.Tensor) and frame.is_cuda):
            return frame
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        resize_from = None
        if isinstance(frame, torch.Tensor) and frame.is_pinned():
            src = frame
        else:
            if isinstance(frame, torch.Tensor):
                frame = frame.numpy()
            if frame.shape[1] != self.width or frame.shape[0] != self.height:
                # the capture delivers another size (cv.py:464 resizes with INTER_NEAREST): upload the frame as it is
                # and resize on the device
                resize_from = torch.from_numpy(np.ascontiguousarray(frame)).pin_memory()
            if resize_from is None:
                if self._stage is None:
                    self._stage = [torch.empty((self.height, self.width, 3), dtype=torch.uint8).pin_memory()
                                   for _ in range(2)]
                    self._stage_events = [None, None]
                k = self._stage_index = (getattr(self, "_stage_index", 0) + 1) & 1
                if self._stage_events[k] is not None:
                    self._stage_events[k].synchronize()     # the previous copy out of this buffer is done
                self._stage[k].numpy()[...] = frame
                src = self._stage[k]
            else:
                src = resize_from
        if tuple(src.shape[:2]) != (self.height, self.width):
            resize_from = src
        # preallocated ring of device frames (no allocation per frame): a slot is overwritten once the
        # gray conversion that read its previous frame has run
        if self._dev_ring is None:
            n = max(2, int(self.pairs_in_flight)) + 2
            self._dev_ring = [torch.empty((self.height, self.width, 3), dtype=torch.uint8, device="cuda")
                              for _ in range(n)]
            self._dev_consumed = [None] * n
            self._dev_index = 0
        slot = self._dev_index
        self._dev_index = (slot + 1) % len(self._dev_ring)
        dev = self._dev_ring[slot]
        with torch.cuda.stream(self._copy_stream):
            if self._dev_consumed[slot] is not None:
                self._copy_stream.wait_event(self._dev_consumed[slot])
            if resize_from is not None:
                raw = resize_from.cuda(non_blocking=True)       # (rare path: one allocation per frame)
                ops.resize_nearest_bgr(raw, self.height, self.width, out=dev)
                raw.record_stream(self._copy_stream)
                self._keep_pinned = resize_from                  # alive until the next upload
            else:
                dev.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        if resize_from is None and src is not frame and self._stage is not None:
            self._stage_events[self._stage_index] = ev
        return _Uploaded(dev, ev, slot)

    def _gray_on_device(self, frame) -> torch.Tensor:
        up = self._upload(frame)
        if not isinstance(up, _Uploaded):
            return ops.gray_from_bgr(up)
        torch.cuda.current_stream().wait_event(up.ready)
        gray = ops.gray_from_bgr(up.frame)
        done = torch.cuda.Event()
        done.record()
        self._dev_consumed[up.slot] = done
        return gray

    def _read_frame(self):
        """capture.read() with a one-frame look-ahead: the next frame's upload overlaps this
        frame's kernels."""
        if self._lookahead is not None:
            item = self._lookahead
            self._lookahead = None
        else:
            ok, frame = self.capture.read()
            item = (ok, self._upload(frame) if (ok and frame is not None) else None)
        if self.prefetch and item[0] and item[1] is not None:
            ok, frame = self.capture.read()
            self._lookahead = (ok, self._upload(frame) if (ok and frame is not None) else None)
        return item

    def _method_engine(self):
        c = self.config
        m = c.method
        if m == CvFlowSource.Method.FARNEBACK:
            key = ("fb", c.fb_pyr_scale, c.fb_levels, c.fb_winsize, c.fb_iterations, c.fb_poly_n, c.fb_poly_sigma,
                   c.fb_flags)
            make = lambda: ops.Farneback(self.height, self.width, c.fb_pyr_scale, c.fb_levels, c.fb_winsize,  # noqa: E731
                                         c.fb_iterations, c.fb_poly_n, c.fb_poly_sigma, c.fb_flags,
                                         lanes=max(1, min(2, self.pairs_in_flight)))
        elif m == CvFlowSource.Method.HORN_SCHUNCK:
            key = ("hs",)
            make = lambda: ops.HornSchunck(self.height, self.width)  # noqa: E731
        elif m == CvFlowSource.Method.LUKAS_KANADE:
            key = ("lk", c.lk_window_size, c.lk_max_level, c.lk_step)
            make = lambda: ops.LucasKanade(self.height, self.width, c.lk_window_size, c.lk_max_level, c.lk_step)  # noqa: E731
        elif m == CvFlowSource.Method.LITEFLOWNET:
            raise ImportError("LiteFlowNet method cannot be used: it is outside the accelerated path (SURVEY.md #9)")
        else:
            raise ValueError(f"Unknown flow method '{m}'")
        if key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            self._engine = make()
            self._engine_key = key
            self._prepared = False
        return self._engine

    def rewind(self):
        import cv2
        FlowSource.rewind(self)
        self.capture.set(cv2.CAP_PROP_POS_MSEC, 0)
        self._lookahead = None
        frame = None
        for i in range(self.input_frame_index + 1):
            success, frame = self.capture.read()
            if not success or frame is None:
                raise RuntimeError(f"An error occurred while reading frame at index {i}")
        self.prev_gray = self._gray_on_device(frame)
        self._prepared = False
        self.prev_flow = None
        self._drain()

    def _drain(self):
        for _, _, ev in self._inflight:
            ev.synchronize()
        self._inflight.clear()
        self._submitted = 0
        self._eos = False

    def _next_farneback_lanes(self, engine, forward: bool) -> torch.Tensor:
        """Farneback with up to two pairs in flight: frame t's expansion lives in slot t % 3, pair (t, t+1) is
        solved on lane t % 2 / its own stream; flows are handed out in frame order."""
        main = torch.cuda.current_stream()
        if self._lane_streams is None:
            self._lane_streams = [torch.cuda.Stream() for _ in range(2)]
        if not self._prepared:
            self._drain()
            engine.prepare(0, self.prev_gray)
            self._prepared = True
            for s in self._lane_streams:
                s.wait_stream(main)
        while len(self._inflight) < self.pairs_in_flight and not self._eos:
            success, frame = self._read_frame()
            if frame is None or not success:
                self._eos = True
                break
            n = self._submitted
            lane, old, new = n & 1, n % 3, (n + 1) % 3
            with torch.cuda.stream(self._lane_streams[lane]):
                gray = self._gray_on_device(frame)
                flow = (engine.step(new, gray, old, new, lane=lane) if forward
                        else engine.step(new, gray, new, old, lane=lane))
                ev = torch.cuda.Event()
                ev.record()
            self._inflight.append((flow, gray, ev))
            self._submitted = n + 1
        if not self._inflight:
            raise StopIteration
        flow, gray, ev = self._inflight.popleft()
        main.wait_event(ev)
        flow.record_stream(main)
        self.prev_gray = gray
        return flow

    def _reads_prev_flow(self) -> bool:
        # Horn-Schunck starts from the previous (post-processed, aliased) flow
        return super()._reads_prev_flow() or self.config.method == CvFlowSource.Method.HORN_SCHUNCK

    def next(self) -> torch.Tensor:
        if (self.config.method == CvFlowSource.Method.FARNEBACK and self.pairs_in_flight > 1
                and self.prev_gray is not None
                and self.direction in (FlowSource.Direction.FORWARD, FlowSource.Direction.BACKWARD)):
            return self._next_farneback_lanes(self._method_engine(),
                                              self.direction == FlowSource.Direction.FORWARD)
        success, frame = self._read_frame()
        if frame is None or not success:
            raise StopIteration
        gray = self._gray_on_device(frame)
        if self.direction not in (FlowSource.Direction.FORWARD, FlowSource.Direction.BACKWARD):
            raise ValueError(f"Invalid flow direction '{self.direction}'")
        if self.prev_gray is None:
            raise ValueError("Missing reference frames")
        forward = self.direction == FlowSource.Direction.FORWARD
        engine = self._method_engine()
        c = self.config
        if c.method == CvFlowSource.Method.FARNEBACK:
            if not self._prepared:
                engine.prepare(self._slot, self.prev_gray)
                self._prepared = True
            cur = self._slot ^ 1
            # the new frame's pyramid / expansion overlaps the coarse levels of the solve
            flow = (engine.step(cur, gray, self._slot, cur) if forward else engine.step(cur, gray, cur, self._slot))
            self._slot = cur
        else:
            left, right = (self.prev_gray, gray) if forward else (gray, self.prev_gray)
            if c.method == CvFlowSource.Method.HORN_SCHUNCK:
                prev = None if self.prev_flow is None else self.prev_flow.clone()
                flow = engine(left, right, prev, c.hs_alpha, c.hs_iterations, c.hs_decay, c.hs_delta)
            else:
                flow = engine(left, right)
        self.prev_gray = gray
        return flow

    def close(self):
        self._drain()
        self.capture.release()
        if self._engine is not None:
            self._engine.close()
            self._engine = None


_METHOD_NAMES = {"FARNEBACK": "farneback", "HORN_SCHUNCK": "horn-schunck", "LUKAS_KANADE": "lukas-kanade",
                 "LITEFLOWNET": "liteflownet"}
