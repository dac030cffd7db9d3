"""Flow source plugin surface (drop-in for ``transflow/flow/sources/source.py``).

Kept from the reference: ``FlowSource.from_args(...) -> Builder`` (``source.py:365-411``); the
``Builder`` context manager whose ``__enter__`` runs ``build()``, constructs
``cls(*args(), **kwargs())`` and calls ``validate()`` (``:199-204``); ``width / height /
framerate / length``; iteration yielding ``(H, W, 2) float32`` flows; ``next() / rewind() /
close()``; the seek / duration / repeat / lock bookkeeping (``:125-197``, ``:293-321``); and
``post_process`` (``:337-363``), which runs on the device here.

Yielded flows are NumPy arrays by default (what ``pipeline.py`` expects from the queue); set
``source.output = "device"`` to receive CUDA tensors and keep the frame on the GPU.
"""
import enum
import logging
import os
import warnings

import numpy as np
import torch

from ..filters import FlowFilter
from ... import ops
from ...utils import load_float_mask, parse_lambda_expression

logger = logging.getLogger(__name__)


class FlowSource:

    @enum.unique
    class Direction(enum.Enum):
        FORWARD = 0   # past to present
        BACKWARD = 1  # present to past

        @classmethod
        def from_arg(cls, arg):
            if arg is None:
                return cls.FORWARD
            if isinstance(arg, cls):
                return arg
            if isinstance(arg, enum.Enum):      # the reference's own Direction enum (same member names)
                return cls[arg.name]
            if isinstance(arg, int):
                return cls(arg)
            table = {"forward": cls.FORWARD, "backward": cls.BACKWARD}
            if arg in table:
                return table[arg]
            raise ValueError(f"Invalid Flow Direction: {arg}")

    @enum.unique
    class LockMode(enum.Enum):
        STAY = 0
        SKIP = 1

        @classmethod
        def from_arg(cls, arg):
            if arg is None:
                return cls.STAY
            if isinstance(arg, cls):
                return arg
            if isinstance(arg, enum.Enum):      # the reference's own LockMode enum
                return cls[arg.name]
            if isinstance(arg, int):
                return cls(arg)
            table = {"stay": cls.STAY, "skip": cls.SKIP}
            if arg in table:
                return table[arg]
            raise ValueError(f"Invalid Lock Mode: {arg}")

    class Builder:
        """Collects the arguments, resolves frame ranges, then builds the source on ``__enter__``."""

        def __init__(self, direction="backward", mask_path=None, kernel_path=None, flow_filters=None,
                     seek_ckpt=None, seek_time=None, duration_time=None, repeat=1, lock_expr=None,
                     lock_mode="stay"):
            self.direction = FlowSource.Direction.from_arg(direction)
            self.width = None
            self.height = None
            self.framerate = 30
            self.mask_path = mask_path
            self.mask = None
            self.kernel_path = kernel_path
            self.kernel = None
            self.flow_filters = []
            self.flow_filters_string = flow_filters
            self.seek_ckpt = seek_ckpt
            self.seek_time = seek_time
            self.duration_time = duration_time
            self.is_stream = False
            self.base_length = None
            self.length = None
            self.start_frame = 0
            self.ckpt_start_frame = 0
            self.end_frame = 0
            self.repeat = repeat
            self.lock_expr_string = lock_expr
            self.lock_expr_stay = None
            self.lock_expr_skip = None
            self.lock_mode = FlowSource.LockMode.from_arg(lock_mode)
            self.source = None

        @property
        def cls(self):
            return FlowSource

        def args(self) -> list:
            return [self.direction, self.width, self.height, self.framerate, self.length, self.start_frame,
                    self.ckpt_start_frame, self.end_frame]

        def kwargs(self) -> dict:
            return dict(mask=self.mask, kernel=self.kernel, flow_filters=self.flow_filters,
                        lock_mode=self.lock_mode, lock_expr_stay=self.lock_expr_stay,
                        lock_expr_skip=self.lock_expr_skip)

        def _parse_options(self):
            if self.mask_path is not None:
                m = load_float_mask(self.mask_path)
                self.mask = m.reshape((*m.shape, 1))
            if self.kernel_path is not None:
                self.kernel = np.load(self.kernel_path)
            if self.lock_expr_string is not None:
                if self.lock_mode == FlowSource.LockMode.STAY:
                    text = self.lock_expr_string
                    if "(" not in text:
                        text = f"({text})"
                    self.lock_expr_stay = tuple(eval(f"[{text},]"))  # noqa: S307 - trusted CLI input
                else:
                    self.lock_expr_skip = parse_lambda_expression(self.lock_expr_string)
            if self.flow_filters_string is not None:
                self.flow_filters = []
                for item in self.flow_filters_string.strip().split(";"):
                    if not item.strip():
                        continue
                    name, _, rest = item.partition("=")
                    self.flow_filters.append(FlowFilter.from_args(name.strip(), tuple(rest.strip().split(":"))))

        def _resolve_range(self):
            if self.base_length is not None and self.base_length <= 0:
                self.base_length = None
            self.is_stream = self.base_length is None
            if self.is_stream and self.repeat > 1:
                warnings.warn("Flow source is a stream, cannot repeat it!")
                self.repeat = 1
            if self.is_stream and self.seek_time is not None and self.seek_time > 0:
                warnings.warn("Flow source is a stream, seek time is ignored!")
                self.seek_time = None
            first = int(self.seek_time * self.framerate) if (self.seek_time is not None and not self.is_stream) else 0
            if self.duration_time is not None:
                # round before flooring to dodge float noise
                last = first + int(round(self.duration_time * self.framerate, 3))
                if self.base_length is not None:
                    last = min(last, self.base_length)
            elif self.base_length is not None:
                last = self.base_length
            else:
                last = 0
            self.start_frame, self.end_frame = first, last
            if self.repeat == 0:
                self.length = None
            elif self.is_stream:
                self.length = last
            else:
                self.length = self.repeat * (last - first)
            if (self.length is not None and self.lock_mode == FlowSource.LockMode.STAY
                    and self.lock_expr_stay is not None):
                for _, hold in self.lock_expr_stay:
                    self.length += int(hold * self.framerate)
            self.ckpt_start_frame = first
            if self.seek_ckpt is not None:
                self.ckpt_start_frame += self.seek_ckpt % (last - first)

        def build(self):
            self._parse_options()
            self._resolve_range()

        def __enter__(self):
            self.build()
            self.source = self.cls(*self.args(), **self.kwargs())
            self.source.validate()
            logger.debug("Built '%s'", type(self.source).__name__)
            return self.source

        def __exit__(self, exc_type, exc_value, exc_traceback):
            if self.source is not None:
                self.source.close()

    # ---------------------------------------------------------------------------------------------
    def __init__(self, direction, width, height, framerate, length, start_frame, ckpt_start_frame, end_frame,
                 mask=None, kernel=None, flow_filters=None, lock_mode=None, lock_expr_stay=None,
                 lock_expr_skip=None):
        self.direction = direction
        self.width = width
        self.height = height
        self.framerate = framerate
        self.length = length
        self.end_frame = end_frame
        self.mask = mask
        self.kernel = kernel
        self.flow_filters = [] if flow_filters is None else flow_filters
        self.lock_mode = FlowSource.LockMode.STAY if lock_mode is None else lock_mode
        self.lock_expr_stay = lock_expr_stay
        self.lock_expr_skip = lock_expr_skip
        self.input_frame_index = 0
        self.output_frame_index = 0
        self.prev_flow = None
        self.lock_start = None
        self.lock_expr_stay_index = 0
        #: "numpy" (reference behaviour), "device" (CUDA tensors, no D2H per frame) or "claims": as "device", but a
        #: FORWARD flow whose only consumer is the compositor is handed out as ``ops.ForwardClaims`` (the scatter pass's
        #: claim plane; ``Compositor.step`` consumes it, ``.tensor()`` forms the flow for anyone else)
        self.output = "numpy"
        self.start_frame = ckpt_start_frame
        self.rewind()
        self.start_frame = start_frame
        self._post = None  # built lazily: needs a CUDA device

    def __len__(self):
        return self.length

    def assert_type(self, attr: str, *types: type):
        if not isinstance(getattr(self, attr), types):
            raise ValueError(f"Attribute {attr} has incorrect type {type(getattr(self, attr))}")

    def validate(self):
        self.assert_type("direction", FlowSource.Direction)
        self.assert_type("width", int)
        self.assert_type("height", int)
        self.assert_type("framerate", float)
        self.assert_type("length", int, type(None))
        self.assert_type("start_frame", int)
        self.assert_type("end_frame", int)
        self.assert_type("mask", np.ndarray, type(None))
        self.assert_type("kernel", np.ndarray, type(None))
        self.assert_type("flow_filters", list)
        self.assert_type("lock_mode", FlowSource.LockMode)
        self.assert_type("lock_expr_stay", tuple, type(None))

    @property
    def t(self) -> float:
        return 0 if self.framerate is None else self.output_frame_index / self.framerate

    def next(self):
        raise NotImplementedError()

    def rewind(self):
        self.input_frame_index = self.start_frame

    def read_next_flow(self):
        if self.input_frame_index == self.end_frame:
            self.rewind()
        flow = self.next()
        self.input_frame_index += 1
        return flow

    def _is_locked(self) -> bool:
        if self.lock_mode == FlowSource.LockMode.STAY and self.lock_expr_stay is not None:
            held = self.lock_start is not None
            locked = False
            if held:
                locked = (self.t - self.lock_start) < self.lock_expr_stay[self.lock_expr_stay_index][1]
                if not locked:
                    self.lock_expr_stay_index += 1
                    self.lock_start = None
            if not held or not locked:
                locked = self.t >= self.lock_expr_stay[self.lock_expr_stay_index][0]
                if locked:
                    self.lock_start = self.t
            return locked
        if self.lock_mode == FlowSource.LockMode.SKIP and self.lock_expr_skip is not None:
            return bool(self.lock_expr_skip(self.t))
        return False

    def __next__(self):
        if self.length is not None and self.output_frame_index >= self.length:
            raise StopIteration
        locked = self._is_locked()
        if locked:
            if self.prev_flow is None:
                raise RuntimeError("Flow is locked but has not been initialized. Maybe lock the flow later?")
            flow = self.prev_flow
        else:
            flow = self.read_next_flow()
        # prev_flow aliases ``flow``: post_process edits it in place exactly where the reference does (quirk Q5) --
        # the filters always; the clip / forward scatter only when no mask and no kernel is set
        self.prev_flow = flow
        if locked and self.lock_mode == FlowSource.LockMode.SKIP:
            self.read_next_flow()
        self.output_frame_index += 1
        out = self.post_process(flow)
        if self.output in ("device", "claims"):
            return out
        return out.cpu().numpy()

    def __iter__(self):
        return self

    def _apply_scalar_ops(self, flow, pending):
        arr, n = ops._pack_flow_ops(pending)
        ops.check(ops._lib.load().tf_flow_filters(ops.ptr(flow), arr, n, None, ops.ptr(flow), self.height,
                                                  self.width, ops.stream_ptr()))

    def _reads_prev_flow(self) -> bool:
        """Is ``prev_flow`` (which aliases the buffer ``post_process`` edits in place, quirk Q5) read again later?"""
        return self.lock_expr_stay is not None or self.lock_expr_skip is not None

    def post_process(self, raw):
        """filters -> mask -> kernel -> [forward: clip, round, scatter] -> clip, on the device.

        Runs of scalar filters (scale / threshold / clip) are folded into the post-process kernel together with
        the mask multiply; a ``polar`` filter (arbitrary array expressions) splits the run.  With a convolution
        kernel the order of the reference is kept: filters + mask, float64 convolution, then the clip / scatter.

        What happens to ``raw`` follows the reference (source.py:337-363), because ``prev_flow`` aliases it and a
        locked flow (or Horn-Schunck's decay) reads it again: without mask and kernel every step edits ``raw`` in
        place; with a mask or a kernel only the filters do (``numpy.multiply`` / ``numpy.stack`` copy there) and the
        result is a new tensor.
        """
        flow = raw if isinstance(raw, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(raw)).cuda()
        if self._post is None:
            mask = None
            if self.mask is not None:
                mask = torch.from_numpy(np.ascontiguousarray(self.mask, dtype=np.float32)).cuda()
            self._post = ops.PostProcess(self.height, self.width, self.direction == FlowSource.Direction.FORWARD,
                                         mask=mask, kernel=self.kernel)
        pending = []
        t = self.t
        for flt in self.flow_filters:
            if flt.kind is not None and len(pending) < ops.MAX_FLOW_OPS:
                pending.append(flt.op(t))
                continue
            if pending:
                self._apply_scalar_ops(flow, pending)
                pending = []
            if flt.kind is not None:
                pending.append(flt.op(t))
            else:
                flt.apply(flow, t)
        if (self.output == "claims" and self._post.forward and self._post.kernel is None
                and not self._reads_prev_flow()):
            # nobody reads ``raw`` again, so the in-place edits of the reference are unobservable: scatter pass only
            return self._post.claims(flow, ops=pending)
        if not self._post.copies:
            return self._post(flow, ops=pending)
        if pending:
            self._apply_scalar_ops(flow, pending)      # in place: prev_flow carries the filter edits only
        return self._post(flow, out=torch.empty_like(flow))

    @classmethod
    def from_args(cls, flow_path, use_mvs=False, mask_path=None, kernel_path=None, cv_config=None,
                  flow_filters=None, size=None, direction=None, seek_ckpt=None, seek_time=None,
                  duration_time=None, repeat=1, lock_expr=None, lock_mode=None):
        path = flow_path.split("::")[-1] if isinstance(flow_path, str) and "::" in flow_path else flow_path
        common = dict(direction=direction, mask_path=mask_path, kernel_path=kernel_path, flow_filters=flow_filters,
                      seek_ckpt=seek_ckpt, seek_time=seek_time, duration_time=duration_time, repeat=repeat,
                      lock_expr=lock_expr, lock_mode=lock_mode)
        if isinstance(path, str) and path.endswith(".flow.zip"):
            from .archive import ArchiveFlowSource
            return ArchiveFlowSource.Builder(path, **common)
        if use_mvs:
            raise NotImplementedError("motion-vector flow sources have no arithmetic to accelerate (SURVEY.md #11)")
        from .cv import CvFlowConfig, CvFlowSource
        if isinstance(cv_config, CvFlowConfig):
            config = cv_config
        elif cv_config is not None and cv_config != "window" and os.path.isfile(cv_config):
            config = CvFlowConfig.from_file(cv_config)
        else:
            config = CvFlowConfig()
        return CvFlowSource.Builder(path, config, size, **common)

    def close(self):
        pass
