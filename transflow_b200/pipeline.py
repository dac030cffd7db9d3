"""In-process device pipeline (the GPU-mode counterpart of ``transflow/pipeline.py``).

The reference runs one OS process per flow source / pixmap source / output and ships whole
ndarrays through ``multiprocessing.Queue`` (66 MB pickled per 4K flow, ``pipeline.py:326``);
here decode -> flow -> accumulate -> remap stay on one device in one process, and the only
per-frame host traffic is the input frame (H2D) and the output frame (D2H), each on its own
CUDA stream so they overlap the kernels of neighbouring frames.

``Pipeline(config).run()`` keeps the reference's hot loop (``pipeline.py:545-575``):
``flow = next(flow_source)`` -> ``compositor.update(flow)`` -> ``compositor.render()`` ->
outputs, with ``cursor``, ``Status`` messages, cancel event and ``.ckpt.zip`` checkpoints
(``pipeline.py:225-242``: ``meta.json`` + pickled compositor; a run whose flow path is such an archive
resumes from it, ``pipeline.py:290-303``).
"""
import collections
import io
import json
import os
import pickle
import time
import zipfile
from dataclasses import dataclass, field

import numpy as np
import torch

from .compositor import Compositor
from .compositor.pixmap_source_interface import PixmapSourceInterface
from .config import LayerConfig, PixmapSourceConfig
from .flow import FlowSource
from .pixmap.source import PixmapSource
from .utils import load_bool_mask, parse_timestamp


@dataclass
class Config:
    """The subset of ``transflow.config.Config`` (``config.py:160-256``) that reaches the hot path."""
    flow_path: object                      # video path, or an open capture (e.g. ArrayCapture)
    mask_path: str | None = None
    kernel_path: str | None = None
    cv_config: object = None               # JSON path or a CvFlowConfig
    flow_filters: str | None = None
    direction: object = "forward"
    seek_time: object = None
    duration_time: object = None
    repeat: int = 1
    lock_expr: str | None = None
    lock_mode: object = None
    pixmap_sources: list = field(default_factory=list)
    layers: list = field(default_factory=list)
    compositor_background: str | None = None
    output_path: object = None             # None | "dir/%d.png" | callable(frame_index, rgb ndarray)
    size: tuple | None = None
    seed: int | None = None
    extra_flow_paths: list = field(default_factory=list)   # more flow sources, merged per frame (pipeline.py:328-334)
    flows_merging_function: str = "first"                  # key of FLOW_MERGING_FUNCTIONS (pipeline.py:149-158)
    export_flow: object = None                             # None | path of the ".flow.zip" to write (pipeline.py:363-377)
    round_flow: bool = False                               # export rounded integer flows (pipeline.py:505)
    view_flow: bool = False                                # output render2d(flow) instead of the composite (pipeline.py:512)
    view_flow_magnitude: bool = False                      # output render1d(|flow|) (pipeline.py:514-516)
    render_scale: float = 1
    render_colors: object = None                           # "c1,c2,..." | list | tuple | None (config.py:248-252)
    render_binary: bool = False

    def __post_init__(self):
        self.seek_time = parse_timestamp(self.seek_time) or 0
        self.duration_time = parse_timestamp(self.duration_time)
        if self.compositor_background is None:
            self.compositor_background = "#FFFFFF"
        if isinstance(self.render_colors, str):
            self.render_colors = tuple(self.render_colors.split(","))
        elif isinstance(self.render_colors, list):
            self.render_colors = tuple(self.render_colors)
        if self.seed is None:
            self.seed = int.from_bytes(os.urandom(4), "little")
        known = {layer.index for layer in self.layers}
        for pm in self.pixmap_sources:
            for li in pm.layers:
                if li not in known:
                    self.layers.append(LayerConfig(li))
                    known.add(li)

    _PLAIN_FIELDS = ("mask_path", "kernel_path", "flow_filters", "seek_time", "duration_time", "repeat", "lock_expr",
                     "compositor_background", "size", "seed", "extra_flow_paths", "flows_merging_function",
                     "export_flow", "round_flow", "view_flow", "view_flow_magnitude", "render_scale", "render_colors",
                     "render_binary")

    def todict(self) -> dict:
        """Same keys as the reference's ``Config.todict`` (config.py:290-323) for the fields kept here."""
        d = {k: getattr(self, k) for k in self._PLAIN_FIELDS}
        d["flow_path"] = self.flow_path if isinstance(self.flow_path, str) else repr(self.flow_path)
        d["cv_config"] = self.cv_config if isinstance(self.cv_config, (str, type(None))) else self.cv_config.to_dict()
        d["direction"] = FlowSource.Direction.from_arg(self.direction).value
        d["lock_mode"] = FlowSource.LockMode.from_arg(self.lock_mode).value
        d["pixmap_sources"] = [p.todict() for p in self.pixmap_sources]
        d["layers"] = [layer.todict() for layer in self.layers]
        d["output_path"] = self.output_path if isinstance(self.output_path, (str, type(None))) else None
        d["export_flow"] = None if self.export_flow is None else str(self.export_flow)
        d["size"] = None if self.size is None else list(self.size)
        return d

    @classmethod
    def fromdict(cls, d: dict):
        """Inverse of ``todict`` (reference: config.py:258-288); unknown keys of a reference-written dict are ignored."""
        from .flow.sources.cv import CvFlowConfig
        cv = d.get("cv_config")
        if isinstance(cv, dict):
            cv = CvFlowConfig(**cv)
        kwargs = {k: d[k] for k in cls._PLAIN_FIELDS if k in d}
        if kwargs.get("size") is not None:
            kwargs["size"] = tuple(kwargs["size"])
        if kwargs.get("extra_flow_paths") is None:
            kwargs.pop("extra_flow_paths", None)
        return cls(d["flow_path"], cv_config=cv, direction=d.get("direction", "forward"),
                   lock_mode=d.get("lock_mode"), output_path=d.get("output_path"),
                   pixmap_sources=[PixmapSourceConfig.fromdict(x) for x in d.get("pixmap_sources", [])],
                   layers=[LayerConfig.fromdict(x) for x in d.get("layers", [])], **kwargs)

    def get_secondary_output_path(self, suffix: str) -> str:
        """``<output or flow path without extension><suffix>`` (config.py:325-341)."""
        import re
        base = self.output_path if isinstance(self.output_path, str) else None
        if base is None:
            base = self.flow_path if isinstance(self.flow_path, str) else "transflow"
        path = os.path.splitext(base)[0]
        if path.endswith(".flow") or path.endswith(".ckpt"):
            path = path[:-5]
        if re.match(r".*\.(\d{3})$", path):
            path = path[:-4]
        return path + suffix


class _IteratorQueue:
    """Adapts an in-process pixmap iterator to the ``queue.get(timeout)`` the interface calls."""

    def __init__(self, iterator):
        self.iterator = iterator

    def get(self, timeout=None):
        try:
            return next(self.iterator)
        except StopIteration:
            return None


class Pipeline:

    Status = collections.namedtuple("Status", ["cursor", "total", "elapsed", "error"])

    def __init__(self, config: Config, status_queue=None, cancel_event=None, checkpoint_every=None,
                 checkpoint_end=False, checkpoint_path=None, keep_frames_on_device=False):
        self.config = config
        self.status_queue = status_queue
        self.cancel_event = cancel_event
        self.checkpoint_every = checkpoint_every
        self.checkpoint_end = checkpoint_end
        self.checkpoint_path = checkpoint_path
        self.keep_frames_on_device = keep_frames_on_device
        self.compositor: Compositor | None = None
        self.cursor = 0
        self.expected_length = None
        self.flow_source = None
        self._flow_builder = None
        self._pixmaps = []
        self._ckpt_meta = {}
        self.last_frame = None
        self._down_stream = None
        self._ring = None
        self._pending = None

    # -- setup (pipeline.py:290-455) -----------------------------------------------------------------
    def _setup_checkpoint(self):
        """Resume from a ``.ckpt.zip`` given as the flow path (pipeline.py:290-303): the archive's config replaces
        ours, shifted by ``cursor / framerate``; the pickled compositor carries the accumulated state."""
        self._ckpt_meta = {}
        path = self.config.flow_path
        if not (isinstance(path, str) and path.endswith(".ckpt.zip")):
            return
        with zipfile.ZipFile(path) as archive:
            meta = json.loads(archive.read("meta.json").decode())
            self.compositor = pickle.loads(archive.read("compositor.bin"))
        self._ckpt_meta = meta
        self.config = Config.fromdict(meta["config"])
        shift = meta["cursor"] / meta["framerate"]
        self.config.seek_time += shift
        if self.config.duration_time is not None:
            self.config.duration_time -= shift
        self.cursor = int(meta["cursor"])

    def _setup_flow_source(self):
        c = self.config
        self._flow_builder = FlowSource.from_args(
            c.flow_path, mask_path=c.mask_path, kernel_path=c.kernel_path, cv_config=c.cv_config,
            flow_filters=c.flow_filters, size=c.size, direction=c.direction,
            seek_ckpt=self._ckpt_meta.get("cursor"), seek_time=c.seek_time, duration_time=c.duration_time,
            repeat=c.repeat, lock_expr=c.lock_expr, lock_mode=c.lock_mode)
        self.flow_source = self._flow_builder.__enter__()
        self.flow_source.output = "device"
        self.expected_length = self.flow_source.length
        from . import ops
        if c.flows_merging_function not in ops.MERGE_MODES:
            raise ValueError(f"Unknown flows merging function {c.flows_merging_function}")
        self._extra_builders, self.extra_flow_sources = [], []
        for path in c.extra_flow_paths:        # extra sources take no mask / kernel / filters (pipeline.py:331)
            builder = FlowSource.from_args(path, cv_config=c.cv_config, size=c.size, direction=c.direction,
                                           seek_time=c.seek_time, duration_time=c.duration_time, repeat=c.repeat)
            src = builder.__enter__()
            src.output = "device"
            self._extra_builders.append(builder)
            self.extra_flow_sources.append(src)
        self.flow_output = None
        if c.export_flow:
            from .output import NumpyOutput
            fs = self.flow_source
            self.flow_output = NumpyOutput(str(c.export_flow), replace=True)
            self.flow_output.write_meta({"path": c.flow_path if isinstance(c.flow_path, str) else repr(c.flow_path),
                                         "width": fs.width, "height": fs.height, "framerate": fs.framerate,
                                         "direction": fs.direction.value, "seek_time": c.seek_time})

    def _next_flow(self) -> torch.Tensor:
        """``Pipeline._update_flow`` (pipeline.py:492-507): one flow per source, merged, upscaled, exported."""
        from . import ops
        flows = [next(self.flow_source)]
        for src in self.extra_flow_sources:
            flows.append(next(src))
        flow = flows[0]
        if len(flows) > 1 or self.config.flows_merging_function != "first":
            flow = ops.merge_flows(flows, self.config.flows_merging_function)
        flow = self._upscale(flow)
        if self.flow_output is not None:
            self.flow_output.write_array(torch.round(flow).to(torch.int64) if self.config.round_flow else flow)
        return flow

    def _setup_pixmaps_and_compositor(self):
        fw, fh = self.flow_source.width, self.flow_source.height
        size = self.config.size if self.config.size is not None else (fw, fh)
        opened = []
        for pc in self.config.pixmap_sources:
            src = PixmapSource.from_args(pc.path, size, seek=self._ckpt_meta.get("cursor"), seed=self.config.seed,
                                         seek_time=pc.seek_time, alteration_path=pc.alteration_path,
                                         repeat=pc.repeat, flow_path=self.config.flow_path
                                         if isinstance(self.config.flow_path, str) else None)
            src.device = True
            opened.append(src.__enter__())
        self._pixmaps = opened
        width, height = fw, fh
        if opened:
            pw, ph = opened[0].width, opened[0].height
            for s in opened[1:]:
                if (s.width, s.height) != (pw, ph):
                    raise ValueError(f"Pixmap sources must have the same dimensions, found {pw}x{ph} and "
                                     f"{s.width}x{s.height}")
            if (pw, ph) != (fw, fh):
                if pw % fw != 0 or ph % fh != 0:
                    raise ValueError(f"Resolutions do not match: flow is {fw}x{fh} while pixmap is {pw}x{ph}.")
                width, height = pw, ph
        self._scale = (width // fw, height // fh)
        if self.compositor is None:
            self.compositor = Compositor.from_args(height, width, self.config.layers,
                                                   background_color=self.config.compositor_background,
                                                   seed=self.config.seed)
        interfaces = {}
        for pc, src in zip(self.config.pixmap_sources, opened):
            q = _IteratorQueue(src)
            for li in pc.layers:
                mask = load_bool_mask(pc.introduction_path, (height, width), True)
                interfaces.setdefault(li, []).append(PixmapSourceInterface(q, mask))
        self.compositor.set_sources(interfaces)
        # the compositor is the flow's only consumer: forward flows may travel as the scatter pass's claim plane
        c = self.config
        if (not self.extra_flow_sources and c.flows_merging_function == "first" and self._scale == (1, 1)
                and self.flow_output is None and not c.view_flow and not c.view_flow_magnitude
                and os.environ.get("TFB200_FORWARD_CLAIMS", "1") != "0"):
            self.flow_source.output = "claims"

    # -- per-frame pieces ----------------------------------------------------------------------------
    def _upscale(self, flow: torch.Tensor) -> torch.Tensor:
        wf, hf = self._scale
        if wf == 1 and hf == 1:
            return flow
        # utils.upscale_array (utils.py:417-418): vector scaling then block replicate
        from . import ops
        return ops.upscale_flow(flow, wf, hf)

    def _output_frame(self, flow: torch.Tensor):
        """``compositor.update(flow)`` then ``_update_output`` (pipeline.py:562-564, 509-521): the flow visualisers
        replace the composite when asked for."""
        c = self.config
        if c.view_flow or c.view_flow_magnitude:
            from .output import render
            self.compositor.update(flow)
            if c.view_flow:
                return render.render2d(flow, c.render_scale, c.render_colors)
            return render.render_magnitude(flow, c.render_scale, c.render_colors, c.render_binary)
        return self.compositor.step(flow)

    def _deliver(self, index: int, host: np.ndarray):
        out = self.config.output_path
        if callable(out):
            out(index, host)
        elif isinstance(out, str):
            import PIL.Image
            path = out % index if "%" in out else out
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            PIL.Image.fromarray(host).save(path)

    def _emit(self, index: int, frame: torch.Tensor):
        """Asynchronous D2H: frame t is copied on the download stream into a pinned ring while the
        kernels of frame t+1 run; the sink receives frame t-1 (the ring is flushed at the end)."""
        self.last_frame = frame
        if self.config.output_path is None or self.keep_frames_on_device:
            return
        if self._down_stream is None:
            self._down_stream = torch.cuda.Stream()
            self._ring = [torch.empty(tuple(frame.shape), dtype=torch.uint8).pin_memory() for _ in range(2)]
            self._pending = [None, None]
        k = index & 1
        self._flush_slot(k)
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(self._down_stream):
            self._down_stream.wait_event(done)
            self._ring[k].copy_(frame, non_blocking=True)
            frame.record_stream(self._down_stream)
            copied = torch.cuda.Event()
            copied.record(self._down_stream)
        self._pending[k] = (index, copied)

    def _flush_slot(self, k: int):
        if self._pending is None or self._pending[k] is None:
            return
        index, copied = self._pending[k]
        copied.synchronize()
        self._pending[k] = None
        self._deliver(index, self._ring[k].numpy())

    def _flush_outputs(self):
        if self._pending is None:
            return
        order = sorted((p[0], k) for k, p in enumerate(self._pending) if p is not None)
        for _, k in order:
            self._flush_slot(k)

    def export_checkpoint(self, path: str | None = None):
        """``meta.json`` + ``compositor.bin`` (pickled compositor with sources detached)."""
        assert self.compositor is not None
        saved = [layer.sources for layer in self.compositor.layers]
        for layer in self.compositor.layers:
            layer.sources = []
        try:
            blob = pickle.dumps(self.compositor)
        finally:
            for layer, src in zip(self.compositor.layers, saved):
                layer.sources = src
        meta = {"config": self.config.todict(), "cursor": self.cursor,
                "framerate": self.flow_source.framerate if self.flow_source else None, "timestamp": time.time()}
        if path is None:
            path = self.checkpoint_path or self.config.get_secondary_output_path(f"_{self.cursor:05d}.ckpt.zip")
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        self.last_checkpoint_path = path
        with zipfile.ZipFile(path, "w") as archive:
            archive.writestr("meta.json", json.dumps(meta))
            archive.writestr("compositor.bin", blob)

    # -- run -----------------------------------------------------------------------------------------
    def run(self):
        start = time.time()
        self._setup_checkpoint()
        self._setup_flow_source()
        self._setup_pixmaps_and_compositor()
        error = None
        try:
            while True:
                if self.cancel_event is not None and self.cancel_event.is_set():
                    break
                try:
                    flow = self._next_flow()
                except StopIteration:
                    break
                frame = self._output_frame(flow)
                if frame is not None:
                    self._emit(self.cursor, frame)
                self.cursor += 1
                if self.checkpoint_every is not None and self.cursor % self.checkpoint_every == 0:
                    self.export_checkpoint()
                if self.status_queue is not None:
                    self.status_queue.put(Pipeline.Status(self.cursor, self.expected_length, time.time() - start, None))
            self._flush_outputs()
            for layer in self.compositor.layers:
                layer.check_indices()
        except Exception as err:  # a CUDA error surfaces here as a Python exception
            error = err
            if self.status_queue is not None:
                self.status_queue.put(Pipeline.Status(self.cursor, self.expected_length, time.time() - start, str(err)))
            raise
        finally:
            if self.checkpoint_end and error is None:
                self.export_checkpoint()
            self.close()
        return self.cursor

    def close(self):
        if self._flow_builder is not None:
            self._flow_builder.__exit__(None, None, None)
            self._flow_builder = None
        for builder in getattr(self, "_extra_builders", []):
            builder.__exit__(None, None, None)
        self._extra_builders = []
        if getattr(self, "flow_output", None) is not None:
            self.flow_output.close()
            self.flow_output = None
        for src in self._pixmaps:
            src.__exit__(None, None, None)
        self._pixmaps = []
