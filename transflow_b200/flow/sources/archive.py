"""Replay of exported flows: the ``.flow.zip`` container as a flow source.

Container (what ``transflow/output/numpy.py`` and ``pipeline.py:363-377`` of the reference write, and
``transflow_b200.output`` here): ``meta.json`` = {width, height, framerate, direction (0 forward / 1 backward;
missing in old archives = forward), ...} plus one ``%09d.npy`` array (H, W, 2) per frame.  Plugin surface of
``transflow/flow/sources/archive.py:9-48`` (``ArchiveFlowSource.Builder(path, **kwargs)``, ``next()``,
``close()``).

The frames are decoded on the host (zip + npy are host formats) into a PINNED staging buffer and uploaded on
the caller's stream; filters, mask, kernel and the direction conversion then run on the device like for any
other source.  Unlike the reference's builder, the frame range is resolved here, so the source ends after the
last archived frame.
"""
import io
import json
import zipfile

import numpy as np
import torch

from .source import FlowSource

META_NAME = "meta.json"


class FlowArchive:
    """Random-access reader of a ``.flow.zip``: ``meta`` dict, ``len()``, ``read(i) -> ndarray``."""

    def __init__(self, path: str):
        self.path = path
        self.zip = zipfile.ZipFile(path)
        self.meta = json.loads(self.zip.read(META_NAME).decode())
        members = set(self.zip.namelist())
        count = 0
        while self.member(count) in members:
            count += 1
        self.count = count

    @staticmethod
    def member(index: int) -> str:
        return f"{index:09d}.npy"

    def __len__(self) -> int:
        return self.count

    def read(self, index: int) -> np.ndarray:
        return np.load(io.BytesIO(self.zip.read(self.member(index))), allow_pickle=False)

    def close(self):
        self.zip.close()


class ArchiveFlowSource(FlowSource):

    class Builder(FlowSource.Builder):

        def __init__(self, path: str, **kwargs):
            super().__init__(**kwargs)
            self.path = path
            self.archive = None

        @property
        def cls(self):
            return ArchiveFlowSource

        def args(self):
            return [self.archive] + FlowSource.Builder.args(self)

        def build(self):
            self.archive = FlowArchive(self.path)
            meta = self.archive.meta
            self.direction = FlowSource.Direction(meta.get("direction", FlowSource.Direction.FORWARD.value))
            self.width, self.height, self.framerate = meta["width"], meta["height"], meta["framerate"]
            self.base_length = len(self.archive)
            super().build()

    def __init__(self, archive: FlowArchive, *args, **kwargs):
        self.archive = archive
        self._staging = None   # pinned (H, W, 2) float32, reused for every frame
        self._uploaded = None  # event: the previous upload has left the staging buffer
        FlowSource.__init__(self, *args, **kwargs)

    def validate(self):
        super().validate()
        self.assert_type("archive", FlowArchive)

    def next(self):
        frame = self.archive.read(self.input_frame_index)
        if tuple(frame.shape) != (self.height, self.width, 2):
            raise ValueError(f"archived flow {self.input_frame_index} has shape {tuple(frame.shape)}, "
                             f"expected ({self.height}, {self.width}, 2)")
        if self._staging is None:
            self._staging = torch.empty((self.height, self.width, 2), dtype=torch.float32).pin_memory()
        if self._uploaded is not None:
            self._uploaded.synchronize()
        # archives exported with round_flow hold integers; the device flow type is float32
        self._staging.numpy()[...] = frame
        flow = self._staging.cuda(non_blocking=True)
        self._uploaded = torch.cuda.Event()
        self._uploaded.record()
        return flow

    def close(self):
        self.archive.close()
