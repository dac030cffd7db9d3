"""``.flow.zip`` archives as a flow source (drop-in for ``transflow/flow/sources/archive.py:9-48``).

Format (written by ``output/numpy.py`` here and in the reference, ``pipeline.py:363-377``): a zip holding
``meta.json`` (``width``, ``height``, ``framerate``, ``direction`` -- archives older than the direction field are
forward) and one ``NNNNNNNNN.npy`` (H, W, 2) array per frame.  Frames are read on the host and uploaded; the
post-process (filters, mask, kernel, direction conversion) then runs on the device like for any other source.
"""
import json
import zipfile

import numpy as np
import torch

from .source import FlowSource


class ArchiveFlowSource(FlowSource):

    class Builder(FlowSource.Builder):

        def __init__(self, path: str, **kwargs):
            super().__init__(**kwargs)
            self.path = path
            self.archive = None

        @property
        def cls(self):
            return ArchiveFlowSource

        def build(self):
            self.archive = zipfile.ZipFile(self.path)
            with self.archive.open("meta.json") as file:
                data = json.loads(file.read().decode())
            self.direction = FlowSource.Direction(data.get("direction", FlowSource.Direction.FORWARD.value))
            self.width = data["width"]
            self.height = data["height"]
            self.framerate = data["framerate"]
            self.base_length = len(self.archive.infolist()) - 1
            super().build()

        def args(self):
            return [self.archive, *FlowSource.Builder.args(self)]

    def __init__(self, archive: zipfile.ZipFile, *args, **kwargs):
        self.archive = archive
        FlowSource.__init__(self, *args, **kwargs)

    def validate(self):
        super().validate()
        self.assert_type("archive", zipfile.ZipFile)

    def next(self):
        with self.archive.open(f"{self.input_frame_index:09d}.npy") as file:
            flow = np.load(file)
        # archives written with --round-flow hold integers; the flow type on the device is float32
        return torch.from_numpy(np.ascontiguousarray(flow, dtype=np.float32)).cuda(non_blocking=True)

    def close(self):
        self.archive.close()
