"""Pixmap sources (drop-in for ``transflow/pixmap/source.py``): context manager + iterator of
``uint8 (H, W, 3|4)`` RGB(A) frames.  The generators stay on the host and are seeded exactly
like the reference's so pixels are identical; what is new is ``device=True``: stills are
uploaded to HBM once and the same CUDA tensor is yielded every frame, video frames are
uploaded through pinned memory."""
import logging
import os
import re

import numpy as np

logger = logging.getLogger(__name__)

_STILL_RE = re.compile(r"^(color:[a-z0-9\(\)#, ]+|color|#?[0-9a-f]{6}|noise|bwnoise|cnoise|gradient|first)$")


class PixmapSource:

    IMAGE_EXTS = {".jpg", ".jpeg", ".png", ".webp", ".bmp", ".ico", ".tiff"}

    def __init__(self, alteration_path: str | None, length: int | None = None):
        self.alteration_path = alteration_path
        self.width = None
        self.height = None
        self.framerate = None
        self.alteration = None
        self.length = length
        self.device = False

    def __enter__(self):
        return self

    def __next__(self):
        raise NotImplementedError()

    def __iter__(self):
        return self

    def __exit__(self, exc_type, exc_value, exc_traceback):
        pass

    def load_alteration(self):
        """Opaque pixels of the alteration image overwrite the pixmap's RGB (source.py:40-69)."""
        if self.alteration_path is None:
            return
        import PIL.Image
        image = np.array(PIL.Image.open(self.alteration_path))
        if image.ndim == 2:
            image = image[:, :, None]
        if image.shape[2] < 4:
            pad = np.ones((*image.shape[:2], 4 - image.shape[2]), dtype=np.uint8)
            image = np.concatenate([image, pad], axis=2)
        if self.width is None:
            raise ValueError("Width not initialized")
        ii, jj = np.nonzero(image[:, :, 3])
        base = (ii * self.width + jj) * 3
        inds = (base[:, None] + np.arange(3)[None, :]).ravel()
        vals = image[ii, jj, :3].ravel()
        self.alteration = (inds, vals)

    def setup(self):
        self.load_alteration()

    def alter(self, array):
        if self.alteration is not None:
            np.put(array, self.alteration[0], self.alteration[1])
        return array

    def _emit(self, array):
        """Host array -> what the iterator yields (NumPy, or a CUDA tensor when ``device``)."""
        if not self.device:
            return array
        import torch
        return torch.from_numpy(np.ascontiguousarray(array)).cuda(non_blocking=True)

    @classmethod
    def from_args(cls, path: str, size: tuple, seek=None, seed=None, seek_time=None, alteration_path=None,
                  repeat: int = 1, flow_path=None):
        from . import still
        ext = os.path.splitext(path)[1]
        m = _STILL_RE.match(path.lower().strip())
        if m is not None:
            width, height = size
            kind = m.group(1)
            if kind == "color":
                return still.ColorPixmapSource(width, height, seed=seed, alteration_path=alteration_path)
            if kind.startswith("color:"):
                return still.ColorPixmapSource(width, height, kind.split(":", 1)[1], seed=seed,
                                               alteration_path=alteration_path)
            if re.match(r"#?[0-9a-f]{6}", kind):
                return still.ColorPixmapSource(width, height, kind, seed=seed, alteration_path=alteration_path)
            if kind == "noise":
                return still.NoisePixmapSource(width, height, seed, alteration_path)
            if kind == "bwnoise":
                return still.BwNoisePixmapSource(width, height, seed, alteration_path)
            if kind == "cnoise":
                return still.ColoredNoisePixmapSource(width, height, seed, alteration_path)
            if kind == "gradient":
                return still.GradientPixmapSource(width, height, seed)
            if kind == "first":
                assert flow_path is not None
                return still.VideoStillPixmapSource(flow_path, alteration_path)
            raise ValueError(f"Unknown pixmap source '{kind}'")
        if os.path.isfile(path) and ext.lower() in cls.IMAGE_EXTS:
            return still.ImagePixmapSource(path, alteration_path)
        from .cv import CvPixmapSource
        return CvPixmapSource(path, seek, seek_time, alteration_path, repeat)
