"""Still pixmap sources (``transflow/pixmap/still.py``): colour, three kinds of noise, a random
expression-tree gradient, an image file, the first frame of the flow video.  Seeded with
``numpy.random.seed`` / ``random.seed`` exactly like the reference (``still.py:45,60,69,78,172``)."""
import random

import numpy as np

from .source import PixmapSource
from ..utils import parse_color


class StillPixmapSource(PixmapSource):

    def __init__(self, width=None, height=None, seed=None, alteration_path=None):
        PixmapSource.__init__(self, alteration_path, length=None)
        self.width = width
        self.height = height
        self.seed = seed
        self.array = None
        self._resident = None

    def _init_array(self):
        raise NotImplementedError()

    def _need_size(self):
        if self.width is None or self.height is None:
            raise ValueError("Width or height not initialized")
        return self.height, self.width

    def __enter__(self):
        self.array = self._init_array()
        self.height, self.width = self.array.shape[:2]
        self.setup()
        return self

    def __next__(self):
        assert self.array is not None
        if self.device:
            if self._resident is None:          # upload once; a still never changes
                self._resident = self._emit(self.alter(self.array.copy()))
            return self._resident
        return self.alter(self.array.copy())


class ColorPixmapSource(StillPixmapSource):

    def __init__(self, width, height, color=None, seed=None, alteration_path=None):
        StillPixmapSource.__init__(self, width, height, seed, alteration_path)
        self.color = color

    def _init_array(self):
        np.random.seed(self.seed)
        if self.color is None:
            color = list(np.random.randint(0, 256, size=(3), dtype=np.uint8))
        else:
            color = parse_color(self.color)
        h, w = self._need_size()
        out = np.zeros((h, w, 3), dtype=np.uint8)
        out[:, :, :] = color
        return out


class NoisePixmapSource(StillPixmapSource):
    def _init_array(self):
        np.random.seed(self.seed)
        h, w = self._need_size()
        return np.repeat(np.random.randint(0, 256, size=(h, w, 1), dtype=np.uint8), 3, axis=2)


class BwNoisePixmapSource(StillPixmapSource):
    def _init_array(self):
        np.random.seed(self.seed)
        h, w = self._need_size()
        return np.repeat(np.random.choice([0, 255], size=(h, w, 1)), 3, axis=2).astype(np.uint8)


class ColoredNoisePixmapSource(StillPixmapSource):
    def _init_array(self):
        np.random.seed(self.seed)
        h, w = self._need_size()
        return np.random.randint(0, 256, size=(h, w, 3), dtype=np.uint8)


class GradientPixmapSource(StillPixmapSource):
    """Random expression tree over (row, column) in [-1, 1] (``still.py:84-184``).  The tree is
    drawn with the same ``random`` calls as the reference; evaluation is vectorised over the
    pixel grid in float64, which is the arithmetic the per-pixel Python loop performs."""

    NODE_I, NODE_J, NODE_RGB, NODE_MIX, NODE_TRIPLE, NODE_Z, NODE_B = range(7)

    # The tree grammar of the reference (still.py:94-118), written as three mutually recursive productions.  A
    # seeded run must issue the same `random.random()` calls in the same order to give the same picture:
    #   inner(kind, d)  = leaf                      if d <= 0      (kind: MIX or TRIPLE)
    #                   = (kind, operand(d-1) x 3)  otherwise, operands left to right
    #   operand(d)      = leaf                      if d <= 0      (no draw)
    #                   = leaf if draw < 1/4 else inner(MIX, d-1)
    #   leaf            = I if draw < .333, J if draw < .666, else RGB with three more draws in [-1, 1)
    def _leaf(self) -> tuple:
        pick = random.random()
        if pick < .333:
            return (self.NODE_I, None, None, None)
        if pick < .666:
            return (self.NODE_J, None, None, None)
        channels = [random.random() * 2 - 1 for _ in range(3)]
        return (self.NODE_RGB, *channels)

    def _operand(self, depth: int) -> tuple:
        if depth <= 0 or random.random() < .25:
            return self._leaf()
        return self._inner(self.NODE_MIX, depth - 1)

    def _inner(self, kind: int, depth: int) -> tuple:
        if depth <= 0:
            return self._leaf()
        return (kind, *[self._operand(depth - 1) for _ in range(3)])

    def generate(self, node_type: int, depth: int) -> tuple:
        if node_type in (self.NODE_TRIPLE, self.NODE_MIX):
            return self._inner(node_type, depth)
        if node_type == self.NODE_B:
            return self._operand(depth)
        if node_type == self.NODE_Z:
            return self._leaf()
        raise ValueError(f"Unknown node type {node_type}")

    def evaluate(self, tree: tuple, i, j):
        """-> three float64 arrays (or scalars) for channel r, g, b."""
        kind, a, b, c = tree
        if kind == self.NODE_TRIPLE:
            return (self.evaluate(a, i, j)[0], self.evaluate(b, i, j)[1], self.evaluate(c, i, j)[2])
        if kind == self.NODE_MIX:
            ea, eb, ec = self.evaluate(a, i, j), self.evaluate(b, i, j), self.evaluate(c, i, j)
            out = []
            for k in range(3):
                w = (1 + ea[k]) / 2
                out.append((1 - w) * eb[k] + w * ec[k])
            return tuple(out)
        if kind == self.NODE_RGB:
            return (a, b, c)
        h, w = self._need_size()
        if kind == self.NODE_I:
            z = 2 * (i / (h - 1)) - 1
            return (z, z, z)
        if kind == self.NODE_J:
            z = 2 * (j / (w - 1)) - 1
            return (z, z, z)
        raise NotImplementedError(f"Unknown node type {kind}")

    def _init_array(self):
        random.seed(self.seed)
        tree = self.generate(self.NODE_TRIPLE, 5)
        h, w = self._need_size()
        ii = np.arange(h, dtype=np.float64)[:, None] * np.ones((1, w))
        jj = np.ones((h, 1)) * np.arange(w, dtype=np.float64)[None, :]
        out = np.zeros((h, w, 3), dtype=np.uint8)
        for k, chan in enumerate(self.evaluate(tree, ii, jj)):
            out[:, :, k] = (255 * (np.asarray(chan, np.float64) * np.ones((h, w)) + 1) / 2).astype(np.uint8)
        return out


class ImagePixmapSource(StillPixmapSource):

    def __init__(self, path: str, alteration_path=None):
        StillPixmapSource.__init__(self, alteration_path=alteration_path)
        self.path = path

    def _init_array(self):
        import PIL.Image
        with PIL.Image.open(self.path) as image:
            array = np.array(image)[:, :, :]
        assert array.shape[2] in (3, 4), f"Pixmap image has unsupported dimension: {array.shape}"
        return array


class VideoStillPixmapSource(ImagePixmapSource):

    def _init_array(self):
        import cv2
        capture = cv2.VideoCapture(self.path)
        success, frame = capture.read()
        assert success, "Could not open video for still bitmap source"
        capture.release()
        return np.array(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
