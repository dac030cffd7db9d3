"""What a compositor layer pulls pixmaps from (``transflow/compositor/pixmap_source_interface.py``).

Same three members as the reference (``next`` / ``get`` / ``frame_number`` plus
``introduction_mask``).  The queue may deliver NumPy ``uint8 (H, W, 3|4)`` arrays (the
reference's transport) or CUDA tensors of the same layout (in-process device pipeline: stills
are uploaded once and re-yielded)."""
import numpy as np


class EndOfPixmap(StopIteration):
    pass


class PixmapSourceInterface:

    def __init__(self, queue, introduction_mask):
        self.queue = queue
        self.image = None
        self.counter = -1
        self.introduction_mask = np.asarray(introduction_mask, dtype=bool)

    def get(self):
        assert self.image is not None
        return self.image

    def next(self, timeout: float = 1):
        image = self.queue.get(timeout=timeout)
        if image is None:
            raise EndOfPixmap
        if not (hasattr(image, "shape") and len(image.shape) == 3 and image.shape[2] in (3, 4)):
            raise AssertionError("pixmap must be (H, W, 3|4)")
        self.image = image
        self.counter += 1
        return image

    @property
    def frame_number(self) -> int:
        return self.counter


class StillQueue:
    """In-process stand-in for the pixmap SourceProcess queue: yields frames[min(i, last)]."""

    def __init__(self, frames):
        self.frames = list(frames) if isinstance(frames, (list, tuple)) else [frames]
        self.i = 0

    def get(self, timeout=None):
        f = self.frames[min(self.i, len(self.frames) - 1)]
        self.i += 1
        return f
