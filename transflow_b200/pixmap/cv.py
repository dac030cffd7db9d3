"""Video pixmap source: the RGB frames of a video (or of an already-open capture such as ``ArrayCapture``), with
an initial seek and ``repeat`` passes (0 = loop forever).  Interface of ``transflow/pixmap/cv.py:11-66``:
``CvPixmapSource(path, seek, seek_time, alteration_path, repeat)``, context manager, iterator of ``uint8
(H, W, 3)``; ``width``, ``height``, ``framerate``, ``length`` are known after ``__enter__``.

The passes over the video are written as one generator (``_passes``) instead of loop bookkeeping inside
``__next__``; frames that already live on the device (a capture yielding CUDA tensors) are channel-flipped there
and never touch the host.
"""
import warnings

import numpy as np

from .source import PixmapSource


def _is_capture(obj) -> bool:
    return hasattr(obj, "read") and hasattr(obj, "get")


class CvPixmapSource(PixmapSource):

    def __init__(self, path, seek=None, seek_time=None, alteration_path=None, repeat: int = 1):
        super().__init__(alteration_path)
        self.path = path
        self.seek = seek
        self.seek_time = seek_time
        self.repeat = repeat
        self.capture = None
        self.loop_index = 1      # 1-based index of the pass being read
        self._frames = None

    # -- positioning ------------------------------------------------------------------------
    def rewind(self):
        """Back to the first frame of a pass: start of the video, then ``seek`` frames skipped."""
        import cv2
        if self.capture is None:
            raise ValueError("Capture not initialized")
        self.capture.set(cv2.CAP_PROP_POS_MSEC, 0)
        skipped = 0
        while skipped < (self.seek or 0):
            self.capture.read()
            skipped += 1

    def _passes(self):
        """Raw BGR frames of every pass, in order."""
        while True:
            self.rewind()
            while True:
                ok, frame = self.capture.read()
                if not ok or frame is None:
                    break
                yield frame
            if self.repeat != 0 and self.loop_index >= self.repeat:
                return
            self.loop_index += 1

    # -- context manager --------------------------------------------------------------------
    def __enter__(self):
        import cv2
        self.setup()
        self.capture = self.path if _is_capture(self.path) else cv2.VideoCapture(self.path)
        props = {name: self.capture.get(getattr(cv2, "CAP_PROP_" + name))
                 for name in ("FRAME_WIDTH", "FRAME_HEIGHT", "FPS", "FRAME_COUNT")}
        self.width, self.height = int(props["FRAME_WIDTH"]), int(props["FRAME_HEIGHT"])
        self.framerate = round(props["FPS"])
        per_pass = int(props["FRAME_COUNT"]) if props["FRAME_COUNT"] is not None else 0
        if self.seek_time is not None:
            self.seek = int(self.seek_time * self.framerate)
        if self.repeat > 0 and per_pass > 0:
            lost = self.seek if self.seek_time is not None else 0   # only a time seek shortens the announced length
            self.length = (per_pass - lost) * self.repeat
        self._frames = self._passes()
        return self

    def __exit__(self, exc_type, exc_value, exc_traceback):
        if self.capture is not None:
            self.capture.release()

    # -- iteration --------------------------------------------------------------------------
    def __next__(self):
        assert self.capture is not None and self._frames is not None
        opened = getattr(self.capture, "isOpened", None)
        if opened is not None and not opened():
            warnings.warn("Attempt to read frame from pixmap capture, which was not opened")
            raise StopIteration
        bgr = next(self._frames)            # StopIteration after the last pass
        if not isinstance(bgr, np.ndarray):
            return bgr.flip(-1).contiguous()    # device-resident frame: BGR -> RGB on the device
        rgb = np.ascontiguousarray(bgr[:, :, ::-1])     # cv2.cvtColor(frame, COLOR_BGR2RGB)
        return self._emit(self.alter(rgb))
