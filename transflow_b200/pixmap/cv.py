"""Video pixmap source (``transflow/pixmap/cv.py:11-66``): RGB frames of a video, with seek and
repeat.  Also accepts an already-open capture object (e.g. ``ArrayCapture``)."""
import warnings

import numpy as np

from .source import PixmapSource


class CvPixmapSource(PixmapSource):

    def __init__(self, path, seek=None, seek_time=None, alteration_path=None, repeat: int = 1):
        PixmapSource.__init__(self, alteration_path)
        self.path = path
        self.capture = None
        self.seek = seek
        self.seek_time = seek_time
        self.repeat = repeat
        self.loop_index = 1

    def rewind(self):
        import cv2
        if self.capture is None:
            raise ValueError("Capture not initialized")
        self.capture.set(cv2.CAP_PROP_POS_MSEC, 0)
        for _ in range(self.seek or 0):
            self.capture.read()

    def __enter__(self):
        import cv2
        self.setup()
        self.capture = self.path if hasattr(self.path, "read") else cv2.VideoCapture(self.path)
        self.width = int(self.capture.get(cv2.CAP_PROP_FRAME_WIDTH))
        self.height = int(self.capture.get(cv2.CAP_PROP_FRAME_HEIGHT))
        self.framerate = round(self.capture.get(cv2.CAP_PROP_FPS))
        count = self.capture.get(cv2.CAP_PROP_FRAME_COUNT)
        if self.repeat > 0 and count is not None and int(count) > 0:
            self.length = int(count) * self.repeat
        if self.seek_time is not None:
            self.seek = int(self.seek_time * self.framerate)
            if self.length is not None:
                self.length -= self.seek * self.repeat
        self.rewind()
        return self

    def __next__(self):
        import cv2
        assert self.capture is not None
        if hasattr(self.capture, "isOpened") and not self.capture.isOpened():
            warnings.warn("Attempt to read frame from pixmap capture, which was not opened")
            raise StopIteration
        while True:
            success, frame = self.capture.read()
            if success and frame is not None:
                break
            if self.repeat == 0 or self.loop_index < self.repeat:
                self.loop_index += 1
                self.rewind()
                continue
            raise StopIteration
        if not isinstance(frame, np.ndarray):      # device-resident BGR frame: flip channels there
            return frame.flip(-1).contiguous()
        return self._emit(self.alter(np.array(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))))

    def __exit__(self, exc_type, exc_value, exc_traceback):
        if self.capture is not None:
            self.capture.release()
