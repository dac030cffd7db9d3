"""Pixel-carrying "stack" layer (``introduction.py:8-67``): 8-int records
(r, g, b, alpha, source, i, j, frame) moved by the flow, then refreshed from the pixmaps."""
import ctypes as C

from .layer import Layer
from ..._lib import check


class IntroductionLayer(Layer):
    KIND = "introduction"
    DEPTH = 8
    INDEX_I = 5
    INDEX_J = 6
    INDEX_ALPHA = 3
    INDEX_SOURCE = 4

    @property
    def introduced_once(self) -> bool:
        once = C.c_int()
        check(self._lib.tf_layer_get_counters(self._handle, None, C.byref(once)))
        return bool(once.value)

    def _needs_pixmaps(self) -> bool:
        # the reference returns before source.next() once introduce_once has fired (introduction.py:21-22)
        return not (self.config.introduce_once and self.introduced_once)
