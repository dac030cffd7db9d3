"""The "map accumulator": move (``movement.py:20-60``) -> reset -> integer remap
(``reference.py:58-105``), one fused gather kernel on device (csrc/compositor.cu)."""
from .layer import Layer


class MoveReferenceLayer(Layer):
    KIND = "moveref"
