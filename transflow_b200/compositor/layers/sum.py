"""The "sum accumulator" (``sum.py:7-14``): data[..., (i, j)] += floor(flow) -> reset -> remap."""
from .layer import Layer


class SumLayer(Layer):
    KIND = "sum"
