"""Static layer (``static.py:7-17``): rgba[mask] = pixmap[mask], alpha preset to 1."""
from .layer import Layer


class StaticLayer(Layer):
    KIND = "static"

    @property
    def data(self):
        raise AttributeError("StaticLayer has no data array")
