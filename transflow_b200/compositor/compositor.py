"""Device compositor (drop-in for ``transflow/compositor/compositor.py``).

``update(flow)`` / ``render()`` keep the reference's call pattern (``pipeline.py:565``, ``:518``);
``step(flow)`` is the same pair fused: every layer's update kernel also applies
``Layer.render`` and overwrites the output frame where it is opaque
(``compositor.py:31-40``), so a single-layer compositor is one kernel per frame.
"""
import logging
from collections.abc import Sequence

import numpy as np
import torch

from .. import _lib
from .._lib import check, ptr, stream_ptr
from ..config import LayerConfig
from ..ops import ForwardClaims
from ..utils import parse_color
from .layers import Layer
from .pixmap_source_interface import PixmapSourceInterface

logger = logging.getLogger(__name__)


class Compositor:

    def __init__(self, height: int, width: int, layers: Sequence[Layer], background_color: str = "#ffffff"):
        self.height = height
        self.width = width
        self.background_color = parse_color(background_color)
        self.layers = layers

    @property
    def background(self) -> np.ndarray:
        bg = np.zeros((self.height, self.width, 3), dtype=np.uint8)
        bg[:, :] = self.background_color
        return bg

    @property
    def _bg_word(self) -> int:
        r, g, b = self.background_color
        return (int(r) << 16) | (int(g) << 8) | int(b)

    def update(self, flow):
        if isinstance(flow, ForwardClaims):
            flow = flow.tensor()
        for layer in self.layers:
            layer.update(flow)

    def render_device(self, out: torch.Tensor | None = None) -> torch.Tensor:
        """RGB uint8 (H, W, 3) on the device."""
        if out is None:
            out = torch.empty((self.height, self.width, 3), dtype=torch.uint8, device="cuda")
        images = [layer.render_device() for layer in self.layers]
        import ctypes as C
        arr = (C.c_void_p * max(len(images), 1))(*[im.data_ptr() for im in images])
        check(_lib.load().tf_composite(arr, len(images), self._bg_word, ptr(out), self.height, self.width,
                                       stream_ptr()))
        return out

    def render(self) -> np.ndarray:
        """RGB array of shape (height, width, 3), on the host like the reference."""
        return self.render_device().cpu().numpy()

    def step(self, flow, out: torch.Tensor | None = None) -> torch.Tensor:
        """update(flow) + render() in one pass over HBM per layer; returns the device frame."""
        if out is None:
            out = torch.empty((self.height, self.width, 3), dtype=torch.uint8, device="cuda")
        if not self.layers:
            out[:, :] = torch.tensor(self.background_color, dtype=torch.uint8, device="cuda")
            return out
        if isinstance(flow, ForwardClaims):
            # a single move-reference layer reads the claim plane itself (no flow in HBM at all)
            if len(self.layers) == 1 and flow.live and self.layers[0].takes_claims():
                self.layers[0]._update_claims(flow, out, background=self._bg_word)
                return out
            flow = flow.tensor()
        for i, layer in enumerate(self.layers):
            layer._update(flow, rgb_inout=out, first_layer=(i == 0), background=self._bg_word)
        return out

    @classmethod
    def from_args(cls, height: int, width: int, layer_configs: list[LayerConfig],
                  background_color: str = "#ffffff", seed: int | None = None):
        """``seed`` (``Config.seed``) keys the device-side random-reset draws per layer; the reference seeds the
        global NumPy generator instead (pipeline.py), which the ``reset_rng = "numpy"`` parity mode still follows."""
        layers = [Layer.from_args(config, height, width, []) for config in layer_configs]
        for layer in layers:
            if hasattr(layer, "set_seed"):
                layer.set_seed(seed)
        return cls(height, width, layers, background_color=background_color)

    def set_sources(self, pixmap_interfaces: dict[int, list[PixmapSourceInterface]]):
        for i, layer in enumerate(self.layers):
            layer.set_sources(pixmap_interfaces.get(i, []))
