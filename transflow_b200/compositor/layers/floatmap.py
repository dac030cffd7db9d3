"""``floatmap`` layer (extension): float displacement-map accumulation with bilinear (or
nearest) sampling and remap -- the reference's WebGL variant (``extra/www/shaders/acc.frag``, ``remap.frag``)
behind the compositor's layer interface.  The Python reference has no such layer; select it with
``classname="floatmap"`` and optional ``:``-separated settings, e.g. ``"floatmap:nearest:decay=0.01:blur=3"``
(``linear`` | ``nearest``, ``scale=``, ``decay=``, ``blur=``: the uniforms of ``acc.frag:8-12``).

State: ``map`` float32 (H, W, 2), ping-pong on the device (the update is a gather from the previous map).
"""
import numpy as np
import torch

from ... import _lib
from ..._lib import check, ptr, stream_ptr
from .layer import _to_device_u8, flow_to_device


class FloatMapLayer:
    KIND = "floatmap"

    def __init__(self, config, height: int, width: int, sources):
        self.config = config
        self.height, self.width = int(height), int(width)
        self.sources = list(sources)
        self.linear, self.scale, self.decay, self.blur_size = True, 1.0, 0.0, 1
        for item in str(config.classname).split(":")[1:]:
            if item in ("linear", "nearest"):
                self.linear = item == "linear"
            elif item.startswith("scale="):
                self.scale = float(item[6:])
            elif item.startswith("decay="):
                self.decay = float(item[6:])
            elif item.startswith("blur="):
                self.blur_size = int(item[5:])
            else:
                raise ValueError(f"Unknown floatmap setting '{item}'")
        if not 1 <= self.blur_size <= 15:
            raise ValueError(f"floatmap blur size must be in [1, 15], got {self.blur_size}")
        self._lib = _lib.load()
        self._maps = [torch.zeros((self.height, self.width, 2), dtype=torch.float32, device="cuda") for _ in range(2)]
        self._cur = 0
        self._pixmap = None

    def set_sources(self, sources):
        self.sources = list(sources)

    # -- per frame ------------------------------------------------------------------------------
    def _update(self, flow, rgb_inout=None, first_layer=False, background=0):
        fl = flow_to_device(flow)
        if tuple(fl.shape) != (self.height, self.width, 2):
            raise ValueError(f"flow must be ({self.height}, {self.width}, 2), got {tuple(fl.shape)}")
        if len(self.sources) != 1:
            raise ValueError(f"a floatmap layer takes exactly one pixmap source, got {len(self.sources)}")
        self._pixmap = _to_device_u8(self.sources[0].next())
        if tuple(self._pixmap.shape[:2]) != (self.height, self.width):
            raise ValueError(f"pixmap must be ({self.height}, {self.width}, C), got {tuple(self._pixmap.shape)}")
        src, dst = self._maps[self._cur], self._maps[self._cur ^ 1]
        check(self._lib.tf_floatmap_accumulate(ptr(src), ptr(fl), ptr(dst), self.height, self.width, self.scale,
                                               self.decay, self.blur_size, int(self.linear), stream_ptr()))
        self._cur ^= 1
        if rgb_inout is not None:
            check(self._lib.tf_floatmap_remap(ptr(dst), ptr(self._pixmap), int(self._pixmap.shape[2]), int(self.linear),
                                              None, ptr(rgb_inout), int(bool(first_layer)), int(background),
                                              self.height, self.width, stream_ptr()))
        self._keepalive = fl

    def update(self, flow):
        self._update(flow)

    def render_device(self) -> torch.Tensor:
        if self._pixmap is None:
            raise RuntimeError("floatmap layer rendered before its first update")
        out = torch.empty((self.height, self.width, 4), dtype=torch.uint8, device="cuda")
        check(self._lib.tf_floatmap_remap(ptr(self._maps[self._cur]), ptr(self._pixmap), int(self._pixmap.shape[2]),
                                          int(self.linear), ptr(out), None, 0, 0, self.height, self.width, stream_ptr()))
        return out

    def render(self) -> np.ndarray:
        return self.render_device().cpu().numpy()

    def check_indices(self):
        pass  # every sample is clamped to the edge

    # -- state ----------------------------------------------------------------------------------
    @property
    def map(self) -> np.ndarray:
        return self._maps[self._cur].cpu().numpy()

    @map.setter
    def map(self, value):
        self._maps[self._cur].copy_(torch.from_numpy(np.ascontiguousarray(value, dtype=np.float32)))

    def __getstate__(self):
        return {"config": self.config, "height": self.height, "width": self.width, "sources": self.sources,
                "map": self.map}

    def __setstate__(self, state):
        self.__init__(state["config"], state["height"], state["width"], state["sources"])
        self.map = state["map"]

    def close(self):
        self._maps = []
