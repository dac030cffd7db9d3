"""Flow visualisers on device: ``render1d`` / ``render2d`` of ``transflow/output/render.py:9-48``
(used by the pipeline when ``view_flow`` / ``view_flow_magnitude`` is set, ``pipeline.py:509-516``).
Same signatures and defaults; arrays are CUDA tensors (NumPy arrays are uploaded), the result is a
device ``uint8 (H, W, 3)`` tensor, bit-exact with the reference's float32 NumPy arithmetic."""
import numpy as np
import torch

from .. import ops
from ..utils import parse_color


def _device(arr) -> torch.Tensor:
    if isinstance(arr, torch.Tensor):
        return arr
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).cuda()


def render1d(arr, scale: float = 1, colors=None, binary: bool = False) -> torch.Tensor:
    if colors is None:
        colors = ("#000000", "#ffffff")
    return ops.render_flow(_device(arr), ops.RENDER_1D, scale, [parse_color(c) for c in colors[:2]], binary)


def render2d(arr, scale: float = 1, colors=None) -> torch.Tensor:
    if colors is None:
        colors = ("#ffff00", "#0000ff", "#ff00ff", "#00ff00")
    return ops.render_flow(_device(arr), ops.RENDER_2D, scale, [parse_color(c) for c in colors[:4]])


def render_magnitude(flow, scale: float = 1, colors=None, binary: bool = False) -> torch.Tensor:
    """``render1d(numpy.sqrt(numpy.sum(numpy.power(flow, 2), axis=2)), ...)`` (pipeline.py:515-516) in one kernel."""
    if colors is None:
        colors = ("#000000", "#ffffff")
    return ops.render_flow(_device(flow), ops.RENDER_MAGNITUDE, scale, [parse_color(c) for c in colors[:2]], binary)
