"""Flow filters (``transflow/flow/filters.py``): in-place edits of the flow with expressions of the time ``t``.

``scale``, ``threshold`` and ``clip`` are one scalar per frame and run inside the post-process kernel
(``tf_flow_postprocess_ex``: filters -> mask -> clip / scatter in ONE pass over the flow), with
NumPy's promotion rules reproduced (a Python scalar keeps float32 arithmetic, a NumPy float64 scalar promotes).
``polar`` evaluates arbitrary user expressions of per-pixel arrays ``(r, a)``: those run as device tensor
expressions between two kernel calls.
"""
import os

import torch

from .. import ops
from ..utils import parse_lambda_expression


class FlowFilter:
    #: name of the fused elementwise op, None when the filter needs tensor expressions
    kind: str | None = None

    def op(self, t: float):
        """-> ("scale" | "threshold" | "clip", scalar for this frame), consumed by ``ops.PostProcess``."""
        raise NotImplementedError()

    def apply(self, flow: torch.Tensor, t: float) -> None:
        """Stand-alone, in place on the device flow (same arithmetic as the fused path)."""
        h, w = flow.shape[:2]
        arr, n = ops._pack_flow_ops([self.op(t)])
        lib = ops._lib.load()
        ops.check(lib.tf_flow_filters(ops.ptr(flow), arr, n, None, ops.ptr(flow), h, w, ops.stream_ptr()))

    @classmethod
    def from_args(cls, filter_name: str, filter_args: tuple):
        table = {"scale": (ScaleFlowFilter, 1), "threshold": (ThresholdFlowFilter, 1),
                 "clip": (ClipFlowFilter, 1), "polar": (PolarFlowFilter, 2)}
        if filter_name not in table:
            raise ValueError(f"Unknown filter name '{filter_name}'")
        klass, arity = table[filter_name]
        if len(filter_args) != arity:
            raise ValueError(f"Invalid number of arguments: {filter_name} {filter_args}")
        return klass(filter_args)


class _ScalarFlowFilter(FlowFilter):
    def __init__(self, args):
        self.expr = parse_lambda_expression(args[0])

    def op(self, t):
        return self.kind, self.expr(t)


class ScaleFlowFilter(_ScalarFlowFilter):
    """``flow *= expr(t)`` (filters.py:36-42)."""
    kind = "scale"


class ThresholdFlowFilter(_ScalarFlowFilter):
    """``flow[norm <= expr(t)] = 0`` (filters.py:45-54)."""
    kind = "threshold"


class ClipFlowFilter(_ScalarFlowFilter):
    """``flow *= expr(t) / norm`` where ``norm >= expr(t)`` (filters.py:57-70)."""
    kind = "clip"


class _DeviceNumpy:
    """What ``numpy`` means inside a ``polar`` expression: the reference hands the expressions NumPy arrays
    ``(r, a)`` (filters.py:79-84) and USAGE.md writes them with ``numpy.sin(a)`` etc.  Here ``r`` and ``a`` are CUDA
    tensors, so a NumPy function called with a tensor argument runs as the torch function of the same name on the
    device; anything else is plain NumPy."""

    _ALIASES = {"power": "pow", "absolute": "abs", "arctan2": "atan2", "arcsin": "asin", "arccos": "acos",
                "arctan": "atan", "mod": "remainder", "concatenate": "cat"}

    def __getattr__(self, name):
        import numpy
        target = getattr(numpy, name)
        if not callable(target) or isinstance(target, type):
            return target
        on_device = getattr(torch, self._ALIASES.get(name, name), None)

        def call(*args, **kwargs):
            ref = next((a for a in args if isinstance(a, torch.Tensor)), None)
            if ref is None:
                return target(*args, **kwargs)
            if on_device is None:
                raise NotImplementedError(f"numpy.{name} has no device counterpart")
            args = [a if isinstance(a, torch.Tensor) else torch.as_tensor(a, dtype=ref.dtype, device=ref.device)
                    for a in args]
            return on_device(*args, **kwargs)
        return call


class PolarFlowFilter(FlowFilter):
    """Polar re-parametrisation with user expressions of ``(t, r, a)`` arrays (filters.py:73-87).  The expressions
    are arbitrary Python: they run on the device tensors when they are made of operators and ``numpy`` functions
    torch also has.  An expression that needs anything else raises (there is no silent host path); setting
    ``PolarFlowFilter.allow_host_expressions = True`` (or ``TRANSFLOW_B200_POLAR_HOST_EXPR=1``) lets exactly the user's
    lambda -- nothing of this package's arithmetic -- be evaluated on host copies of ``r`` and ``a``, as the reference
    evaluates it."""

    allow_host_expressions = os.environ.get("TRANSFLOW_B200_POLAR_HOST_EXPR", "0") == "1"

    def __init__(self, args):
        self.exprs = [parse_lambda_expression(a, ("t", "r", "a"), numpy_module=_DeviceNumpy()) for a in args]
        self.host_exprs = [parse_lambda_expression(a, ("t", "r", "a")) for a in args]
        self.expr_radius, self.expr_theta = self.exprs

    def apply(self, flow, t):
        radius = torch.linalg.vector_norm(flow, dim=2)
        theta = torch.atan2(flow[:, :, 1], flow[:, :, 0])
        try:
            new_radius = self.expr_radius(t, radius, theta)
            new_theta = self.expr_theta(t, radius, theta)
        except (NotImplementedError, TypeError, RuntimeError, ValueError, AttributeError) as err:
            if not self.allow_host_expressions:
                raise NotImplementedError(
                    f"polar filter expression cannot run on device tensors ({err}); use operators and numpy functions "
                    "torch also provides, or set PolarFlowFilter.allow_host_expressions = True") from err
            r, a = radius.cpu().numpy(), theta.cpu().numpy()
            new_radius, new_theta = (e(t, r, a) for e in self.host_exprs)

        def tensor(v):
            if isinstance(v, torch.Tensor):
                return v.to(device=flow.device, dtype=torch.float32)
            import numpy
            return torch.as_tensor(numpy.broadcast_to(numpy.asarray(v, dtype=numpy.float32), tuple(theta.shape)).copy(),
                                   device=flow.device)
        new_radius, new_theta = tensor(new_radius), tensor(new_theta)
        flow[:, :, 1] = new_radius * torch.sin(new_theta)
        flow[:, :, 0] = new_radius * torch.cos(new_theta)
