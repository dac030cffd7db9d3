"""Flow filters (``transflow/flow/filters.py``): in-place elementwise edits of the flow with
expressions of the time ``t``.  Next-tier row 8f-1: they run on the device tensor (so the flow
never returns to the host between estimation and accumulation) but as plain tensor
expressions for now, not hand-written kernels."""
import torch

from ..utils import parse_lambda_expression


class FlowFilter:

    def apply(self, flow: torch.Tensor, t: float) -> None:
        raise NotImplementedError()

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


class ScaleFlowFilter(FlowFilter):
    def __init__(self, args):
        self.expr = parse_lambda_expression(args[0])

    def apply(self, flow, t):
        flow *= self.expr(t)


class ThresholdFlowFilter(FlowFilter):
    def __init__(self, args):
        self.expr = parse_lambda_expression(args[0])

    def apply(self, flow, t):
        norm = torch.linalg.vector_norm(flow, dim=2)
        flow[norm <= self.expr(t)] = 0


class ClipFlowFilter(FlowFilter):
    def __init__(self, args):
        self.expr = parse_lambda_expression(args[0])

    def apply(self, flow, t):
        threshold = self.expr(t)
        norm = torch.linalg.vector_norm(flow, dim=2)
        factors = torch.where(norm >= threshold, threshold / norm, torch.ones_like(norm))
        flow *= factors.unsqueeze(2)


class PolarFlowFilter(FlowFilter):
    def __init__(self, args):
        self.expr_radius = parse_lambda_expression(args[0], ("t", "r", "a"))
        self.expr_theta = parse_lambda_expression(args[1], ("t", "r", "a"))

    def apply(self, flow, t):
        radius = torch.linalg.vector_norm(flow, dim=2)
        theta = torch.atan2(flow[:, :, 1], flow[:, :, 0])
        new_radius = self.expr_radius(t, radius, theta)
        new_theta = self.expr_theta(t, radius, theta)
        new_theta = new_theta if isinstance(new_theta, torch.Tensor) else torch.full_like(theta, float(new_theta))
        flow[:, :, 1] = new_radius * torch.sin(new_theta)
        flow[:, :, 0] = new_radius * torch.cos(new_theta)
