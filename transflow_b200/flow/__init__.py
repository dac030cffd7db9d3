"""Flow plugin surface: ``FlowSource`` (``from_args`` -> ``Builder``) and its two enums, as ``transflow.flow`` exports
them."""
from .sources.source import FlowSource

__all__ = ["FlowSource", "Direction", "LockMode"]

Direction, LockMode = FlowSource.Direction, FlowSource.LockMode
