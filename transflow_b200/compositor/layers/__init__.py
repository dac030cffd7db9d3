"""Device-backed compositor layers; ``Layer.from_args`` dispatches on ``LayerConfig.classname``."""
from .layer import Layer

__all__ = ["Layer"]
