from .layer import Layer
