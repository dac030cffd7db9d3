from .compositor import Compositor
