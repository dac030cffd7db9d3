"""Horn-Schunck on device, same signature as ``flow/methods/horn_schunck.py:9-16``."""
from ... import ops


def calc_optical_flow_horn_schunck(prev_grey, next_grey, flow=None, alpha=1, max_iters=3, decay=0, delta=1,
                                   _cache={}):
    h, w = prev_grey.shape
    if (h, w) not in _cache:
        _cache.clear()
        _cache[(h, w)] = ops.HornSchunck(h, w)
    return _cache[(h, w)](prev_grey, next_grey, flow, alpha, max_iters, decay, delta)
