"""Farneback on device: the call ``cv2.calcOpticalFlowFarneback(prev, next, flow, ...)`` of
``flow/sources/cv.py:477-490`` as a function of two device gray frames."""
from ... import ops


def calc_optical_flow_farneback(prev_grey, next_grey, pyr_scale=0.5, levels=3, winsize=15, iterations=3,
                                poly_n=5, poly_sigma=1.2, flags=0, _cache={}):
    h, w = prev_grey.shape
    key = (h, w, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags)
    if key not in _cache:
        _cache.clear()
        _cache[key] = ops.Farneback(h, w, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags)
    return _cache[key](prev_grey, next_grey)
