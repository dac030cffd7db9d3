"""Pyramidal Lucas-Kanade on device, same signature as ``flow/methods/lukas_kanade.py:9-14``."""
from ... import ops


def calc_optical_flow_lukas_kanade(prev_grey, next_grey, win_size, max_level, step, _cache={}):
    h, w = prev_grey.shape
    key = (h, w, win_size, max_level, step)
    if key not in _cache:
        _cache.clear()
        _cache[key] = ops.LucasKanade(h, w, win_size, max_level, step)
    return _cache[key](prev_grey, next_grey)
