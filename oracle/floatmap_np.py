"""CPU restatement (TEST INFRASTRUCTURE ONLY) of the reference's WebGL variant: float displacement-map
accumulation and bilinear / nearest remap.

Follows ``extra/www/shaders/acc.frag:17-41`` (box blur of the flow, ``u = feedback(uv + flow) + flow``, decay) and
``extra/www/shaders/remap.frag:10-18`` (``bitmap(uv + mapping(uv))``) with the sampler state of
``extra/www/transflow.js:358-364`` (CLAMP_TO_EDGE, LINEAR or NEAREST), restated in PIXEL units: a texture
coordinate ``uv`` of pixel ``p`` is ``(p + 0.5) / size``, so an offset of ``f`` pixels samples the texel grid at
index ``p + f``.  The mirror flags and GL's bottom-up row order are presentation details and are left out.

Parity unpinned: the Python reference has no counterpart of this mode (its compositor is integer / nearest
only) and the shaders cannot run here; this file IS the specification the CUDA kernels are checked against.
"""
import numpy as np


def sample(tex: np.ndarray, x: np.ndarray, y: np.ndarray, linear: bool) -> np.ndarray:
    """``texture2D`` at texel-index coordinates (x, y) (texel centres at integers), CLAMP_TO_EDGE."""
    h, w = tex.shape[:2]
    x = np.asarray(x, np.float32)
    y = np.asarray(y, np.float32)
    if not linear:
        xi = np.clip(np.floor(x + np.float32(0.5)).astype(np.int64), 0, w - 1)
        yi = np.clip(np.floor(y + np.float32(0.5)).astype(np.int64), 0, h - 1)
        return tex[yi, xi].astype(np.float32)
    x0f, y0f = np.floor(x), np.floor(y)
    fx, fy = (x - x0f)[..., None], (y - y0f)[..., None]
    x0 = np.clip(x0f.astype(np.int64), 0, w - 1)
    x1 = np.clip(x0f.astype(np.int64) + 1, 0, w - 1)
    y0 = np.clip(y0f.astype(np.int64), 0, h - 1)
    y1 = np.clip(y0f.astype(np.int64) + 1, 0, h - 1)
    t = tex.astype(np.float32).reshape(h, w, -1)
    top = t[y0, x0] * (1 - fx) + t[y0, x1] * fx
    bot = t[y1, x0] * (1 - fx) + t[y1, x1] * fx
    return (top * (1 - fy) + bot * fy).astype(np.float32)


def accumulate(map_prev: np.ndarray, flow: np.ndarray, scale=1.0, decay=0.0, blur_size=1, linear=True) -> np.ndarray:
    """acc.frag: ``map(p) = v(u)``, ``u = map_prev(p + f) + f``, ``f = scale * boxblur(flow)(p)``."""
    h, w = flow.shape[:2]
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
    half = np.float32(blur_size // 2)
    weight = np.float32(1.0 / float(blur_size) ** 2)
    color = np.zeros((h, w, 2), np.float32)
    for j in range(blur_size):          # acc.frag:24-35 (offsets land on texel centres)
        for i in range(blur_size):
            color += sample(flow, xs + (np.float32(j) - half), ys + (np.float32(i) - half), linear) * weight
    f = np.float32(scale) * color
    u = sample(map_prev, xs + f[..., 0], ys + f[..., 1], linear) + f      # acc.frag:39
    return (u - np.sign(u) * np.float32(decay) * np.abs(u)).astype(np.float32)   # acc.frag:40


def remap(mapping: np.ndarray, bitmap: np.ndarray, linear=True) -> np.ndarray:
    """remap.frag: ``out(p) = bitmap(p + mapping(p))`` -> uint8 (H, W, C), rounded like an 8-bit framebuffer."""
    h, w = mapping.shape[:2]
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
    v = sample(bitmap, xs + mapping[..., 0], ys + mapping[..., 1], linear)
    return np.clip(np.floor(v + np.float32(0.5)), 0, 255).astype(np.uint8)
