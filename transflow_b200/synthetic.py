"""Seeded synthetic video for parity tests and ``bench.py`` (no datasets exist offline).

Follows the recipe fixed in SURVEY.md section 8(d): a multi-octave smooth-noise texture
(Gaussian-blurred white noise at sigma 2 / 8 / 24 px, weights 0.5 / 1 / 2), cropped at a
(2t, t) px offset and composed with a slow rotation + zoom so the flow is non-uniform with
|flow| <= ~4 px per frame.  Three channels share the field with gains (1.0, 0.9, 0.8) so the
BGR -> gray conversion is exercised.  Frames are BGR uint8, the layout ``cv2.VideoCapture``
hands to the reference (``transflow/flow/sources/cv.py:461``).
"""
import numpy as np


def _smooth_noise(rng, h, w, sigma):
    import cv2
    z = rng.standard_normal((h, w)).astype(np.float32)
    z = cv2.GaussianBlur(z, (0, 0), sigma, borderType=cv2.BORDER_REFLECT_101)
    return z / (z.std() + 1e-12)


def texture_canvas(height: int, width: int, margin: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    h, w = height + 2 * margin, width + 2 * margin
    field = np.zeros((h, w), np.float32)
    for sigma, weight in ((2.0, 0.5), (8.0, 1.0), (24.0, 2.0)):
        field += weight * _smooth_noise(rng, h, w, sigma)
    field -= field.min()
    field *= 255.0 / max(float(field.max()), 1e-12)
    return field


def synthetic_clip(height: int, width: int, frames: int, seed: int = 0) -> np.ndarray:
    """Return ``(frames, H, W, 3) uint8`` BGR frames."""
    import cv2
    margin = 64 * ((frames + 15) // 16)
    canvas = texture_canvas(height, width, margin, seed)
    gains = (1.0, 0.9, 0.8)
    out = np.empty((frames, height, width, 3), np.uint8)
    cx, cy = width / 2.0, height / 2.0
    for t in range(frames):
        ang = np.deg2rad(0.05 * t)
        zoom = 1.0 + 0.0002 * t
        ca, sa = np.cos(ang) / zoom, np.sin(ang) / zoom
        # dst (x, y) -> canvas position: rotate/zoom about the frame centre, then shift by the
        # margin plus the (2t, t) translation
        m = np.array([[ca, -sa, cx - ca * cx + sa * cy + margin + 2 * t],
                      [sa, ca, cy - sa * cx - ca * cy + margin + t]], np.float64)
        frame = cv2.warpAffine(canvas, m, (width, height),
                               flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                               borderMode=cv2.BORDER_REFLECT_101)
        for c, gain in enumerate(gains):
            out[t, :, :, c] = np.clip(np.rint(frame * gain), 0, 255).astype(np.uint8)
    return out


def radial_mask(height: int, width: int) -> np.ndarray:
    """Float32 radial gradient in [0, 1] quantised to 8 bits (stand-in for assets/Mask.png)."""
    y = (np.arange(height, dtype=np.float32) - height / 2) / (height / 2)
    x = (np.arange(width, dtype=np.float32) - width / 2) / (width / 2)
    r = np.sqrt(x[None, :] ** 2 + y[:, None] ** 2) / np.sqrt(2.0)
    q = np.clip(np.rint(r * 255), 0, 255).astype(np.uint8)
    return (q.astype(np.float32) / 255).astype(np.float32)


def cnoise_pixmap(height: int, width: int, seed: int = 0) -> np.ndarray:
    """Seeded colour-noise pixmap, uint8 (H, W, 3)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(height, width, 3), dtype=np.uint8)
