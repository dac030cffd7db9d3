"""Stage-by-stage NumPy restatement of ``cv2.calcOpticalFlowFarneback`` (CPU path, flags=0).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The algorithm lives in the third-party
dependency ``opencv-python`` (``modules/video/src/optflowgf.cpp``, OpenCV 4.x; the image ships
4.13.0), not under ``/root/reference``; the reference reaches it at
``transflow/flow/sources/cv.py:479-490``.  This file restates the published algorithm so a
failing CUDA stage can be localised; ``tests/test_oracle.py`` pins it against the live ``cv2``.

Every stage mirrors the precision cv2 uses (float32 storage, float64 accumulators in the
horizontal poly-expansion pass and in the box sums).
"""
import numpy as np


def cv_round(x: float) -> int:
    """cvRound: round half to even (SSE2 cvtsd2si)."""
    return int(np.rint(x))


def level_plan(width: int, height: int, pyr_scale: float, levels: int):
    """optflowgf.cpp calc(): level crop (min_size 32) and per-level (w, h, ksz, sigma).

    Returns a list ordered coarse -> fine of dicts(k, w, h, ksz, sigma, scale).
    """
    k, scale = 0, 1.0
    while k < levels:
        scale *= pyr_scale
        if width * scale < 32 or height * scale < 32:
            break
        k += 1
    plan = []
    for lvl in range(k, -1, -1):
        scale = 1.0
        for _ in range(lvl):
            scale *= pyr_scale
        sigma = (1.0 / scale - 1.0) * 0.5
        ksz = max(cv_round(sigma * 5) | 1, 3)
        plan.append(dict(k=lvl, w=cv_round(width * scale), h=cv_round(height * scale),
                         ksz=ksz, sigma=sigma, scale=scale))
    return plan


def gaussian_kernel(ksz: int, sigma: float) -> np.ndarray:
    """cv::getGaussianKernel(ksz, sigma, CV_32F): fixed tables for sigma<=0 and ksz<=7."""
    small = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
             7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}
    if sigma <= 0 and ksz in small:
        return np.asarray(small[ksz], np.float32)
    sig = sigma if sigma > 0 else ((ksz - 1) * 0.5 - 1) * 0.3 + 0.8
    x = np.arange(ksz, dtype=np.float64) - (ksz - 1) * 0.5
    g = np.exp(-0.5 / (sig * sig) * x * x)
    g /= g.sum()
    return g.astype(np.float32)


def _reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    if n == 1:
        return np.zeros_like(idx)
    p = 2 * (n - 1)
    idx = np.abs(idx) % p
    return np.where(idx >= n, p - idx, idx)


def gaussian_blur(img: np.ndarray, ksz: int, sigma: float) -> np.ndarray:
    """Separable f32 Gaussian, BORDER_REFLECT_101 (cv::GaussianBlur on CV_32F)."""
    g = gaussian_kernel(ksz, sigma)
    r = ksz // 2
    h, w = img.shape
    cols = _reflect101(np.arange(-r, w + r), w)
    rows = _reflect101(np.arange(-r, h + r), h)
    tmp = np.zeros((h, w), np.float32)
    ext = img[:, cols]
    for j in range(ksz):
        tmp += g[j] * ext[:, j:j + w]
    out = np.zeros((h, w), np.float32)
    ext = tmp[rows, :]
    for j in range(ksz):
        out += g[j] * ext[j:j + h, :]
    return out


def linear_coeffs(dst: int, src: int):
    """cv::resize INTER_LINEAR source index / weight table for one axis (float32 weights)."""
    scale = src / dst
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    t = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s[lo], t[lo] = 0, 0
    hi = s >= src - 1
    s[hi], t[hi] = src - 1, 0
    return s, t


def resize_linear(img: np.ndarray, w: int, h: int) -> np.ndarray:
    """cv::resize(..., INTER_LINEAR) for float32, 1 or more channels (last axis)."""
    sh, sw = img.shape[:2]
    if (sw, sh) == (w, h):
        return img.astype(np.float32).copy()
    sx, tx = linear_coeffs(w, sw)
    sy, ty = linear_coeffs(h, sh)
    sx1 = np.minimum(sx + 1, sw - 1)
    sy1 = np.minimum(sy + 1, sh - 1)
    if img.ndim == 3:
        tx = tx[None, :, None]
        ty = ty[:, None, None]
    else:
        tx = tx[None, :]
        ty = ty[:, None]
    one = np.float32(1)
    rows0 = img[sy][:, sx] * (one - tx) + img[sy][:, sx1] * tx
    rows1 = img[sy1][:, sx] * (one - tx) + img[sy1][:, sx1] * tx
    return (rows0 * (one - ty) + rows1 * ty).astype(np.float32)


def pyramid_image(gray_u8: np.ndarray, lvl: dict) -> np.ndarray:
    """calc(): convertTo(f32) -> GaussianBlur(ksz, sigma) of the FULL-RES image -> resize."""
    f = gaussian_blur(gray_u8.astype(np.float32), lvl["ksz"], lvl["sigma"])
    return resize_linear(f, lvl["w"], lvl["h"])


def prepare_gaussian(n: int, sigma: float):
    """FarnebackPrepareGaussian: g, xg, xxg (f32, index 0 = x=-n) and the four inverse-Gram terms."""
    if sigma < np.finfo(np.float32).eps:
        sigma = n * 0.3
    x = np.arange(-n, n + 1, dtype=np.float64)
    g = np.exp(-x * x / (2 * sigma * sigma)).astype(np.float32)
    s = 1.0 / g.astype(np.float64).sum()
    g = (g.astype(np.float64) * s).astype(np.float32)
    xg = (x * g.astype(np.float64)).astype(np.float32)
    xxg = (x * x * g.astype(np.float64)).astype(np.float32)
    gd = g.astype(np.float64)
    gg = np.outer(gd, gd)  # [y, x]
    X = x[None, :]
    Y = x[:, None]
    G = np.zeros((6, 6))
    G[0, 0] = gg.sum()
    G[1, 1] = (gg * X * X).sum()
    G[3, 3] = (gg * X ** 4).sum()
    G[5, 5] = (gg * X * X * Y * Y).sum()
    G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
    G[4, 4] = G[3, 3]
    G[3, 4] = G[4, 3] = G[5, 5]
    inv = np.linalg.inv(G)
    return g, xg, xxg, inv[1, 1], inv[0, 3], inv[3, 3], inv[5, 5]


def poly_exp(img: np.ndarray, n: int, sigma: float) -> np.ndarray:
    """FarnebackPolyExp: (h, w) f32 -> (h, w, 5) f32 = (d/dy, d/dx, yy, xx, xy) coefficients."""
    h, w = img.shape
    g, xg, xxg, ig11, ig03, ig33, ig55 = prepare_gaussian(n, sigma)
    c = n  # centre index
    rows = np.clip(np.arange(-n, h + n), 0, h - 1)
    ext = img[rows, :]
    r0 = (ext[n:n + h] * g[c]).astype(np.float32)
    r1 = np.zeros((h, w), np.float32)
    r2 = np.zeros((h, w), np.float32)
    for k in range(1, n + 1):
        up = ext[n - k:n - k + h]
        dn = ext[n + k:n + k + h]
        p = up + dn
        r0 = r0 + g[c + k] * p
        r1 = r1 + xg[c + k] * (dn - up)
        r2 = r2 + xxg[c + k] * p
    cols = np.clip(np.arange(-n, w + n), 0, w - 1)
    e0, e1, e2 = r0[:, cols], r1[:, cols], r2[:, cols]
    b1 = e0[:, n:n + w].astype(np.float64) * float(g[c])
    b3 = e1[:, n:n + w].astype(np.float64) * float(g[c])
    b5 = e2[:, n:n + w].astype(np.float64) * float(g[c])
    b2 = np.zeros((h, w))
    b4 = np.zeros((h, w))
    b6 = np.zeros((h, w))
    for k in range(1, n + 1):
        gk, xgk, xxgk = float(g[c + k]), float(xg[c + k]), float(xxg[c + k])
        lo0, hi0 = e0[:, n - k:n - k + w], e0[:, n + k:n + k + w]
        lo1, hi1 = e1[:, n - k:n - k + w], e1[:, n + k:n + k + w]
        lo2, hi2 = e2[:, n - k:n - k + w], e2[:, n + k:n + k + w]
        tg = (hi0 + lo0).astype(np.float64)  # float32 add, then widened (as in cv2)
        b1 += tg * gk
        b4 += tg * xxgk
        b2 += (hi0 - lo0).astype(np.float64) * xgk
        b3 += (hi1 + lo1).astype(np.float64) * gk
        b6 += (hi1 - lo1).astype(np.float64) * xgk
        b5 += (hi2 + lo2).astype(np.float64) * gk
    out = np.empty((h, w, 5), np.float32)
    out[..., 1] = b2 * ig11
    out[..., 0] = b3 * ig11
    out[..., 3] = b1 * ig03 + b4 * ig33
    out[..., 2] = b1 * ig03 + b5 * ig33
    out[..., 4] = b6 * ig55
    return out


BORDER = np.asarray([0.14, 0.14, 0.4472, 0.4472, 0.4472], np.float32)


def border_scale(w: int, h: int) -> np.ndarray:
    """Per-pixel attenuation used by FarnebackUpdateMatrices within 5 px of the edges."""
    def axis(n):
        s = np.ones(n, np.float32)
        for i in range(n):
            v = np.float32(1)
            if i < 5:
                v = v * BORDER[i]
            if i >= n - 5:
                v = v * BORDER[n - i - 1]
            s[i] = v
        return s
    sx, sy = axis(w), axis(h)
    # cv2 multiplies (x-left * x-right) * y-top * y-bottom, left to right, in float32
    return (sx[None, :] * sy[:, None]).astype(np.float32)


def update_matrices(R0: np.ndarray, R1: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """FarnebackUpdateMatrices over the full image: -> M (h, w, 5) float32."""
    h, w = flow.shape[:2]
    f32 = np.float32
    xs = np.arange(w, dtype=f32)[None, :]
    ys = np.arange(h, dtype=f32)[:, None]
    dx, dy = flow[..., 0].astype(f32), flow[..., 1].astype(f32)
    fx = xs + dx
    fy = ys + dy
    x1 = np.floor(fx).astype(np.int64)
    y1 = np.floor(fy).astype(np.int64)
    fx = (fx - x1.astype(f32)).astype(f32)
    fy = (fy - y1.astype(f32)).astype(f32)
    inside = (x1 >= 0) & (x1 < w - 1) & (y1 >= 0) & (y1 < h - 1)
    xc = np.clip(x1, 0, max(w - 2, 0))
    yc = np.clip(y1, 0, max(h - 2, 0))
    one = f32(1)
    a00 = (one - fx) * (one - fy)
    a01 = fx * (one - fy)
    a10 = (one - fx) * fy
    a11 = fx * fy
    if w > 1 and h > 1:
        samp = (a00[..., None] * R1[yc, xc] + a01[..., None] * R1[yc, xc + 1]
                + a10[..., None] * R1[yc + 1, xc] + a11[..., None] * R1[yc + 1, xc + 1]).astype(f32)
    else:
        samp = np.zeros_like(R0)
    r2 = np.where(inside, samp[..., 0], f32(0))
    r3 = np.where(inside, samp[..., 1], f32(0))
    r4 = np.where(inside, (R0[..., 2] + samp[..., 2]) * f32(0.5), R0[..., 2])
    r5 = np.where(inside, (R0[..., 3] + samp[..., 3]) * f32(0.5), R0[..., 3])
    r6 = np.where(inside, (R0[..., 4] + samp[..., 4]) * f32(0.25), R0[..., 4] * f32(0.5))
    r2 = (R0[..., 0] - r2) * f32(0.5)
    r3 = (R0[..., 1] - r3) * f32(0.5)
    r2 = r2 + (r4 * dy + r6 * dx)
    r3 = r3 + (r6 * dy + r5 * dx)
    sc = border_scale(w, h)
    r2, r3, r4, r5, r6 = (v * sc for v in (r2, r3, r4, r5, r6))
    M = np.empty((h, w, 5), f32)
    M[..., 0] = r4 * r4 + r6 * r6
    M[..., 1] = (r4 + r5) * r6
    M[..., 2] = r5 * r5 + r6 * r6
    M[..., 3] = r4 * r2 + r6 * r3
    M[..., 4] = r6 * r2 + r5 * r3
    return M


def blur_solve(M: np.ndarray, winsize: int) -> np.ndarray:
    """FarnebackUpdateFlow_Blur without the lagging matrix update: box sums in double, 2x2 solve."""
    h, w = M.shape[:2]
    m = winsize // 2
    rows = np.clip(np.arange(-m, h + m), 0, h - 1)
    cols = np.clip(np.arange(-m, w + m), 0, w - 1)
    Md = M.astype(np.float64)
    cs = np.concatenate([np.zeros((1, w, 5)), np.cumsum(Md[rows], axis=0)], axis=0)
    vs = cs[2 * m + 1:] - cs[:h]
    cs = np.concatenate([np.zeros((h, 1, 5)), np.cumsum(vs[:, cols], axis=1)], axis=1)
    s = (cs[:, 2 * m + 1:] - cs[:, :w]) * (1.0 / (winsize * winsize))
    g11, g12, g22, h1, h2 = (s[..., i] for i in range(5))
    idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
    flow = np.empty((h, w, 2), np.float32)
    flow[..., 0] = (g11 * h2 - g12 * h1) * idet
    flow[..., 1] = (g22 * h1 - g12 * h2) * idet
    return flow


def upsample_flow(prev_flow: np.ndarray, w: int, h: int, pyr_scale: float) -> np.ndarray:
    """calc(): resize(prevFlow, INTER_LINEAR) then flow *= 1/pyr_scale."""
    return (resize_linear(prev_flow, w, h) * np.float32(1.0 / pyr_scale)).astype(np.float32)


def farneback(prev_u8, next_u8, pyr_scale=0.5, levels=3, winsize=15, iterations=3,
              poly_n=5, poly_sigma=1.2, trace=None):
    """Full flags=0 pipeline.  ``trace`` (a dict) receives every intermediate per level."""
    H, W = prev_u8.shape
    flow = None
    for lvl in level_plan(W, H, pyr_scale, levels):
        w, h = lvl["w"], lvl["h"]
        flow = np.zeros((h, w, 2), np.float32) if flow is None else upsample_flow(flow, w, h, pyr_scale)
        I0, I1 = pyramid_image(prev_u8, lvl), pyramid_image(next_u8, lvl)
        R0, R1 = poly_exp(I0, poly_n, poly_sigma), poly_exp(I1, poly_n, poly_sigma)
        if trace is not None:
            trace[lvl["k"]] = dict(I0=I0, I1=I1, R0=R0, R1=R1, flow_init=flow.copy(), flows=[])
        M = update_matrices(R0, R1, flow)
        for it in range(iterations):
            flow = blur_solve(M, winsize)
            if trace is not None:
                trace[lvl["k"]]["flows"].append(flow.copy())
            if it < iterations - 1:
                M = update_matrices(R0, R1, flow)
    return flow
