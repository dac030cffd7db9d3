"""NumPy restatement of ``cv2.calcOpticalFlowPyrLK`` (u8, 1 channel, flags 0, criteria
COUNT+EPS 30 / 0.01, minEigThreshold 1e-4) as used by the reference at
``transflow/flow/methods/lukas_kanade.py:26-32``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The algorithm lives in the third-party
dependency ``opencv-python`` (``modules/video/src/lkpyramid.cpp``, OpenCV 4.x; image ships
4.13.0); this file restates it so the CUDA tracker can be checked stage by
stage, and is itself pinned against the live ``cv2`` in ``tests/test_oracle.py``.
All points are processed at once (vectorised over points).
"""
import numpy as np

W_BITS = 14


def _reflect101(i, n):
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.abs(i) % p
    return np.where(i >= n, p - i, i)


def pyr_down(img: np.ndarray) -> np.ndarray:
    """cv::pyrDown for uint8: separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int32)
    src = img.astype(np.int32)
    cols = _reflect101(2 * np.arange(ow)[:, None] + np.arange(-2, 3)[None, :], w)   # (ow, 5)
    hsum = (src[:, cols] * k).sum(axis=2)                                           # (h, ow)
    rows = _reflect101(2 * np.arange(oh)[:, None] + np.arange(-2, 3)[None, :], h)   # (oh, 5)
    vsum = (hsum[rows, :] * k[None, :, None]).sum(axis=1)                           # (oh, ow)
    return ((vsum + 128) >> 8).astype(np.uint8)


def build_pyramid(img: np.ndarray, win: int, max_level: int):
    """buildOpticalFlowPyramid: stop when the NEXT level would be <= the window."""
    pyr = [img]
    h, w = img.shape
    for _ in range(max_level):
        w, h = (w + 1) // 2, (h + 1) // 2
        if w <= win or h <= win:
            break
        pyr.append(pyr_down(pyr[-1]))
    return pyr


def scharr(img: np.ndarray):
    """ScharrDerivInvoker: int16 (Ix, Iy), BORDER_REFLECT_101 on the level image itself."""
    h, w = img.shape
    s = img.astype(np.int32)
    ru = s[_reflect101(np.arange(h) - 1, h)]
    rd = s[_reflect101(np.arange(h) + 1, h)]
    t0 = (ru + rd) * 3 + s * 10
    t1 = rd - ru
    cl = _reflect101(np.arange(w) - 1, w)
    cr = _reflect101(np.arange(w) + 1, w)
    ix = t0[:, cr] - t0[:, cl]
    iy = (t1[:, cr] + t1[:, cl]) * 3 + t1 * 10
    return ix.astype(np.int16), iy.astype(np.int16)


def _descale(v, n):
    return (v + (1 << (n - 1))) >> n


def _weights(a, b):
    f = np.float32
    one = f(1)
    s = f(1 << W_BITS)
    iw00 = np.rint((one - a) * (one - b) * s).astype(np.int32)
    iw01 = np.rint(a * (one - b) * s).astype(np.int32)
    iw10 = np.rint((one - a) * b * s).astype(np.int32)
    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10
    return iw00, iw01, iw10, iw11


def _patch(padded, ix, iy, win, pad):
    """(npts, win+1, win+1) int32 window at integer top-left (ix, iy) of a padded image."""
    r = np.arange(win + 1)
    yy = (iy[:, None] + r[None, :] + pad)[:, :, None]
    xx = (ix[:, None] + r[None, :] + pad)[:, None, :]
    return padded[yy, xx].astype(np.int32)


def _interp(p, w, shift):
    iw00, iw01, iw10, iw11 = (v[:, None, None] for v in w)
    v = p[:, :-1, :-1] * iw00 + p[:, :-1, 1:] * iw01 + p[:, 1:, :-1] * iw10 + p[:, 1:, 1:] * iw11
    return _descale(v, shift)


def track_level(I, J, prev_pts, next_pts, win, level, is_top, max_count=30, eps=0.01, min_eig_thr=1e-4):
    """LKTrackerInvoker for one pyramid level; returns the updated nextPts (float32 (n, 2))."""
    f = np.float32
    h, w = I.shape
    half = f((win - 1) * 0.5)
    eps2 = float(min(max(eps, 0.0), 10.0)) ** 2
    max_count = min(max(max_count, 0), 100)
    Ipad = np.pad(I, win, mode="reflect")
    Jpad = np.pad(J, win, mode="reflect")
    dx, dy = scharr(I)
    dxp = np.pad(dx, win, mode="constant")
    dyp = np.pad(dy, win, mode="constant")

    prev = (prev_pts * f(1.0 / (1 << level))).astype(f)
    nxt = prev.copy() if is_top else (next_pts * f(2)).astype(f)
    out = nxt.copy()                       # nextPts[ptidx] = nextPt (stored before any test)
    pp = prev - half
    ipx = np.floor(pp[:, 0]).astype(np.int32)
    ipy = np.floor(pp[:, 1]).astype(np.int32)
    ok = ~((ipx < -win) | (ipx >= w) | (ipy < -win) | (ipy >= h))
    idx = np.nonzero(ok)[0]
    if idx.size == 0:
        return out
    a = (pp[idx, 0] - ipx[idx].astype(f)).astype(f)
    b = (pp[idx, 1] - ipy[idx].astype(f)).astype(f)
    wts = _weights(a, b)
    Iwin = _interp(_patch(Ipad, ipx[idx], ipy[idx], win, win), wts, W_BITS - 5)
    dIx = _interp(_patch(dxp, ipx[idx], ipy[idx], win, win), wts, W_BITS)
    dIy = _interp(_patch(dyp, ipx[idx], ipy[idx], win, win), wts, W_BITS)
    scale = f(1.0 / (1 << 20))
    A11 = (dIx * dIx).sum(axis=(1, 2)).astype(f) * scale
    A12 = (dIx * dIy).sum(axis=(1, 2)).astype(f) * scale
    A22 = (dIy * dIy).sum(axis=(1, 2)).astype(f) * scale
    D = A11 * A22 - A12 * A12
    min_eig = (A22 + A11 - np.sqrt((A11 - A22) * (A11 - A22) + f(4) * A12 * A12)) / f(2 * win * win)
    good = ~((min_eig < f(min_eig_thr)) | (D < np.finfo(f).eps))
    idx, Iwin, dIx, dIy = idx[good], Iwin[good], dIx[good], dIy[good]
    A11, A12, A22, D = A11[good], A12[good], A22[good], D[good]
    Dinv = (f(1) / D).astype(f)
    npt = (nxt[idx] - half).astype(f)
    prev_delta = np.zeros_like(npt)
    active = np.ones(idx.size, bool)
    for j in range(max_count):
        if not active.any():
            break
        act = np.nonzero(active)[0]
        inx = np.floor(npt[act, 0]).astype(np.int32)
        iny = np.floor(npt[act, 1]).astype(np.int32)
        oob = (inx < -win) | (inx >= w) | (iny < -win) | (iny >= h)
        active[act[oob]] = False
        act, inx, iny = act[~oob], inx[~oob], iny[~oob]
        if act.size == 0:
            break
        a = (npt[act, 0] - inx.astype(f)).astype(f)
        b = (npt[act, 1] - iny.astype(f)).astype(f)
        Jwin = _interp(_patch(Jpad, inx, iny, win, win), _weights(a, b), W_BITS - 5)
        diff = Jwin - Iwin[act]
        b1 = (diff * dIx[act]).sum(axis=(1, 2)).astype(f) * scale
        b2 = (diff * dIy[act]).sum(axis=(1, 2)).astype(f) * scale
        delta = np.stack([(A12[act] * b2 - A22[act] * b1) * Dinv[act],
                          (A12[act] * b1 - A11[act] * b2) * Dinv[act]], axis=1).astype(f)
        npt[act] = npt[act] + delta
        out[idx[act]] = npt[act] + half
        small = (delta.astype(np.float64) ** 2).sum(axis=1) <= eps2
        osc = np.zeros(act.size, bool)
        if j > 0:
            osc = (~small & (np.abs(delta[:, 0] + prev_delta[act, 0]) < 0.01)
                   & (np.abs(delta[:, 1] + prev_delta[act, 1]) < 0.01))
            out[idx[act[osc]]] = out[idx[act[osc]]] - delta[osc] * f(0.5)
        prev_delta[act] = delta
        active[act[small | osc]] = False
    return out


def pyr_lk(prev_img, next_img, pts, win=15, max_level=2):
    """``cv2.calcOpticalFlowPyrLK(prev, next, pts, None, winSize=(win, win), maxLevel=max_level)[0]``."""
    pa = build_pyramid(prev_img, win, max_level)
    pb = build_pyramid(next_img, win, max_level)
    top = len(pa) - 1
    pts = np.asarray(pts, np.float32).reshape(-1, 2)
    nxt = pts.copy()
    for level in range(top, -1, -1):
        nxt = track_level(pa[level], pb[level], pts, nxt, win, level, level == top)
    return nxt


def dense_flow(prev_img, next_img, win=15, max_level=2, step=1):
    """The reference wrapper (lukas_kanade.py:9-36) on top of the restated tracker."""
    h, w = prev_img.shape
    gx, gy = np.meshgrid(np.arange(0, w, step), np.arange(0, h, step), indexing="xy")
    p0 = np.stack([gx, gy], axis=-1).astype(np.float32)
    rows, cols = p0.shape[:2]
    p1 = pyr_lk(prev_img, next_img, p0.reshape(-1, 2), win, max_level)
    flow = (p1 - p0.reshape(-1, 2)).reshape(rows, cols, 2)
    if step == 1:
        return flow
    return np.repeat(np.repeat(flow, step, axis=0), step, axis=1)[:h, :w].astype(np.float32)
